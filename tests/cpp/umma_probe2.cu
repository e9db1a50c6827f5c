// Second UMMA probe (dual build like umma_probe.cu: nvcc for a B200, g++ -DLYRA_EMU for the emulator): the tcgen05 features
// the product kernels of round 2 rely on beyond what umma_probe.cu pins.
//   case 3: A operand in TENSOR MEMORY (written by tcgen05.st, one row per thread), split-precision TF32 (3 MMAs per product),
//           B from shared memory, D read back with 16-column tcgen05.ld; M = 128, N = 64, K = 64.
//   case 4: the same with only TMEM lanes 0..31 of A written (rows 0..31 valid, the rest stale) - rows of a GEMM are independent,
//           so rows 0..31 must still match (this is how a 160-row tile uses a second, mostly empty, 128-row block).
//   case 5: kind::i8, both operands in shared memory (K-major, 16-byte core-matrix rows), s32 accumulators; exact.
//   case 6: kind::i8 with the A operand in tensor memory (4 int8 per 32-bit column); exact.
//   case 7: TF32 with N = 160 and N = 16 (the decoder_2/simple half-width and last_layer shapes), A from shared memory.
//   case 8: bulk asynchronous store shared -> global (cp.async.bulk.global.shared::cta) round trip.
// Exit code 0 iff every case matches.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "device_compat.h"

namespace {

constexpr int K = 64;

struct ProbeShared { LyraMbar bar; uint32_t tmem_base; };

__device__ inline int CanonF32(int row, int k, int rows) { return ((k / 4) * (rows / 8) + row / 8) * 32 + (row % 8) * 4 + k % 4; }
__device__ inline int CanonI8(int row, int k, int rows) { return ((k / 16) * (rows / 8) + row / 8) * 128 + (row % 8) * 16 + k % 16; }   // byte index
__device__ inline void Split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(__fsub_rn(x, __uint_as_float(hi)));
}

// mode 3 / 4: OUT[128 x 64] = A * W^T (A in TMEM, split precision); mode 4 writes only lanes 0..31 of A
// mode 7: OUT[128 x (160 | 16)] = A * W^T, A and W in shared memory, single pass (operands pre-truncated by the host check)
__global__ void __launch_bounds__(128)
ProbeF32Kernel(int mode, int N, const float* A, const float* W, float* OUT) {
  float* w_hi = reinterpret_cast<float*>(LYRA_DYN_SMEM());
  float* w_lo = w_hi + N * K;
  float* a_sm = w_lo + N * K;                       // mode 7 only
  LYRA_STATIC_SMEM(ProbeShared, sh, 1);
  const int tid = (int)threadIdx.x, warp = tid / 32;
  for (int i = tid; i < N * K; i += 128) {
    uint32_t h, l;
    Split(W[i], h, l);
    w_hi[CanonF32(i / K, i % K, N)] = __uint_as_float(h);
    w_lo[CanonF32(i / K, i % K, N)] = __uint_as_float(l);
  }
  if (mode == 7) for (int i = tid; i < 128 * K; i += 128) a_sm[CanonF32(i / K, i % K, 128)] = A[i];
  if (tid == 0) { lyra_mbar_init(&sh->bar, 1); lyra_mbar_fence_init(); }
  lyra_fence_proxy_async();
  if (warp == 0) lyra_tmem_alloc(&sh->tmem_base, 512);
  lyra_tc_fence_before_sync();
  __syncthreads();
  lyra_tc_fence_after_sync();
  const uint32_t tmem = sh->tmem_base;
  const uint32_t colD = 0, colAhi = 256, colAlo = 256 + 64;
  const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
  if (mode != 7 && (mode == 3 || warp == 0)) {
    // one row per thread: split and store 16 columns at a time
    for (int c0 = 0; c0 < K; c0 += 16) {
      uint32_t hi[16], lo[16];
      for (int j = 0; j < 16; ++j) Split(A[tid * K + c0 + j], hi[j], lo[j]);
      lyra_tmem_st<16>(tmem + lane_base + colAhi + (uint32_t)c0, hi);
      lyra_tmem_st<16>(tmem + lane_base + colAlo + (uint32_t)c0, lo);
    }
    lyra_tmem_wait_st();
  }
  lyra_tc_fence_before_sync();
  __syncthreads();
  lyra_tc_fence_after_sync();
  if (tid == 0) {
    const uint32_t idesc = lyra_umma_idesc_tf32(128, N);
    const uint32_t lboW = (uint32_t)(N / 8) * 128u, lboA = 16u * 128u;
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t bh = lyra_umma_desc(reinterpret_cast<const char*>(w_hi) + (size_t)ks * 2 * lboW, lboW, 128);
      const uint64_t bl = lyra_umma_desc(reinterpret_cast<const char*>(w_lo) + (size_t)ks * 2 * lboW, lboW, 128);
      if (mode == 7) {
        lyra_umma_tf32(tmem + colD, lyra_umma_desc(reinterpret_cast<const char*>(a_sm) + (size_t)ks * 2 * lboA, lboA, 128), bh, idesc, ks > 0);
      } else {
        lyra_umma_tf32_ts(tmem + colD, tmem + colAlo + (uint32_t)(8 * ks), bh, idesc, ks > 0);
        lyra_umma_tf32_ts(tmem + colD, tmem + colAhi + (uint32_t)(8 * ks), bl, idesc, true);
        lyra_umma_tf32_ts(tmem + colD, tmem + colAhi + (uint32_t)(8 * ks), bh, idesc, true);
      }
    }
    lyra_umma_commit(&sh->bar);
  }
  lyra_mbar_wait(&sh->bar, 0);
  lyra_tc_fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    lyra_tmem_ld<16>(tmem + lane_base + colD + (uint32_t)c0, v);
    lyra_tmem_wait_ld();
    for (int j = 0; j < 16; ++j) OUT[tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  lyra_tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) lyra_tmem_dealloc(tmem, 512);
}

// mode 5: A (s8) from shared memory; mode 6: A from TMEM.  OUT (s32) [128 x 64] = A [128 x 64] * W [64 x 64]^T
__global__ void __launch_bounds__(128)
ProbeI8Kernel(int mode, const int8_t* A, const int8_t* W, int* OUT) {
  constexpr int N = 64;
  int8_t* w_sm = reinterpret_cast<int8_t*>(LYRA_DYN_SMEM());
  int8_t* a_sm = w_sm + N * K;
  LYRA_STATIC_SMEM(ProbeShared, sh, 1);
  const int tid = (int)threadIdx.x, warp = tid / 32;
  for (int i = tid; i < N * K; i += 128) w_sm[CanonI8(i / K, i % K, N)] = W[i];
  for (int i = tid; i < 128 * K; i += 128) a_sm[CanonI8(i / K, i % K, 128)] = A[i];
  if (tid == 0) { lyra_mbar_init(&sh->bar, 1); lyra_mbar_fence_init(); }
  lyra_fence_proxy_async();
  if (warp == 0) lyra_tmem_alloc(&sh->tmem_base, 128);
  lyra_tc_fence_before_sync();
  __syncthreads();
  lyra_tc_fence_after_sync();
  const uint32_t tmem = sh->tmem_base, colD = 0, colA = 64;
  const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
  if (mode == 6) {
    uint32_t w[16];          // K = 64 int8 = 16 columns
    for (int j = 0; j < 16; ++j) {
      uint32_t x = 0;
      for (int b = 0; b < 4; ++b) x |= (uint32_t)(uint8_t)A[tid * K + 4 * j + b] << (8 * b);
      w[j] = x;
    }
    lyra_tmem_st<16>(tmem + lane_base + colA, w);
    lyra_tmem_wait_st();
  }
  lyra_tc_fence_before_sync();
  __syncthreads();
  lyra_tc_fence_after_sync();
  if (tid == 0) {
    const uint32_t idesc = lyra_umma_idesc_i8(128, N);
    const uint32_t lboW = (uint32_t)(N / 8) * 128u, lboA = 16u * 128u;
    for (int ks = 0; ks < K / 32; ++ks) {
      const uint64_t b = lyra_umma_desc(reinterpret_cast<const char*>(w_sm) + (size_t)ks * 2 * lboW, lboW, 128);
      if (mode == 5) lyra_umma_i8(tmem + colD, lyra_umma_desc(reinterpret_cast<const char*>(a_sm) + (size_t)ks * 2 * lboA, lboA, 128), b, idesc, ks > 0);
      else lyra_umma_i8_ts(tmem + colD, tmem + colA + (uint32_t)(8 * ks), b, idesc, ks > 0);
    }
    lyra_umma_commit(&sh->bar);
  }
  lyra_mbar_wait(&sh->bar, 0);
  lyra_tc_fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    lyra_tmem_ld<16>(tmem + lane_base + colD + (uint32_t)c0, v);
    lyra_tmem_wait_ld();
    for (int j = 0; j < 16; ++j) OUT[tid * N + c0 + j] = (int)v[j];
  }
  lyra_tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) lyra_tmem_dealloc(tmem, 128);
}

__global__ void __launch_bounds__(128)
ProbeBulkStoreKernel(const float* in, float* out) {
  float* buf = reinterpret_cast<float*>(LYRA_DYN_SMEM());
  const int tid = (int)threadIdx.x;
  for (int i = tid; i < 2048; i += 128) buf[i] = in[i] * 2.0f + 1.0f;
  lyra_fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    lyra_bulk_s2g(out, buf, 4096u);
    lyra_bulk_s2g(out + 1024, buf + 1024, 4096u);
    lyra_bulk_commit();
    lyra_bulk_wait_all();
  }
}

float Tf32(float x) { uint32_t b; std::memcpy(&b, &x, 4); b &= 0xffffe000u; std::memcpy(&x, &b, 4); return x; }

template <typename T>
T* ToDevice(const std::vector<T>& v) {
  void* p = nullptr;
  if (cudaMalloc(&p, v.size() * sizeof(T)) != cudaSuccess) return nullptr;
  cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return static_cast<T*>(p);
}

}  // namespace

int main() {
  int bad = 0;
  srand(11);
  auto rnd = [] { return (float)(rand() % 20001 - 10000) / 10000.0f * 1.37f; };
  {
    std::vector<float> A(128 * K), W(160 * K), OUT(128 * 160);
    for (auto& v : A) v = rnd();
    for (auto& v : W) v = rnd() * 0.25f;
    float *dA = ToDevice(A), *dW = ToDevice(W), *dO = ToDevice(OUT);
    if (!dA || !dW || !dO) { std::printf("allocation failed\n"); return 2; }
    struct Cfg { int mode, N, rows; bool split; const char* name; };
    const Cfg cfgs[] = {{3, 64, 128, true, "case 3 (A in TMEM, split TF32)"}, {4, 64, 32, true, "case 4 (A in TMEM, lanes 0..31 only)"},
                        {7, 160, 128, false, "case 7a (N = 160)"}, {7, 16, 128, false, "case 7b (N = 16)"}};
    for (const Cfg& c : cfgs) {
      const size_t smem = (size_t)(2 * c.N * K + 128 * K) * 4;
      LYRA_SET_MAX_SMEM(ProbeF32Kernel, smem);
      cudaMemset(dO, 0, OUT.size() * 4);
      LYRA_LAUNCH(ProbeF32Kernel, dim3(1), dim3(128), smem, 0, c.mode, c.N, dA, dW, dO);
      if (cudaDeviceSynchronize() != cudaSuccess) { std::printf("%s: kernel failed\n", c.name); return 1; }
      cudaMemcpy(OUT.data(), dO, OUT.size() * 4, cudaMemcpyDeviceToHost);
      double worst = 0, scale = 0;
      for (int m = 0; m < c.rows; ++m)
        for (int n = 0; n < c.N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k)
            ref += c.split ? (double)A[m * K + k] * (double)W[n * K + k] : (double)Tf32(A[m * K + k]) * (double)Tf32(W[n * K + k]);
          worst = std::fmax(worst, std::fabs((double)OUT[m * c.N + n] - ref));
          scale = std::fmax(scale, std::fabs(ref));
        }
      const bool ok = worst / scale < 5e-6;
      std::printf("%s: max |OUT - ref| = %.3e (relative %.2e) -> %s\n", c.name, worst, worst / scale, ok ? "MATCH" : "MISMATCH");
      bad |= !ok;
    }
  }
  {
    std::vector<int8_t> A(128 * K), W(64 * K);
    std::vector<int> OUT(128 * 64);
    for (auto& v : A) v = (int8_t)(rand() % 256 - 128);
    for (auto& v : W) v = (int8_t)(rand() % 256 - 128);
    int8_t *dA = ToDevice(A), *dW = ToDevice(W);
    int* dO = ToDevice(OUT);
    if (!dA || !dW || !dO) { std::printf("allocation failed\n"); return 2; }
    for (int mode = 5; mode <= 6; ++mode) {
      const size_t smem = (size_t)(64 * K + 128 * K);
      cudaMemset(dO, 0, OUT.size() * 4);
      LYRA_LAUNCH(ProbeI8Kernel, dim3(1), dim3(128), smem, 0, mode, dA, dW, dO);
      if (cudaDeviceSynchronize() != cudaSuccess) { std::printf("case %d: kernel failed\n", mode); return 1; }
      cudaMemcpy(OUT.data(), dO, OUT.size() * 4, cudaMemcpyDeviceToHost);
      int nbad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          int ref = 0;
          for (int k = 0; k < K; ++k) ref += (int)A[m * K + k] * (int)W[n * K + k];
          nbad += OUT[m * 64 + n] != ref;
        }
      std::printf("case %d (kind::i8, A in %s): %d wrong of %d -> %s\n", mode, mode == 5 ? "shared memory" : "TMEM", nbad, 128 * 64, nbad ? "MISMATCH" : "MATCH");
      bad |= nbad != 0;
    }
  }
  {
    std::vector<float> in(2048), out(2048, 0.0f);
    for (auto& v : in) v = rnd();
    float *dI = ToDevice(in), *dO = ToDevice(out);
    LYRA_LAUNCH(ProbeBulkStoreKernel, dim3(1), dim3(128), (size_t)8192, 0, dI, dO);
    if (cudaDeviceSynchronize() != cudaSuccess) { std::printf("case 8: kernel failed\n"); return 1; }
    cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
    int nbad = 0;
    for (int i = 0; i < 2048; ++i) nbad += out[i] != in[i] * 2.0f + 1.0f;
    std::printf("case 8 (bulk store): %d wrong -> %s\n", nbad, nbad ? "MISMATCH" : "MATCH");
    bad |= nbad != 0;
  }
  return bad;
}
