// UMMA (tcgen05) groundwork probe, built two ways from this one source:
//   * nvcc -gencode arch=compute_100a,code=sm_100a   -> runs on a B200 (the instruction sequences of device_compat.h)
//   * g++ -DLYRA_EMU -x c++ (+ tests/cuda_emu)         -> runs on the CPU tier against the emulator's UMMA / TMEM model
// Case 1: D[128 x 64] = A[128 x 64] * B[64 x 64]^T, kind::tf32, single precision pass (the layout tools/tcgen05_probe.cu pinned
//         on hardware).
// Case 2: OUT = LeakyReLU(A * W1^T + b1) * W2^T + b2 + U for A, U [160 x 64]: both 1x1 convolutions of one decoder_2 residual
//         unit at 8 streams x 20 rows, split precision (three MMAs per product), M = 160 as two overlapping 128-row blocks,
//         the first epilogue rewriting the operand buffers in place.  (Hardware status: DESIGN.md section 9.)
// Exit code 0 iff both cases match their double-precision references.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "device_compat.h"

namespace {

constexpr int K = 64, N = 64;

struct ProbeShared { LyraMbar bar; uint32_t tmem_base; };

// canonical K-major no-swizzle operand: [k/4][row/8][row%8][k%4]
__device__ inline int Canon(int row, int k, int rows) { return ((k / 4) * (rows / 8) + row / 8) * 32 + (row % 8) * 4 + k % 4; }
__device__ inline void Split(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = __fsub_rn(x, hi);
}

// rows [row0, row0 + 128) of an operand with `rows` rows -> TMEM columns [col0, col0 + 64); split: three MMAs per product
__device__ inline void IssueBlock(uint32_t tmem, int col0, const float* a_hi, const float* a_lo, int rows, int row0, const float* w_hi,
                                  const float* w_lo, bool split) {
  const uint32_t idesc = lyra_umma_idesc_tf32(128, N);
  const uint32_t lboA = (uint32_t)(rows / 8) * 128u, lboW = (uint32_t)(N / 8) * 128u;
  for (int ks = 0; ks < K / 8; ++ks) {
    const char* ah = reinterpret_cast<const char*>(a_hi) + (row0 / 8) * 128 + (size_t)ks * 2 * lboA;
    const char* al = reinterpret_cast<const char*>(a_lo) + (row0 / 8) * 128 + (size_t)ks * 2 * lboA;
    const char* wh = reinterpret_cast<const char*>(w_hi) + (size_t)ks * 2 * lboW;
    const char* wl = reinterpret_cast<const char*>(w_lo) + (size_t)ks * 2 * lboW;
    const uint32_t d = tmem + (uint32_t)col0;
    if (split) {
      lyra_umma_tf32(d, lyra_umma_desc(al, lboA, 128), lyra_umma_desc(wh, lboW, 128), idesc, ks > 0);      // small terms first
      lyra_umma_tf32(d, lyra_umma_desc(ah, lboA, 128), lyra_umma_desc(wl, lboW, 128), idesc, true);
      lyra_umma_tf32(d, lyra_umma_desc(ah, lboA, 128), lyra_umma_desc(wh, lboW, 128), idesc, true);
    } else {
      lyra_umma_tf32(d, lyra_umma_desc(ah, lboA, 128), lyra_umma_desc(wh, lboW, 128), idesc, ks > 0);
    }
  }
}

// mode 0: case 1 (M = 128, single pass, OUT = A * W1^T); mode 1: case 2 (M = 160)
__global__ void __launch_bounds__(128)
UmmaProbeKernel(int mode, const float* A, const float* W1, const float* B1, const float* W2, const float* B2, const float* U, float* OUT) {
  const int M = mode == 0 ? 128 : 160;
  float* a_hi = reinterpret_cast<float*>(LYRA_DYN_SMEM());
  float* a_lo = a_hi + M * K;
  float* w1_hi = a_lo + M * K;
  float* w1_lo = w1_hi + N * K;
  float* w2_hi = w1_lo + N * K;
  float* w2_lo = w2_hi + N * K;
  LYRA_STATIC_SMEM(ProbeShared, sh, 1);
  const int tid = (int)threadIdx.x, warp = tid / 32, lane = tid % 32;
  for (int i = tid; i < M * K; i += 128) {
    if (mode == 0) { a_hi[Canon(i / K, i % K, M)] = A[i]; a_lo[Canon(i / K, i % K, M)] = 0.0f; }
    else Split(A[i], a_hi[Canon(i / K, i % K, M)], a_lo[Canon(i / K, i % K, M)]);
  }
  for (int i = tid; i < N * K; i += 128) {
    if (mode == 0) { w1_hi[Canon(i / K, i % K, N)] = W1[i]; w1_lo[Canon(i / K, i % K, N)] = 0.0f; }
    else Split(W1[i], w1_hi[Canon(i / K, i % K, N)], w1_lo[Canon(i / K, i % K, N)]);
    Split(W2[i], w2_hi[Canon(i / K, i % K, N)], w2_lo[Canon(i / K, i % K, N)]);
  }
  if (tid == 0) { lyra_mbar_init(&sh->bar, 1); lyra_mbar_fence_init(); }
  lyra_fence_proxy_async();                       // operands written with ordinary stores -> visible to the tensor-core proxy
  if (warp == 0) lyra_tmem_alloc(&sh->tmem_base, 128);
  lyra_tc_fence_before_sync();
  __syncthreads();
  lyra_tc_fence_after_sync();
  const uint32_t tmem = sh->tmem_base;
  const bool split = mode != 0;

  if (tid == 0) {
    IssueBlock(tmem, 0, a_hi, a_lo, M, 0, w1_hi, w1_lo, split);
    if (M > 128) IssueBlock(tmem, 64, a_hi, a_lo, M, 32, w1_hi, w1_lo, split);
    lyra_umma_commit(&sh->bar);
  }
  lyra_mbar_wait(&sh->bar, 0);
  lyra_tc_fence_after_sync();
  // warp w owns rows 32w + lane of block 0 (TMEM lanes 32w..); warp 3 also rows 128 + lane of block 1 (its TMEM lanes 96..127)
  for (int pass = 0; pass < (M > 128 ? 2 : 1); ++pass) {
    if (pass == 1 && warp != 3) break;
    const int row = pass == 0 ? 32 * warp + lane : 128 + lane;
    const uint32_t lane_base = (uint32_t)(pass == 0 ? 32 * warp : 96) << 16;
    for (int c0 = 0; c0 < N; c0 += 8) {
      float v[8];
      lyra_tmem_ld8(tmem + lane_base + (uint32_t)(pass * 64 + c0), v);
      for (int j = 0; j < 8; ++j) {
        if (mode == 0) { OUT[row * N + c0 + j] = v[j]; continue; }
        float h = __fadd_rn(v[j], B1[c0 + j]);
        h = h > 0.0f ? h : __fmul_rn(h, 0.3f);
        Split(h, a_hi[Canon(row, c0 + j, M)], a_lo[Canon(row, c0 + j, M)]);      // next operand, in place
      }
    }
  }
  if (mode != 0) {
    lyra_fence_proxy_async();
    lyra_tc_fence_before_sync();
    __syncthreads();
    lyra_tc_fence_after_sync();
    if (tid == 0) {
      IssueBlock(tmem, 0, a_hi, a_lo, M, 0, w2_hi, w2_lo, true);
      IssueBlock(tmem, 64, a_hi, a_lo, M, 32, w2_hi, w2_lo, true);
      lyra_umma_commit(&sh->bar);
    }
    lyra_mbar_wait(&sh->bar, 1);
    lyra_tc_fence_after_sync();
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1 && warp != 3) break;
      const int row = pass == 0 ? 32 * warp + lane : 128 + lane;
      const uint32_t lane_base = (uint32_t)(pass == 0 ? 32 * warp : 96) << 16;
      for (int c0 = 0; c0 < N; c0 += 8) {
        float v[8];
        lyra_tmem_ld8(tmem + lane_base + (uint32_t)(pass * 64 + c0), v);
        for (int j = 0; j < 8; ++j) OUT[row * N + c0 + j] = __fadd_rn(__fadd_rn(v[j], B2[c0 + j]), U[row * N + c0 + j]);
      }
    }
  }
  lyra_tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) lyra_tmem_dealloc(tmem, 128);
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
  const int MM = 160;
  std::vector<float> A(MM * K), U(MM * N), W1(N * K), W2(N * K), B1(N), B2(N), OUT(MM * N);
  srand(7);
  auto rnd = [] { return (float)(rand() % 20001 - 10000) / 10000.0f * 1.37f; };
  for (auto& v : A) v = rnd();
  for (auto& v : U) v = rnd();
  for (auto& v : W1) v = rnd() * 0.25f;
  for (auto& v : W2) v = rnd() * 0.25f;
  for (auto& v : B1) v = rnd() * 0.1f;
  for (auto& v : B2) v = rnd() * 0.1f;
  float *dA = ToDevice(A), *dU = ToDevice(U), *dW1 = ToDevice(W1), *dW2 = ToDevice(W2), *dB1 = ToDevice(B1), *dB2 = ToDevice(B2), *dO = ToDevice(OUT);
  if (!dA || !dU || !dW1 || !dW2 || !dB1 || !dB2 || !dO) { std::printf("allocation failed\n"); return 2; }
  const size_t smem = (size_t)(2 * MM * K + 4 * N * K) * 4;
  LYRA_SET_MAX_SMEM(UmmaProbeKernel, smem);
  int bad = 0;
  for (int mode = 0; mode < 2; ++mode) {
    const int M = mode == 0 ? 128 : 160;
    LYRA_LAUNCH(UmmaProbeKernel, dim3(1), dim3(128), smem, 0, mode, dA, dW1, dB1, dW2, dB2, dU, dO);
    if (cudaDeviceSynchronize() != cudaSuccess) { std::printf("case %d: kernel failed\n", mode + 1); return 1; }
    cudaMemcpy(OUT.data(), dO, OUT.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0, scale = 0;
    for (int m = 0; m < M; ++m) {
      double h[N];
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k)
          s += mode == 0 ? (double)Tf32(A[m * K + k]) * (double)Tf32(W1[n * K + k]) : (double)A[m * K + k] * (double)W1[n * K + k];
        if (mode == 0) { h[n] = s; continue; }
        s += B1[n];
        h[n] = (double)(float)(s > 0 ? s : s * (double)0.3f);
      }
      for (int n = 0; n < N; ++n) {
        double ref = h[n];
        if (mode != 0) {
          ref = 0;
          for (int k = 0; k < K; ++k) ref += h[k] * (double)W2[n * K + k];
          ref += (double)B2[n] + (double)U[m * N + n];
        }
        worst = std::fmax(worst, std::fabs((double)OUT[m * N + n] - ref));
        scale = std::fmax(scale, std::fabs(ref));
      }
    }
    const bool ok = worst / scale < 5e-6;
    std::printf("case %d: max |OUT - ref| = %.3e (scale %.2f, relative %.2e) -> %s\n", mode + 1, worst, scale, worst / scale, ok ? "MATCH" : "MISMATCH");
    bad |= !ok;
  }
  return bad;
}
