// Dev probe (GPU, standalone), preparation for moving the decoder's tensor-core mode from mma.sync to UMMA:
// one CTA computes  OUT = LeakyReLU(A * W1^T + b1) * W2^T + b2 + U   for A, U [160 x 64], W1, W2 [64 x 64]
// — the two 1x1 convolutions of one decoder_2 residual unit at 8 streams x 20 rows — with tcgen05.mma kind::tf32 in
// split precision (x = hi + lo, three MMAs per product), operands in shared memory in the canonical no-swizzle K-major
// layout, accumulators in TMEM, M = 160 covered by two overlapping M = 128 row blocks (rows 0-127 and 32-159), and the
// epilogue of the first GEMM rewriting the operand buffers in place for the second.  Checks against a double-precision
// CPU reference (expected error ~1e-6: fp32-level).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o devtools_build/umma_resunit_probe tools/umma_resunit_probe.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

constexpr int M = 160, K = 64, N = 64;
constexpr int kABytes = M * K * 4, kWBytes = N * K * 4;
constexpr int kSmem = 2 * kABytes + 4 * kWBytes;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16 | (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32 | (uint64_t)1 << 46;
}
// canonical K-major no-swizzle: [k/4][row/8][row%8][k%4]
__device__ __forceinline__ int canon(int row, int k, int rows) { return ((k / 4) * (rows / 8) + row / 8) * 32 + (row % 8) * 4 + k % 4; }
__device__ __forceinline__ void split(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = __fsub_rn(x, hi);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DN;\nbra W;\nDN:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}

// issue one split-precision GEMM: rows [0,128) -> TMEM columns [0,64), rows [32,160) -> columns [64,128)
__device__ void issue_gemm(uint32_t tmem, const float* a_hi, const float* a_lo, const float* w_hi, const float* w_lo, unsigned long long* bar) {
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t lboA = (M / 8) * 128, lboW = (N / 8) * 128;     // bytes between k/4 groups
  for (int blk = 0; blk < 2; ++blk) {
    const uint32_t row_off = blk ? (32 / 8) * 128 : 0;           // start at row 32: 4 row groups of 128 bytes
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint32_t ao = row_off + ks * 2 * lboA, wo = ks * 2 * lboW;
      const uint64_t dah = make_desc(smem_u32(a_hi) + ao, lboA, 128), dal = make_desc(smem_u32(a_lo) + ao, lboA, 128);
      const uint64_t dwh = make_desc(smem_u32(w_hi) + wo, lboW, 128), dwl = make_desc(smem_u32(w_lo) + wo, lboW, 128);
      const uint32_t d = tmem + blk * 64;
      mma_tf32(d, dal, dwh, idesc, ks > 0 ? 1u : 0u);            // small terms first
      mma_tf32(d, dah, dwl, idesc, 1u);
      mma_tf32(d, dah, dwh, idesc, 1u);
    }
  }
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(128) probe(const float* A, const float* W1, const float* B1, const float* W2, const float* B2, const float* U, float* OUT) {
  extern __shared__ __align__(128) float dyn[];
  float* a_hi = dyn;
  float* a_lo = a_hi + M * K;
  float* w1_hi = a_lo + M * K;
  float* w1_lo = w1_hi + N * K;
  float* w2_hi = w1_lo + N * K;
  float* w2_lo = w2_hi + N * K;
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  for (int i = tid; i < M * K; i += 128) split(A[i], a_hi[canon(i / K, i % K, M)], a_lo[canon(i / K, i % K, M)]);
  for (int i = tid; i < N * K; i += 128) {
    split(W1[i], w1_hi[canon(i / K, i % K, N)], w1_lo[canon(i / K, i % K, N)]);
    split(W2[i], w2_hi[canon(i / K, i % K, N)], w2_lo[canon(i / K, i % K, N)]);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  // ---- GEMM 1
  if (tid == 0) issue_gemm(tmem, a_hi, a_lo, w1_hi, w1_lo, &mbar);
  mbar_wait(&mbar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // epilogue 1: h = LeakyReLU(acc + b1), written back in place as the next operand (hi / lo, canonical layout).
  // warp w owns rows 32w + lane of block 0; warp 3 also owns rows 128 + lane (block 1, TMEM lanes 96..127).
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1 && warp != 3) break;
    const int row = pass == 0 ? 32 * warp + lane : 128 + lane;
    const uint32_t lane_base = (uint32_t)(pass == 0 ? 32 * warp : 96) << 16;
    for (int c0 = 0; c0 < N; c0 += 8) {
      float v[8];
      tmem_ld8(tmem + lane_base + (uint32_t)(pass * 64 + c0), v);
      for (int j = 0; j < 8; ++j) {
        float h = __fadd_rn(v[j], B1[c0 + j]);
        h = h > 0.0f ? h : __fmul_rn(h, 0.3f);
        split(h, a_hi[canon(row, c0 + j, M)], a_lo[canon(row, c0 + j, M)]);
      }
    }
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // operand rewritten with ordinary stores -> visible to the MMA
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

  // ---- GEMM 2
  if (tid == 0) issue_gemm(tmem, a_hi, a_lo, w2_hi, w2_lo, &mbar);
  mbar_wait(&mbar, 1);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1 && warp != 3) break;
    const int row = pass == 0 ? 32 * warp + lane : 128 + lane;
    const uint32_t lane_base = (uint32_t)(pass == 0 ? 32 * warp : 96) << 16;
    for (int c0 = 0; c0 < N; c0 += 8) {
      float v[8];
      tmem_ld8(tmem + lane_base + (uint32_t)(pass * 64 + c0), v);
      for (int j = 0; j < 8; ++j) OUT[row * N + c0 + j] = __fadd_rn(__fadd_rn(v[j], B2[c0 + j]), U[row * N + c0 + j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;\n" ::"r"(tmem) : "memory");
}

int main() {
  std::vector<float> A(M * K), U(M * N), W1(N * K), W2(N * K), B1(N), B2(N), OUT(M * N);
  srand(7);
  auto rnd = [] { return (float)(rand() % 20001 - 10000) / 10000.0f * 1.37f; };
  for (auto& v : A) v = rnd();
  for (auto& v : U) v = rnd();
  for (auto& v : W1) v = rnd() * 0.25f;
  for (auto& v : W2) v = rnd() * 0.25f;
  for (auto& v : B1) v = rnd() * 0.1f;
  for (auto& v : B2) v = rnd() * 0.1f;
  std::vector<double> ref(M * N);
  double scale = 0;
  for (int m = 0; m < M; ++m) {
    double h[N];
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * (double)W1[n * K + k];
      s += B1[n];
      h[n] = (double)(float)(s > 0 ? s : s * (double)0.3f);      // the device rounds h to fp32 before the second GEMM
    }
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += h[k] * (double)W2[n * K + k];
      ref[m * N + n] = s + B2[n] + U[m * N + n];
      scale = std::fmax(scale, std::fabs(ref[m * N + n]));
    }
  }
  float *dA, *dU, *dW1, *dW2, *dB1, *dB2, *dO;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dU, U.size() * 4)); CK(cudaMalloc(&dW1, W1.size() * 4)); CK(cudaMalloc(&dW2, W2.size() * 4));
  CK(cudaMalloc(&dB1, N * 4)); CK(cudaMalloc(&dB2, N * 4)); CK(cudaMalloc(&dO, OUT.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dU, U.data(), U.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW1, W1.data(), W1.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dW2, W2.data(), W2.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB1, B1.data(), N * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB2, B2.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0, OUT.size() * 4));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  probe<<<1, 128, kSmem>>>(dA, dW1, dB1, dW2, dB2, dU, dO);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { std::printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  CK(cudaMemcpy(OUT.data(), dO, OUT.size() * 4, cudaMemcpyDeviceToHost));
  double worst = 0;
  int worst_row = -1;
  for (int i = 0; i < M * N; ++i) {
    const double d = std::fabs((double)OUT[i] - ref[i]);
    if (d > worst) { worst = d; worst_row = i / N; }
  }
  std::printf("max |OUT - ref| = %.3e (output scale %.2f, relative %.2e), worst row %d -> %s\n", worst, scale, worst / scale, worst_row,
              worst / scale < 5e-6 ? "MATCH at fp32 level" : "mismatch");
  return 0;
}
