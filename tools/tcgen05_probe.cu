// Dev probe (GPU, standalone): one CTA computes D[128 x 64] = A[128 x 64] * B[64 x 64]^T with tcgen05.mma kind::tf32,
// operands in shared memory in the canonical no-swizzle K-major layout, accumulator in TMEM, and checks it against the
// CPU.  Purpose: pin the shared-memory descriptor semantics (LBO / SBO) before moving the decoder's tensor-core mode
// from mma.sync to UMMA.   build: nvcc -gencode arch=compute_100a,code=sm_100a -o devtools_build/tcgen05_probe tools/tcgen05_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;            // descriptor version 1 (Blackwell)
  return d;                          // layout type 0 = no swizzle, base offset 0
}

// canonical K-major no-swizzle: core matrix = 8 rows x 16 bytes (4 tf32), stored [k/4][row/8][row%8][k%4]
__device__ __forceinline__ int canon(int row, int k, int rows) { return ((k / 4) * (rows / 8) + row / 8) * 32 + (row % 8) * 4 + k % 4; }

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int variant) {
  extern __shared__ __align__(128) float dyn[];
  float* sA = dyn;
  float* sB = dyn + M * K;
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  for (int i = tid; i < M * K; i += 128) sA[canon(i / K, i % K, M)] = A[i];
  for (int i = tid; i < N * K; i += 128) sB[canon(i / K, i % K, N)] = B[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic stores above -> visible to the tensor-core (async) proxy
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    // instruction descriptor: D = f32, A = B = tf32, both K-major, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t kchunkA = (M / 8) * 128, kchunkB = (N / 8) * 128;    // bytes between the k/4 groups of core matrices
    for (int ks = 0; ks < K / 8; ++ks) {
      uint64_t da, db;
      if (variant == 0) {      // LBO = distance between the two k-halves, SBO = distance between 8-row groups
        da = make_desc(smem_u32(sA) + ks * 2 * kchunkA, kchunkA, 128);
        db = make_desc(smem_u32(sB) + ks * 2 * kchunkB, kchunkB, 128);
      } else {                 // the two swapped
        da = make_desc(smem_u32(sA) + ks * 2 * kchunkA, 128, kchunkA);
        db = make_desc(smem_u32(sB) + ks * 2 * kchunkB, 128, kchunkB);
      }
      const uint32_t acc = ks > 0 ? 1u : 0u;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // everyone waits for the MMAs
  {
    const uint32_t bar = smem_u32(&mbar);
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DN;\nbra W;\nDN:\n}\n" ::"r"(bar) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // warp w reads TMEM lanes 32w .. 32w+31 (rows), 8 columns at a time
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(size_t)(32 * warp + lane) * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;\n" ::"r"(tmem) : "memory");
}

static float tf32(float x) { uint32_t b; memcpy(&b, &x, 4); b &= 0xffffe000u; memcpy(&x, &b, 4); return x; }

int main() {
  std::vector<float> A(M * K), B(N * K), D(M * N), ref(M * N);
  srand(1);
  for (auto& v : A) v = (float)(rand() % 2001 - 1000) / 1000.0f;
  for (auto& v : B) v = (float)(rand() % 2001 - 1000) / 1000.0f;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)tf32(A[m * K + k]) * (double)tf32(B[n * K + k]);
      ref[m * N + n] = (float)s;
    }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  for (int variant = 0; variant < 2; ++variant) {
    CK(cudaMemset(dD, 0, D.size() * 4));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (M + N) * K * 4));
    probe<<<1, 128, (M + N) * K * 4>>>(dA, dB, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { std::printf("variant %d: kernel failed: %s\n", variant, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (int i = 0; i < M * N; ++i) worst = std::fmax(worst, std::fabs((double)D[i] - (double)ref[i]));
    std::printf("variant %d (%s): max |D - ref| = %.3e  D[0]=%f ref[0]=%f D[last]=%f ref[last]=%f  -> %s\n", variant,
                variant == 0 ? "LBO = k-half stride, SBO = 8-row-group stride" : "swapped", worst, D[0], ref[0], D[M * N - 1], ref[M * N - 1],
                worst < 1e-3 ? "MATCH" : "mismatch");
  }
  return 0;
}
