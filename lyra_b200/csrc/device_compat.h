// Build-mode shim.  The product is compiled by nvcc for sm_100a.  The same kernel sources are also
// compiled by g++ with -DLYRA_EMU against tests/cuda_emu (a test-only CPU block simulator) so the
// CPU test tier can check kernel logic against the oracle without a GPU.  Nothing in the shipped
// library depends on the emulator.
#pragma once

#ifdef LYRA_EMU
#include "cuda_emu.h"
#define LYRA_DYN_SMEM() (reinterpret_cast<unsigned char*>(cuda_emu::g_blk->smem))
#define LYRA_LAUNCH(kernel, grid, block, smem, stream, ...) \
  cuda_emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define LYRA_SET_MAX_SMEM(kernel, bytes) (0)
#define LYRA_DEVICE_CODE 1
#define LYRA_TRAP() std::abort()
#else
#include <cuda_runtime.h>
#define LYRA_DYN_SMEM() (lyra_dyn_smem_raw)
#define LYRA_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define LYRA_SET_MAX_SMEM(kernel, bytes) \
  cudaFuncSetAttribute((kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#define LYRA_TRAP() __trap()
#ifdef __CUDACC__
extern __shared__ __align__(1024) unsigned char lyra_dyn_smem_raw[];
#endif
#endif

#include <stdint.h>

// ---- 16-byte asynchronous global->shared copies (LDGSTS); synchronous memcpy in the emulator ----
#if defined(LYRA_EMU)
static inline void lyra_cp_async16(void* smem_dst, const void* gmem_src) { std::memcpy(smem_dst, gmem_src, 16); }
static inline void lyra_cp_async_commit() {}
template <int N>
static inline void lyra_cp_async_wait() {}
#elif defined(__CUDACC__)
__device__ __forceinline__ void lyra_cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void lyra_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void lyra_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
#endif

// ---- mbarrier + bulk asynchronous copy (TMA, cp.async.bulk): the weight pipeline of the fp32 GEMMs.
//      One elected thread arms a "full" barrier with the byte count and issues one bulk copy per weight chunk; consumer
//      warps wait on its phase parity, and release the stage through an "empty" barrier (one arrival per warp).
//      lyra_mbar_wait(bar, parity) returns once the phase with that parity has completed.
struct alignas(8) LyraMbar { unsigned long long v; };
#if defined(LYRA_EMU)
struct LyraMbarEmu { uint16_t expected, arrived, phase, pad; };
static inline LyraMbarEmu* lyra_mbar_emu(LyraMbar* b) { return reinterpret_cast<LyraMbarEmu*>(b); }
static inline void lyra_mbar_init(LyraMbar* b, unsigned count) { LyraMbarEmu* e = lyra_mbar_emu(b); e->expected = (uint16_t)count; e->arrived = 0; e->phase = 0; e->pad = 0; }
static inline void lyra_mbar_fence_init() {}
static inline void lyra_fence_proxy_async() {}
static inline void lyra_mbar_arrive(LyraMbar* b) {
  LyraMbarEmu* e = lyra_mbar_emu(b);
  if (++e->arrived == e->expected) { e->arrived = 0; e->phase ^= 1; }
}
// one thread arriving for `n` participants at once (mbarrier.arrive with a count operand); the count must not overshoot the phase
static inline void lyra_mbar_arrive_n(LyraMbar* b, unsigned n) {
  LyraMbarEmu* e = lyra_mbar_emu(b);
  e->arrived = (uint16_t)(e->arrived + n);
  if (e->arrived > e->expected) { std::fprintf(stderr, "cuda_emu: mbarrier arrival count overshoots the phase\n"); std::abort(); }
  if (e->arrived == e->expected) { e->arrived = 0; e->phase ^= 1; }
}
// the emulated bulk copy is synchronous and fibers are cooperative, so arming + copying is one atomic step:
// the arrival is counted after the data has been written
static inline void lyra_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, LyraMbar* b) {
  std::memcpy(smem_dst, gmem_src, bytes);
  lyra_mbar_arrive(b);
}
static inline void lyra_mbar_wait(LyraMbar* b, unsigned parity) {
  while (lyra_mbar_emu(b)->phase == parity) cuda_emu::yield();
}
// several bulk copies completing ONE barrier phase: begin(total bytes), copy ..., end (the emulated copies are synchronous, so
// the single arrival is counted by end; on the GPU begin arms the barrier and end is a no-op)
static inline void lyra_bulk_multi_begin(LyraMbar*, unsigned) {}
static inline void lyra_bulk_multi_copy(void* smem_dst, const void* gmem_src, unsigned bytes, LyraMbar*) { std::memcpy(smem_dst, gmem_src, bytes); }
static inline void lyra_bulk_multi_end(LyraMbar* b) { lyra_mbar_arrive(b); }
// one function-local shared object per kernel (the emulator backs them all with the same per-block scratch area)
#define LYRA_STATIC_SMEM(type, name, count) \
  static_assert(sizeof(type) * (count) <= 448, "emulated static shared memory: 448 bytes for objects + 64 for named barriers"); \
  type* name = reinterpret_cast<type*>(cuda_emu::g_blk->static_smem)
#elif defined(__CUDACC__)
__device__ __forceinline__ unsigned lyra_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lyra_mbar_init(LyraMbar* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(lyra_smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void lyra_mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
// orders this thread's earlier generic-proxy accesses to shared memory before later asynchronous-proxy (bulk copy) ones
__device__ __forceinline__ void lyra_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void lyra_mbar_arrive(LyraMbar* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(lyra_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void lyra_mbar_arrive_n(LyraMbar* b, unsigned n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;\n" ::"r"(lyra_smem_u32(b)), "r"(n) : "memory");
}
// arm the barrier with the byte count and issue the bulk copy that completes it
__device__ __forceinline__ void lyra_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, LyraMbar* b) {
  const unsigned bar = lyra_smem_u32(b);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(lyra_smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void lyra_bulk_multi_begin(LyraMbar* b, unsigned total_bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(lyra_smem_u32(b)), "r"(total_bytes) : "memory");
}
__device__ __forceinline__ void lyra_bulk_multi_copy(void* smem_dst, const void* gmem_src, unsigned bytes, LyraMbar* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(lyra_smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(lyra_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void lyra_bulk_multi_end(LyraMbar*) {}
__device__ __forceinline__ void lyra_mbar_wait(LyraMbar* b, unsigned parity) {
  const unsigned bar = lyra_smem_u32(b);
  asm volatile("{\n"
               ".reg .pred p;\n"
               "LYRA_WAIT:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra LYRA_DONE;\n"
               "bra LYRA_WAIT;\n"
               "LYRA_DONE:\n"
               "}\n" ::"r"(bar), "r"(parity) : "memory");
}
#define LYRA_STATIC_SMEM(type, name, count) __shared__ type name##_storage[count]; type* name = name##_storage
#endif

// ---- warp-level int8 tensor-core MMA: D(16x8,s32) += A(16x32,s8,row) * B(32x8,s8,col), fragments as in the PTX ISA
//      (mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32).  Integer accumulation is exact, so results are
//      bit-identical to the dp4a / scalar formulation whatever the internal order.
#if defined(LYRA_EMU)
static inline void lyra_mma_s8_16x8x32(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  const uint32_t mine[6] = {a[0], a[1], a[2], a[3], b[0], b[1]};
  const uint32_t (*all)[8] = cuda_emu::warp_gather(mine, 6);
  const int lane = (int)(threadIdx.x & 31), g = lane >> 2, t = lane & 3;
  auto A = [&](int row, int k) {      // row 0..15, k 0..31
    const int src = (row & 7) * 4 + (k & 15) / 4;
    const int reg = (row >> 3) + 2 * (k >> 4);
    return (int)(int8_t)((all[src][reg] >> (8 * (k & 3))) & 0xff);
  };
  auto B = [&](int k, int col) {
    const int src = col * 4 + (k & 15) / 4;
    return (int)(int8_t)((all[src][4 + (k >> 4)] >> (8 * (k & 3))) & 0xff);
  };
  int d[4] = {c[0], c[1], c[2], c[3]};
  for (int k = 0; k < 32; ++k) {
    d[0] += A(g, k) * B(k, 2 * t);
    d[1] += A(g, k) * B(k, 2 * t + 1);
    d[2] += A(g + 8, k) * B(k, 2 * t);
    d[3] += A(g + 8, k) * B(k, 2 * t + 1);
  }
  cuda_emu::warp_barrier();             // keep the table alive until every lane has read it
  c[0] = d[0]; c[1] = d[1]; c[2] = d[2]; c[3] = d[3];
}
#elif defined(__CUDACC__)
__device__ __forceinline__ void lyra_mma_s8_16x8x32(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
#endif

// ---- warp-level TF32 tensor-core MMA: D(16x8,f32) += A(16x8,tf32,row) * B(8x8,tf32,col)
//      (mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32).  Used only by the decoder's opt-in tensor-core mode
//      (split-precision "3xTF32": fp32-equivalent accuracy, not bit-identical to the fmaf chain).
//      Fragments (g = lane / 4, t = lane % 4): a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4); b0 (k=t, n=g) b1 (k=t+4, n=g);
//      c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1).
#if defined(LYRA_EMU)
static inline float lyra_emu_tf32(uint32_t bits) {          // the tensor core reads the top 19 bits of each operand
  bits &= 0xffffe000u;
  float f;
  std::memcpy(&f, &bits, 4);
  return f;
}
static inline void lyra_mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  const uint32_t mine[6] = {a[0], a[1], a[2], a[3], b[0], b[1]};
  const uint32_t (*all)[8] = cuda_emu::warp_gather(mine, 6);
  const int lane = (int)(threadIdx.x & 31), g = lane >> 2, t = lane & 3;
  auto A = [&](int row, int k) { return (double)lyra_emu_tf32(all[(row & 7) * 4 + (k & 3)][(row >> 3) + 2 * (k >> 2)]); };
  auto B = [&](int k, int col) { return (double)lyra_emu_tf32(all[col * 4 + (k & 3)][4 + (k >> 2)]); };
  double d[4] = {c[0], c[1], c[2], c[3]};
  for (int k = 0; k < 8; ++k) {
    d[0] += A(g, k) * B(k, 2 * t);
    d[1] += A(g, k) * B(k, 2 * t + 1);
    d[2] += A(g + 8, k) * B(k, 2 * t);
    d[3] += A(g + 8, k) * B(k, 2 * t + 1);
  }
  cuda_emu::warp_barrier();
  c[0] = (float)d[0]; c[1] = (float)d[1]; c[2] = (float)d[2]; c[3] = (float)d[3];
}
#elif defined(__CUDACC__)
__device__ __forceinline__ void lyra_mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
#endif

// ---- 5th-generation tensor cores (tcgen05 / UMMA) with accumulators in tensor memory (TMEM).
//      Not used by the shipped kernels yet: these wrappers carry the instruction sequences that
//      tools/tcgen05_probe.cu verified on a B200, plus an emulator model so the CPU test tier can run UMMA kernels.
//      Shared-memory matrix descriptor (no swizzle, K-major): element (row, k) of an operand lives at
//        start + (k / 4) * LBO + (row / 8) * SBO + (row % 8) * 16 + (k % 4) * 4   bytes
//      i.e. 8-row x 16-byte core matrices, LBO = byte distance of the two 4-element k-halves of one MMA (K = 8),
//      SBO = byte distance of consecutive 8-row groups.  kind::tf32 reads the top 19 bits of every operand.
//      TMEM address = (lane << 16) | column; an M = 128 accumulator puts row r in lane r, column n in column n.
__host__ __device__ inline uint32_t lyra_umma_idesc_tf32(int M, int N) {
  // D = f32 (bits 4-5 = 1), A = B = tf32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
#if defined(LYRA_EMU)
static inline uint32_t lyra_emu_smem_addr(const void* p) { return (uint32_t)(reinterpret_cast<const char*>(p) - cuda_emu::g_blk->smem); }
static inline uint64_t lyra_umma_desc(const void* smem, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((lyra_emu_smem_addr(smem) >> 4) & 0x3FFF) | (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16 |
         (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32 | (uint64_t)1 << 46;
}
static inline float* lyra_emu_tmem(uint32_t taddr, int lane_add, int col_add) {
  cuda_emu::BlockState* b = cuda_emu::g_blk;
  if (b->tmem.empty()) b->tmem.assign(128 * 512, 0.0f);
  const unsigned lane = (taddr >> 16) + (unsigned)lane_add, col = (taddr & 0xffffu) + (unsigned)col_add;
  if (lane >= 128 || col >= 512) { std::fprintf(stderr, "cuda_emu: TMEM access out of range (lane %u, column %u)\n", lane, col); std::abort(); }
  return &b->tmem[(size_t)lane * 512 + col];
}
// warp-collective in hardware; here every lane of the calling warp runs it and lane 0 does the work
static inline void lyra_tmem_alloc(uint32_t* smem_slot, int ncols) {
  if ((threadIdx.x & 31) == 0) {
    cuda_emu::BlockState* b = cuda_emu::g_blk;
    if (ncols < 32 || (ncols & (ncols - 1)) || b->tmem_used + (unsigned)ncols > 512) { std::fprintf(stderr, "cuda_emu: bad TMEM allocation\n"); std::abort(); }
    *smem_slot = b->tmem_used;
    b->tmem_used += (unsigned)ncols;
  }
  __syncwarp();
}
static inline void lyra_tmem_dealloc(uint32_t, int) { __syncwarp(); }
static inline void lyra_tc_fence_before_sync() {}
static inline void lyra_tc_fence_after_sync() {}
// D[M x N] (+)= A[M x 8] * B[N x 8]^T, one thread issues it; the emulator executes it on the spot
static inline void lyra_umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  const int N = (int)((idesc >> 17) & 0x3f) << 3, M = (int)((idesc >> 24) & 0x1f) << 4;
  const char* base = cuda_emu::g_blk->smem;
  auto elem = [&](uint64_t d, int row, int k) {
    const uint32_t start = (uint32_t)(d & 0x3FFF) << 4, lbo = (uint32_t)((d >> 16) & 0x3FFF) << 4, sbo = (uint32_t)((d >> 32) & 0x3FFF) << 4;
    uint32_t bits;
    std::memcpy(&bits, base + start + (uint32_t)(k / 4) * lbo + (uint32_t)(row / 8) * sbo + (uint32_t)(row % 8) * 16 + (uint32_t)(k % 4) * 4, 4);
    return (double)lyra_emu_tf32(bits);
  };
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = accumulate ? (double)*lyra_emu_tmem(tmem_d, m, n) : 0.0;
      for (int k = 0; k < 8; ++k) acc += elem(desc_a, m, k) * elem(desc_b, n, k);
      *lyra_emu_tmem(tmem_d, m, n) = (float)acc;
    }
}
// all MMAs issued so far by this thread are complete when the barrier's phase completes
static inline void lyra_umma_commit(LyraMbar* b) { lyra_mbar_arrive(b); }
// 32x32b shape: lane i of the warp reads TMEM lane (address lane) + i, 8 consecutive columns
static inline void lyra_tmem_ld8(uint32_t taddr, float (&v)[8]) {
  const int lane = (int)(threadIdx.x & 31), warp = (int)(threadIdx.x >> 5);
  if ((taddr >> 16) != (uint32_t)(32 * (warp % 4))) {      // hardware rule: warp w of a warpgroup reaches TMEM lanes 32 (w % 4) .. + 31 only
    std::fprintf(stderr, "cuda_emu: warp %d may not read TMEM lanes starting at %u\n", warp, taddr >> 16);
    std::abort();
  }
  for (int j = 0; j < 8; ++j) v[j] = *lyra_emu_tmem(taddr, lane, j);
}

// ---- additions for the product's UMMA kernels: A operand in TMEM, kind::i8, tcgen05.st, wide tcgen05.ld ----
static inline uint32_t* lyra_emu_tmem_u32(uint32_t taddr, int lane_add, int col_add) { return reinterpret_cast<uint32_t*>(lyra_emu_tmem(taddr, lane_add, col_add)); }
static inline void lyra_emu_check_lane_window(uint32_t taddr) {
  const int warp = (int)(threadIdx.x >> 5);
  if ((taddr >> 16) != (uint32_t)(32 * (warp % 4))) {
    std::fprintf(stderr, "cuda_emu: warp %d may not access TMEM lanes starting at %u\n", warp, taddr >> 16);
    std::abort();
  }
}
// D[M x N] (+)= A[M x 8] * B[N x 8]^T with A in TMEM: row r of A is TMEM lane r, element k is column (address column) + k
static inline void lyra_umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  const int N = (int)((idesc >> 17) & 0x3f) << 3, M = (int)((idesc >> 24) & 0x1f) << 4;
  const char* base = cuda_emu::g_blk->smem;
  auto belem = [&](int row, int k) {
    const uint32_t start = (uint32_t)(desc_b & 0x3FFF) << 4, lbo = (uint32_t)((desc_b >> 16) & 0x3FFF) << 4, sbo = (uint32_t)((desc_b >> 32) & 0x3FFF) << 4;
    uint32_t bits;
    std::memcpy(&bits, base + start + (uint32_t)(k / 4) * lbo + (uint32_t)(row / 8) * sbo + (uint32_t)(row % 8) * 16 + (uint32_t)(k % 4) * 4, 4);
    return (double)lyra_emu_tf32(bits);
  };
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = accumulate ? (double)*lyra_emu_tmem(tmem_d, m, n) : 0.0;
      for (int k = 0; k < 8; ++k) acc += (double)lyra_emu_tf32(*lyra_emu_tmem_u32(tmem_a, m, k)) * belem(n, k);
      *lyra_emu_tmem(tmem_d, m, n) = (float)acc;
    }
}
// kind::i8: D (s32) [M x N] (+)= A (s8) [M x 32] * B (s8) [N x 32]^T; K-major no-swizzle operands are 8-row x 16-byte core
// matrices: element (row, k) at start + (k / 16) * LBO + (row / 8) * SBO + (row % 8) * 16 + k % 16 bytes
__host__ __device__ inline uint32_t lyra_umma_idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // D = s32, A = B = s8
}
static inline int lyra_emu_i8_elem(uint64_t d, int row, int k) {
  const char* base = cuda_emu::g_blk->smem;
  const uint32_t start = (uint32_t)(d & 0x3FFF) << 4, lbo = (uint32_t)((d >> 16) & 0x3FFF) << 4, sbo = (uint32_t)((d >> 32) & 0x3FFF) << 4;
  return (int)*reinterpret_cast<const int8_t*>(base + start + (uint32_t)(k / 16) * lbo + (uint32_t)(row / 8) * sbo + (uint32_t)(row % 8) * 16 + (uint32_t)(k % 16));
}
static inline void lyra_umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  const int N = (int)((idesc >> 17) & 0x3f) << 3, M = (int)((idesc >> 24) & 0x1f) << 4;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      int32_t* d = reinterpret_cast<int32_t*>(lyra_emu_tmem(tmem_d, m, n));
      int acc = accumulate ? *d : 0;
      for (int k = 0; k < 32; ++k) acc += lyra_emu_i8_elem(desc_a, m, k) * lyra_emu_i8_elem(desc_b, n, k);
      *d = acc;
    }
}
// A (s8) in TMEM: row r is lane r, 4 consecutive k per 32-bit column (little endian), 8 columns per MMA
static inline void lyra_umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  const int N = (int)((idesc >> 17) & 0x3f) << 3, M = (int)((idesc >> 24) & 0x1f) << 4;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      int32_t* d = reinterpret_cast<int32_t*>(lyra_emu_tmem(tmem_d, m, n));
      int acc = accumulate ? *d : 0;
      for (int k = 0; k < 32; ++k) acc += (int)(int8_t)((*lyra_emu_tmem_u32(tmem_a, m, k / 4) >> (8 * (k % 4))) & 0xff) * lyra_emu_i8_elem(desc_b, n, k);
      *d = acc;
    }
}
template <int NC>
static inline void lyra_tmem_ld(uint32_t taddr, uint32_t (&v)[NC]) {
  lyra_emu_check_lane_window(taddr);
  const int lane = (int)(threadIdx.x & 31);
  for (int j = 0; j < NC; ++j) v[j] = *lyra_emu_tmem_u32(taddr, lane, j);
}
template <int NC>
static inline void lyra_tmem_st(uint32_t taddr, const uint32_t (&v)[NC]) {
  lyra_emu_check_lane_window(taddr);
  const int lane = (int)(threadIdx.x & 31);
  for (int j = 0; j < NC; ++j) *lyra_emu_tmem_u32(taddr, lane, j) = v[j];
}
static inline void lyra_tmem_wait_ld() {}
static inline void lyra_tmem_wait_st() {}
#elif defined(__CUDACC__)
__device__ __forceinline__ uint64_t lyra_umma_desc(const void* smem, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((lyra_smem_u32(smem) >> 4) & 0x3FFF) | (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16 |
         (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32 | (uint64_t)1 << 46;     // descriptor version 1, no swizzle
}
// one whole warp; ncols a power of two >= 32; the TMEM base address is written to *smem_slot (shared memory)
__device__ __forceinline__ void lyra_tmem_alloc(uint32_t* smem_slot, int ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(lyra_smem_u32(smem_slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void lyra_tmem_dealloc(uint32_t taddr, int ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void lyra_tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void lyra_tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void lyra_umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate ? 1u : 0u) : "memory");
}
__device__ __forceinline__ void lyra_umma_commit(LyraMbar* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(lyra_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void lyra_tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}

__host__ __device__ inline uint32_t lyra_umma_idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // D = s32, A = B = s8
}
__device__ __forceinline__ void lyra_umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate ? 1u : 0u) : "memory");
}
__device__ __forceinline__ void lyra_umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate ? 1u : 0u) : "memory");
}
__device__ __forceinline__ void lyra_umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate ? 1u : 0u) : "memory");
}
// 32x32b shape, NC consecutive columns: lane i of the warp accesses TMEM lane (address lane) + i.  The loaded registers are valid
// after lyra_tmem_wait_ld(); stored data is visible to later MMAs after lyra_tmem_wait_st() (+ the usual fences).
template <int NC> __device__ __forceinline__ void lyra_tmem_ld(uint32_t taddr, uint32_t (&v)[NC]);
template <> __device__ __forceinline__ void lyra_tmem_ld<8>(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void lyra_tmem_ld<16>(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}
template <int NC> __device__ __forceinline__ void lyra_tmem_st(uint32_t taddr, const uint32_t (&v)[NC]);
template <> __device__ __forceinline__ void lyra_tmem_st<8>(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
template <> __device__ __forceinline__ void lyra_tmem_st<16>(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void lyra_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void lyra_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
#endif

// ---- bulk asynchronous store shared -> global (TMA, cp.async.bulk.global.shared::cta) and named barriers for warp subsets.
//      Stores are grouped with lyra_bulk_commit(); lyra_bulk_wait_read() returns once the shared-memory source of every committed
//      group may be overwritten, lyra_bulk_wait_all() once the writes themselves are complete.  One thread issues and waits.
#if defined(LYRA_EMU)
static inline void lyra_bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) { std::memcpy(gmem_dst, smem_src, bytes); }
static inline void lyra_prefetch_l2(const void*, unsigned) {}
static inline void lyra_bulk_commit() {}
static inline void lyra_bulk_wait_read() {}
static inline void lyra_bulk_wait_all() {}
// barrier `id` (1..15) over `nthreads` threads (a multiple of 32): every participating thread calls it
static inline void lyra_named_bar_sync(int id, int nthreads) {
  cuda_emu::BlockState* b = cuda_emu::g_blk;
  unsigned* st = reinterpret_cast<unsigned*>(b->static_smem + 448) + 2 * (id & 7);     // {arrived, generation}; ids are used modulo 8 here
  const unsigned gen = st[1];
  if (++st[0] == (unsigned)nthreads) { st[0] = 0; st[1] = gen + 1; }
  while (st[1] == gen) cuda_emu::yield();
}
#elif defined(__CUDACC__)
// asks the L2 to fetch `bytes` (multiple of 16) starting at the 16-byte aligned global address: a hint, no completion to wait for
__device__ __forceinline__ void lyra_prefetch_l2(const void* gmem, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lyra_bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(lyra_smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lyra_bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void lyra_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void lyra_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void lyra_named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory"); }
#endif

// ---- thread-block clusters: a pair of CTAs shares one weight stream (each CTA's producer loads half of every chunk and TMA
//      multicasts it into both CTAs' shared memory; a stage is recycled when BOTH CTAs' MMAs have released it).  The emulator
//      runs one block at a time: clusters do not exist there and kernels take their single-CTA path.
#if defined(LYRA_EMU)
static inline unsigned lyra_cluster_ctarank() { return 0; }
static inline unsigned lyra_cluster_nctarank() { return 1; }
static inline void lyra_cluster_sync() {}
static inline void lyra_bulk_g2s_mc(void* smem_dst, const void* gmem_src, unsigned bytes, LyraMbar*, unsigned) { std::memcpy(smem_dst, gmem_src, bytes); }
static inline void lyra_umma_commit_mc(LyraMbar* b, unsigned) { lyra_mbar_arrive(b); }
#elif defined(__CUDACC__)
__device__ __forceinline__ unsigned lyra_cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned lyra_cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ void lyra_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// bulk copy global -> the same shared-memory offset of every CTA in `cta_mask`; completes `bytes` of transaction count on the
// mbarrier at the same offset in each of them (the barriers are armed by their own CTAs, see lyra_bulk_multi_begin)
__device__ __forceinline__ void lyra_bulk_g2s_mc(void* smem_dst, const void* gmem_src, unsigned bytes, LyraMbar* b, unsigned cta_mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n"
               ::"r"(lyra_smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(lyra_smem_u32(b)), "h"((unsigned short)cta_mask) : "memory");
}
// tcgen05.commit that arrives on the mbarrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void lyra_umma_commit_mc(LyraMbar* b, unsigned cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
               ::"r"(lyra_smem_u32(b)), "h"((unsigned short)cta_mask) : "memory");
}
#endif

