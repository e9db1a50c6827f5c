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
#else
#include <cuda_runtime.h>
#define LYRA_DYN_SMEM() (lyra_dyn_smem_raw)
#define LYRA_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define LYRA_SET_MAX_SMEM(kernel, bytes) \
  cudaFuncSetAttribute((kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
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
