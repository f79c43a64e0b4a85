// Thin device-runtime shim: CUDA runtime in the product build, malloc/loops in the tests-only host
// emulation build (-DIMA_HOSTEMU, see ima_platform.h).  No CPU fallback exists in the product: without
// IMA_HOSTEMU every entry point goes through the CUDA runtime and fails loudly when there is no device.
#pragma once
#include "ima_platform.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>

#if IMA_CUDA
#include <cuda_runtime.h>
namespace ima {
typedef cudaStream_t stream_t;
#define IMA_KERNEL __global__
#define IMA_CONSTANT __constant__
inline int dev_check(cudaError_t e, const char *what, const char *file, int line) {
  if (e != cudaSuccess) {
    fprintf(stderr, "ima2p_b200: CUDA error %d (%s) at %s:%d: %s\n", (int)e, cudaGetErrorString(e), file, line, what);
    return 1;
  }
  return 0;
}
#define IMA_CUDA_OK(x) (::ima::dev_check((x), #x, __FILE__, __LINE__) == 0)
inline void *dev_alloc(size_t bytes) {
  void *p = nullptr;
  if (!IMA_CUDA_OK(cudaMalloc(&p, bytes ? bytes : 1))) return nullptr;
  // cudaMemset runs on the legacy default stream, asynchronously to the host and NOT ordered with the engine's
  // non-blocking streams: wait for it, or it could zero a buffer after a later upload/kernel has written it
  if (!IMA_CUDA_OK(cudaMemset(p, 0, bytes ? bytes : 1)) || !IMA_CUDA_OK(cudaStreamSynchronize(cudaStreamLegacy))) { cudaFree(p); return nullptr; }
  return p;
}
inline void dev_free(void *p) { if (p) cudaFree(p); }
inline bool h2d(void *d, const void *h, size_t n, stream_t s) { return IMA_CUDA_OK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); }
inline bool d2h(void *h, const void *d, size_t n, stream_t s) { return IMA_CUDA_OK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); }
inline bool dev_sync(stream_t s) { return IMA_CUDA_OK(cudaStreamSynchronize(s)); }
IMA_DEV int ima_block() { return blockIdx.x; }
IMA_DEV int ima_warp_in_block() { return threadIdx.x >> 5; }
// Every kernel opens with this line.  griddepcontrol.wait returns at once for a kernel launched the plain way; for one
// launched with programmatic stream serialisation (the step graph, see launch_kernel) it is where the kernel waits for the
// kernel before it on the stream to have finished and flushed its writes -- everything above it (block scheduling, parameter
// loads) overlaps that kernel's tail.
#define IMA_SMEM_DECL extern __shared__ __align__(16) unsigned char ima_dyn_smem[]; asm volatile("griddepcontrol.wait;" ::: "memory");
#define IMA_SMEM ima_dyn_smem
extern bool g_programmatic_launch;       // set by capture_steps around the launches of a step graph (ima_engine.cu)
template <class K, class... A> inline void launch_kernel(K kern, int grid, int threads, size_t smem, cudaStream_t s, A... args) {
  if (!g_programmatic_launch) { kern<<<grid, threads, smem, s>>>(args...); return; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, args...);
}
#define IMA_LAUNCH(kern, grid, warps, smem_bytes_, stream, ...) ::ima::launch_kernel(kern, (int)(grid), (int)(warps) * 32, (size_t)(smem_bytes_), (stream), __VA_ARGS__)
}  // namespace ima
#else
namespace ima {
typedef void *stream_t;
#define IMA_KERNEL
#define IMA_CONSTANT
inline void *dev_alloc(size_t bytes) { return calloc(bytes ? bytes : 1, 1); }
inline void dev_free(void *p) { free(p); }
inline bool h2d(void *d, const void *h, size_t n, stream_t) { memcpy(d, h, n); return true; }
inline bool d2h(void *h, const void *d, size_t n, stream_t) { memcpy(h, d, n); return true; }
inline bool dev_sync(stream_t) { return true; }
struct EmuCtx { int block, warp; unsigned char *smem; };
extern thread_local EmuCtx g_emu;
inline int ima_block() { return g_emu.block; }
inline int ima_warp_in_block() { return g_emu.warp; }
#define IMA_SMEM_DECL
#define IMA_SMEM (::ima::g_emu.smem)
#define IMA_LAUNCH(kern, grid, warps, smem_bytes_, stream, ...)                            \
  do {                                                                                     \
    unsigned char *emu_buf_ = (unsigned char *)malloc((smem_bytes_) + 64);                      \
    for (int b_ = 0; b_ < (int)(grid); b_++)                                               \
      for (int w_ = 0; w_ < (int)(warps); w_++) {                                          \
        ::ima::g_emu.block = b_; ::ima::g_emu.warp = w_; ::ima::g_emu.smem = emu_buf_;     \
        kern(__VA_ARGS__);                                                                 \
      }                                                                                    \
    free(emu_buf_);                                                                        \
  } while (0)
}  // namespace ima
#endif
