// Shared helpers for libv2ce_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/v2ce_b200.h"

namespace v2ce {

// thread-local message of the last failing call (v2ce_last_error)
char* last_error_buffer();
int set_error(int code, const char* fmt, ...);

#define V2CE_CUDA_CHECK(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::v2ce::set_error(V2CE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                             \
  } while (0)

#define V2CE_LAUNCH_CHECK(name)                                                                 \
  do {                                                                                          \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess)                                                                      \
      return ::v2ce::set_error(V2CE_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

#define V2CE_REQUIRE(cond, ...)                                                                 \
  do {                                                                                          \
    if (!(cond)) return ::v2ce::set_error(V2CE_ERR_INVALID, __VA_ARGS__);                       \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct Arena {
  char* base;
  size_t size;
  size_t off;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), size(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
  bool ok() const { return off <= size; }
};

int sm_count_cached();
// Programmatic dependent launch: a conv kernel is launched while its predecessor on the stream is still draining;
// its CTAs become resident as the predecessor's CTAs exit and run their prologue (mbarrier init, TMEM allocation,
// tensor-map fetch) before `griddepcontrol.wait` holds them until the predecessor's memory is visible.
// Opt-in with V2CE_PDL=1 (measured on B200: no gain for this network, the launch gaps are not what the forward waits
// for); without the launch attribute the device instructions are no-ops.
bool pdl_enabled();
// index (< 64) of the current CUDA device: function attributes and constant memory are per device, so the
// one-time-setup caches of the launchers are kept per device (one process may drive several GPUs)
int device_slot();

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace v2ce
