// Shared device/host helpers for libb200llama (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/b2l.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb200llama is written for sm_100a (B200) only"
#endif

namespace b2l {

// ---- error state (thread-local message, returned through b2l_last_error) ----
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);  // records + returns (int)e

#define B2L_CHECK_ARG(cond, ...)                  \
  do {                                            \
    if (!(cond)) {                                \
      b2l::set_error(__VA_ARGS__);                \
      return B2L_E_ARG;                           \
    }                                             \
  } while (0)
#define B2L_CHECK_SUPPORTED(cond, ...)            \
  do {                                            \
    if (!(cond)) {                                \
      b2l::set_error(__VA_ARGS__);                \
      return B2L_E_UNSUPPORTED;                   \
    }                                             \
  } while (0)
#define B2L_CUDA(call)                                          \
  do {                                                          \
    cudaError_t e_ = (call);                                    \
    if (e_ != cudaSuccess) return b2l::cuda_fail(e_, #call);    \
  } while (0)
#define B2L_LAUNCH_CHECK(name)                                       \
  do {                                                               \
    cudaError_t e_ = cudaGetLastError();                             \
    if (e_ != cudaSuccess) return b2l::cuda_fail(e_, "launch " name); \
  } while (0)

int sm_count();  // of the current device, cached per device

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting: remember per device what a kernel was given.
constexpr int B2L_MAX_DEVICES = 64;
struct DynSmemCache {
  size_t bytes[B2L_MAX_DEVICES] = {};
};
template <typename Kernel>
inline int ensure_dyn_smem(Kernel kernel, size_t bytes, DynSmemCache& cache) {
  int dev = 0;
  B2L_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= B2L_MAX_DEVICES) {
    B2L_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
  }
  if (bytes > cache.bytes[dev]) {
    B2L_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    cache.bytes[dev] = bytes;
  }
  return 0;
}

// ---- small device helpers ----
__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ __nv_bfloat16 f2bf(float v) { return __float2bfloat16_rn(v); }
// round a float through bf16 (the reference keeps every intermediate in bf16)
__device__ __forceinline__ float rbf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ float load_sz(const void* p, int dtype, size_t i) {
  return dtype == B2L_BF16 ? bf2f(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                           : reinterpret_cast<const float*>(p)[i];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); `red` is >= 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from a previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  return warp_sum(t);
}

// Debug timeline (tools/diag.py `timeline`): per-launch uint64[8] of %globaltimer nanoseconds,
// slot 0 = min over CTAs (start), slots 1.. = max over CTAs.  nullptr = off.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tl_min(unsigned long long* tr, int slot) {
  if (tr != nullptr) atomicMin(tr + slot, globaltimer_ns());
}
__device__ __forceinline__ void tl_max(unsigned long long* tr, int slot) {
  if (tr != nullptr) atomicMax(tr + slot, globaltimer_ns());
}

// Programmatic dependent launch (PDL) device side.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// Coherent 16-byte global load (ld.global, never ld.global.nc).  Mandatory for data written by the PREVIOUS kernel
// when this kernel runs under programmatic dependent launch: griddepcontrol.wait orders coherent loads only, and a
// `const __restrict__` pointer lets the compiler pick the non-coherent path (LDG.E.CONSTANT).
__device__ __forceinline__ uint4 ld_coherent_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// RMSNorm with the reference's bf16 rounding points (model.py:270-277, no upcast):
//   ms = bf16(mean(bf16(x*x)));  r = bf16(rsqrt(bf16(ms + eps)));  y = bf16(scale * bf16(x * r))
// `sumsq` is the fp32 sum over the row of bf16-rounded squares.
__device__ __forceinline__ float rms_rinv(float sumsq, int C, float eps) {
  float ms = rbf(sumsq / (float)C);
  float t = rbf(ms + eps);
  return rbf(1.0f / sqrtf(t));
}
__device__ __forceinline__ float rms_apply(float x, float rinv, float scale) {
  return rbf(scale * rbf(x * rinv));
}

// Launch helper: optional PDL attribute and cluster dimension.
struct LaunchCfg {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attrs[2];
  LaunchCfg(dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, int cluster_x = 1) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    int n = 0;
    if (pdl) {
      attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attrs[n].val.programmaticStreamSerializationAllowed = 1;
      ++n;
    }
    if (cluster_x > 1) {
      attrs[n].id = cudaLaunchAttributeClusterDimension;
      attrs[n].val.clusterDim.x = cluster_x;
      attrs[n].val.clusterDim.y = 1;
      attrs[n].val.clusterDim.z = 1;
      ++n;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = n;
  }
};

}  // namespace b2l
