// model.py element-wise pieces as stand-alone kernels (module-level drop-ins, prefill).
// In the decode step these are fused into the tcgen05 linear's prologue/epilogue.
#include "b2l_common.cuh"

namespace b2l {

// RMSNorm.forward, model.py:270-277, bf16 rounding points preserved.  One CTA per row.
__global__ void __launch_bounds__(256) rmsnorm_kernel(const __nv_bfloat16* __restrict__ x,
                                                      const __nv_bfloat16* __restrict__ scale,
                                                      __nv_bfloat16* __restrict__ y, int C, float eps) {
  __shared__ float red[32];
  const __nv_bfloat16* xr = x + (size_t)blockIdx.x * C;
  __nv_bfloat16* yr = y + (size_t)blockIdx.x * C;
  float ss = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    float v = bf2f(xr[i]);
    ss += rbf(v * v);
  }
  ss = block_sum(ss, red);
  const float rinv = rms_rinv(ss, C, eps);
  for (int i = threadIdx.x; i < C; i += blockDim.x) yr[i] = f2bf(rms_apply(bf2f(xr[i]), rinv, bf2f(scale[i])));
}

template <typename IdxT>
__global__ void embedding_kernel(const IdxT* __restrict__ idx, const __nv_bfloat16* __restrict__ wte,
                                 __nv_bfloat16* __restrict__ out, int C, int vocab) {
  pdl_launch_dependents();  // the first c_attn of the step may start streaming its weights
  long long t = (long long)idx[blockIdx.x];
  if (t < 0 || t >= vocab) t = 0;  // torch would raise; keep the kernel memory-safe
  const __nv_bfloat16* src = wte + (size_t)t * C;
  __nv_bfloat16* dst = out + (size_t)blockIdx.x * C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) dst[i] = src[i];
}

// silu(a) * b, model.py:252: silu rounds to bf16, then the product rounds to bf16.
__global__ void silu_mul_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                __nv_bfloat16* __restrict__ y, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float av = bf2f(a[i]);
    float s = rbf(av / (1.0f + expf(-av)));
    y[i] = f2bf(s * bf2f(b[i]));
  }
}

__global__ void add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                           __nv_bfloat16* __restrict__ y, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = f2bf(bf2f(a[i]) + bf2f(b[i]));
}

}  // namespace b2l

using namespace b2l;

extern "C" int b2l_rmsnorm(const void* x, const void* scale, void* y, int rows, int C, float eps,
                           b2l_stream_t stream) {
  B2L_CHECK_ARG(x && scale && y && rows >= 0 && C > 0, "b2l_rmsnorm: bad argument");
  if (rows == 0) return 0;
  rmsnorm_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)scale,
                                                         (__nv_bfloat16*)y, C, eps);
  B2L_LAUNCH_CHECK("rmsnorm_kernel");
  return 0;
}

extern "C" int b2l_embedding(const void* idx, int idx_is_i64, const void* wte, void* out, int n, int C, int vocab,
                             b2l_stream_t stream) {
  B2L_CHECK_ARG(idx && wte && out && n >= 0 && C > 0 && vocab > 0, "b2l_embedding: bad argument");
  if (n == 0) return 0;
  if (idx_is_i64)
    embedding_kernel<long long><<<n, 256, 0, (cudaStream_t)stream>>>((const long long*)idx, (const __nv_bfloat16*)wte, (__nv_bfloat16*)out, C, vocab);
  else
    embedding_kernel<int><<<n, 256, 0, (cudaStream_t)stream>>>((const int*)idx, (const __nv_bfloat16*)wte, (__nv_bfloat16*)out, C, vocab);
  B2L_LAUNCH_CHECK("embedding_kernel");
  return 0;
}

extern "C" int b2l_silu_mul(const void* a, const void* b, void* y, size_t n, b2l_stream_t stream) {
  B2L_CHECK_ARG(a && b && y, "b2l_silu_mul: null pointer");
  if (n == 0) return 0;
  silu_mul_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)y, n);
  B2L_LAUNCH_CHECK("silu_mul_kernel");
  return 0;
}

extern "C" int b2l_add(const void* a, const void* b, void* y, size_t n, b2l_stream_t stream) {
  B2L_CHECK_ARG(a && b && y, "b2l_add: null pointer");
  if (n == 0) return 0;
  add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)y, n);
  B2L_LAUNCH_CHECK("add_kernel");
  return 0;
}
