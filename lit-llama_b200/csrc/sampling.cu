// The sampling tail of generate() (generate.py:68-75) up to the probabilities:
//   logits / temperature -> top-k threshold -> where(logits < thr, -inf, logits) -> softmax
// as ONE single-CTA kernel (the reference spends ~10 launches, including a radix sort for
// topk).  torch.multinomial stays in torch so the RNG stream is the reference's.
//
// Rounding points follow the reference as it runs on the GPU in bf16:
//   * logits / temperature is a bf16 tensor: ATen multiplies by the fp32 reciprocal of the
//     scalar and rounds to bf16;
//   * the threshold is the k-th largest of those bf16 values; ties at the threshold are kept
//     (`logits < thr` is false for them), exactly like torch.where;
//   * softmax is evaluated in fp32 (exp(x - max) / sum) and rounded to bf16.
#include "b2l_common.cuh"

namespace b2l {

constexpr int SAMP_THREADS = 1024;

__device__ __forceinline__ uint32_t bf16_key(uint16_t b) {  // monotone map: larger float -> larger key
  return (b & 0x8000u) ? (uint32_t)(uint16_t)~b : (uint32_t)(b | 0x8000u);
}

__global__ void __launch_bounds__(SAMP_THREADS)
    topk_softmax_kernel(const __nv_bfloat16* __restrict__ logits, float inv_temperature, int top_k,
                        __nv_bfloat16* __restrict__ probs, int V) {
  extern __shared__ __align__(16) uint8_t ssm[];
  uint16_t* sv = reinterpret_cast<uint16_t*>(ssm);  // scaled logits as bf16 bits [V]
  __shared__ int hist[256];
  __shared__ int sel_hi, sel_rank;
  __shared__ uint32_t kth_key;
  __shared__ float red[32];
  const int tid = threadIdx.x;

  // 1. scale (bf16 result) and histogram of the high byte of the sortable key
  if (tid < 256) hist[tid] = 0;
  __syncthreads();
  float lmax = -INFINITY;
  for (int i = tid; i < V; i += SAMP_THREADS) {
    const float s = rbf(bf2f(logits[i]) * inv_temperature);
    const __nv_bfloat16 sb = f2bf(s);
    const uint16_t bits = *reinterpret_cast<const uint16_t*>(&sb);
    sv[i] = bits;
    lmax = fmaxf(lmax, s);
    if (top_k > 0 && top_k < V) atomicAdd(&hist[bf16_key(bits) >> 8], 1);
  }
  lmax = warp_max(lmax);
  if ((tid & 31) == 0) red[tid >> 5] = lmax;
  __syncthreads();
  if (tid < 32) {
    float m = red[tid];
    m = warp_max(m);
    if (tid == 0) red[0] = m;
  }
  __syncthreads();
  const float gmax = red[0];

  uint32_t thr_key = 0;  // keep everything
  if (top_k > 0 && top_k < V) {
    // 2. bin of the k-th largest (scan from the top), then its rank inside the bin
    if (tid == 0) {
      int cum = 0, b = 255;
      for (; b >= 0; --b) {
        if (cum + hist[b] >= top_k) break;
        cum += hist[b];
      }
      sel_hi = b;
      sel_rank = top_k - cum;  // k-th largest is the sel_rank-th largest inside bin b
    }
    __syncthreads();
    const int hi = sel_hi;
    __syncthreads();
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < V; i += SAMP_THREADS) {
      const uint32_t key = bf16_key(sv[i]);
      if ((int)(key >> 8) == hi) atomicAdd(&hist[key & 0xFF], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int cum = 0, b = 255;
      for (; b >= 0; --b) {
        if (cum + hist[b] >= sel_rank) break;
        cum += hist[b];
      }
      kth_key = ((uint32_t)hi << 8) | (uint32_t)b;
    }
    __syncthreads();
    thr_key = kth_key;
  }

  // 3. softmax over the kept entries
  float sum = 0.f;
  for (int i = tid; i < V; i += SAMP_THREADS) {
    const uint16_t bits = sv[i];
    if (bf16_key(bits) >= thr_key) sum += expf(__uint_as_float((uint32_t)bits << 16) - gmax);
  }
  sum = warp_sum(sum);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  if (tid < 32) {
    float t = red[tid];
    t = warp_sum(t);
    if (tid == 0) red[0] = t;
  }
  __syncthreads();
  const float total = red[0];
  for (int i = tid; i < V; i += SAMP_THREADS) {
    const uint16_t bits = sv[i];
    const float pv = (bf16_key(bits) >= thr_key) ? expf(__uint_as_float((uint32_t)bits << 16) - gmax) / total : 0.f;
    probs[i] = f2bf(pv);
  }
}

}  // namespace b2l

using namespace b2l;

extern "C" int b2l_topk_softmax(const void* logits, float temperature, int top_k, void* probs, int V, b2l_stream_t stream) {
  B2L_CHECK_ARG(logits && probs && V > 0 && temperature > 0.f && top_k >= 0, "b2l_topk_softmax: bad argument");
  B2L_CHECK_SUPPORTED((size_t)V * 2 <= 200 * 1024, "b2l_topk_softmax: vocabulary %d too large for one CTA", V);
  const size_t smem = ((size_t)V * 2 + 15) & ~(size_t)15;
  static size_t configured = 0;
  if (smem > configured && smem > 48 * 1024) {
    B2L_CUDA(cudaFuncSetAttribute(topk_softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  topk_softmax_kernel<<<1, SAMP_THREADS, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)logits, 1.0f / temperature, top_k,
                                                                      (__nv_bfloat16*)probs, V);
  B2L_LAUNCH_CHECK("topk_softmax_kernel");
  return 0;
}
