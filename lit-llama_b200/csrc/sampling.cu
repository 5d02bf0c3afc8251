// The sampling tail of generate() (generate.py:68-75) up to the probabilities:
//   logits / temperature -> top-k threshold -> where(logits < thr, -inf, logits) -> softmax
// as ONE single-CTA kernel (the reference spends ~10 launches, including a radix sort for
// topk), optionally followed in the same kernel by the draw itself: torch.multinomial(probs, 1) is
// argmax(probs / q) with q ~ Exp(1) (ATen/native/Distributions.cpp); the caller draws q with torch
// (`empty_like(probs).exponential_(1)`, the same RNG consumption as multinomial) so the sampled
// token equals the reference's for the same generator state, without multinomial's ~12 launches.
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

// warp-aggregated shared-memory histogram increment: lanes that hit the same bin elect one to add their count
// (after scaling, most logits share a few exponent bins -- plain atomics would serialise 32-way)
__device__ __forceinline__ void hist_add(int* hist, uint32_t bin, bool valid) {
  const uint32_t key = valid ? bin : 0xFFFFFFFFu;
  const uint32_t peers = __match_any_sync(0xffffffffu, key);
  if (valid && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
}

// warp 0: which of the 256 bins holds the `want`-th largest element (counting from bin 255 down), and the
// rank of that element inside the bin
__device__ __forceinline__ void find_bin(const int* hist, int want, int* sel_bin, int* sel_rank) {
  const int lane = threadIdx.x & 31;
  int c[8], s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j] = hist[lane * 8 + j]; s += c[j]; }
  int incl = s;  // sum over lanes >= lane
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += t;
  }
  const int above = incl - s;
  if (above < want && want <= above + s) {
    int cum = above;
#pragma unroll
    for (int j = 7; j >= 0; --j) {
      if (cum + c[j] >= want) { *sel_bin = lane * 8 + j; *sel_rank = want - cum; break; }
      cum += c[j];
    }
  }
}

__device__ __forceinline__ float bits_f(uint32_t bits16) { return __uint_as_float(bits16 << 16); }

__global__ void __launch_bounds__(SAMP_THREADS)
    topk_softmax_kernel(const __nv_bfloat16* __restrict__ logits, float inv_temperature, int top_k,
                        __nv_bfloat16* __restrict__ probs, const __nv_bfloat16* __restrict__ noise,
                        long long* __restrict__ token, int V) {
  extern __shared__ __align__(16) uint8_t ssm[];
  const int Vp = (V + 7) & ~7;
  uint16_t* sv = reinterpret_cast<uint16_t*>(ssm);  // scaled logits as bf16 bits [Vp]
  uint16_t* sq = sv + Vp;                            // Exp(1) noise as bf16 bits [Vp] (only with `noise`)
  __shared__ int hist[256];
  __shared__ int sel_hi, sel_rank, sel_lo;
  __shared__ float red[32];
  __shared__ int red_i[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool select = top_k > 0 && top_k < V;
  const int nvec = V / 8;  // 16-byte vectors; the (V % 8) tail is handled element-wise by the first threads
  constexpr int PRE = 4;   // rounds whose global loads are issued up front (covers V <= 32768)

  // 0. every global load of the first PRE rounds is in flight before anything waits
  uint4 lpre[PRE], npre[PRE];
#pragma unroll
  for (int r = 0; r < PRE; ++r) {
    const int i = r * SAMP_THREADS + tid;
    lpre[r] = make_uint4(0, 0, 0, 0);
    npre[r] = make_uint4(0, 0, 0, 0);
    if (i < nvec) {
      lpre[r] = reinterpret_cast<const uint4*>(logits)[i];
      if (noise != nullptr) npre[r] = reinterpret_cast<const uint4*>(noise)[i];
    }
  }

  // 1. scale (bf16 result) and histogram of the high byte of the sortable key
  if (tid < 256) hist[tid] = 0;
  __syncthreads();
  float lmax = -INFINITY;
  const int rounds = (nvec + SAMP_THREADS - 1) / SAMP_THREADS;
#pragma unroll 1
  for (int r0 = 0; r0 < rounds; r0 += PRE) {
#pragma unroll
    for (int rr = 0; rr < PRE; ++rr) {
      const int r = r0 + rr;
      if (r >= rounds) break;   // block-uniform
      const int i = r * SAMP_THREADS + tid;
      const bool valid = i < nvec;
      uint4 v = lpre[rr];
      if (r0 > 0) { v = make_uint4(0, 0, 0, 0); if (valid) v = reinterpret_cast<const uint4*>(logits)[i]; }
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      uint32_t o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float a = rbf(__uint_as_float(w[q] << 16) * inv_temperature);
        const float b = rbf(__uint_as_float(w[q] & 0xffff0000u) * inv_temperature);
        const uint32_t ab = __float_as_uint(a) >> 16, bb = __float_as_uint(b) >> 16;
        o[q] = ab | (bb << 16);
        if (valid) lmax = fmaxf(lmax, fmaxf(a, b));
        if (select) {
          hist_add(hist, bf16_key((uint16_t)ab) >> 8, valid);
          hist_add(hist, bf16_key((uint16_t)bb) >> 8, valid);
        }
      }
      if (valid) {
        reinterpret_cast<uint4*>(sv)[i] = make_uint4(o[0], o[1], o[2], o[3]);
        if (noise != nullptr) {
          uint4 n = npre[rr];
          if (r0 > 0) n = reinterpret_cast<const uint4*>(noise)[i];
          reinterpret_cast<uint4*>(sq)[i] = n;
        }
      }
    }
  }
  {
    const int i = nvec * 8 + tid;   // tail (fewer than 8 elements) and zero padding up to Vp
    const bool valid = i < V;
    uint16_t bits = 0xff80;         // padding: -inf, below every threshold
    if (valid) {
      const float sc = rbf(bf2f(logits[i]) * inv_temperature);
      bits = (uint16_t)(__float_as_uint(sc) >> 16);
      lmax = fmaxf(lmax, sc);
    }
    if (i < Vp) {
      sv[i] = bits;
      if (noise != nullptr) sq[i] = valid ? *reinterpret_cast<const uint16_t*>(noise + i) : (uint16_t)0x3f80;
    }
    if (select) hist_add(hist, bf16_key(bits) >> 8, valid);
  }
  lmax = warp_max(lmax);
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  float gmax = red[lane];
  gmax = warp_max(gmax);   // every warp reduces the 32 partials itself

  const int nvp = Vp / 8;
  uint32_t thr_key = 0;  // keep everything
  if (select) {
    // 2. bin of the k-th largest (from the top), then its rank inside the bin
    if (warp == 0) find_bin(hist, top_k, &sel_hi, &sel_rank);
    __syncthreads();
    const int hi = sel_hi, rank = sel_rank;
    __syncthreads();
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int base = 0; base < nvp; base += SAMP_THREADS) {
      const int i = base + tid;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (i < nvp) v = reinterpret_cast<const uint4*>(sv)[i];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const uint32_t key = bf16_key((uint16_t)(hf ? (w[q] >> 16) : (w[q] & 0xffffu)));
          const bool hit = i < nvp && (int)(key >> 8) == hi && (8 * i + 2 * q + hf) < V;
          if (__any_sync(0xffffffffu, hit)) hist_add(hist, key & 0xFF, hit);
        }
      }
    }
    __syncthreads();
    if (warp == 0) { int dummy; find_bin(hist, rank, &sel_lo, &dummy); }
    __syncthreads();
    thr_key = ((uint32_t)hi << 8) | (uint32_t)sel_lo;
  }

  // 3. softmax over the kept entries: e = exp(l - max) is kept in registers for the first PRE rounds
  float sum = 0.f;
  for (int base = 0; base < nvp; base += SAMP_THREADS) {
    const int i = base + tid;
    if (i < nvp) {
      const uint4 v = reinterpret_cast<const uint4*>(sv)[i];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      bool any = false;  // with top-k, 99 % of the vectors hold no kept entry: skip their exponentials
#pragma unroll
      for (int q = 0; q < 4; ++q)
        any = any || bf16_key((uint16_t)(w[q] & 0xffffu)) >= thr_key || bf16_key((uint16_t)(w[q] >> 16)) >= thr_key;
      if (any) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t lo = w[q] & 0xffffu, hi2 = w[q] >> 16;
          if (bf16_key((uint16_t)lo) >= thr_key && lo != 0xff80u) sum += expf(bits_f(lo) - gmax);
          if (bf16_key((uint16_t)hi2) >= thr_key && hi2 != 0xff80u) sum += expf(bits_f(hi2) - gmax);
        }
      }
    }
  }
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  float total = red[lane];
  total = warp_sum(total);

  // 4. probabilities (bf16) and, with `noise` (q ~ Exp(1) drawn by torch), the sample argmax(p / q) --
  // torch.multinomial's own algorithm for one draw (ATen/native/Distributions.cpp), ties to the lower index
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  for (int base = 0; base < nvp; base += SAMP_THREADS) {
    const int i = base + tid;
    if (i < nvp) {
      const uint4 v = reinterpret_cast<const uint4*>(sv)[i];
      uint4 n = make_uint4(0, 0, 0, 0);
      if (noise != nullptr) n = reinterpret_cast<const uint4*>(sq)[i];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      const uint32_t nw[4] = {n.x, n.y, n.z, n.w};
      uint32_t o[4] = {0u, 0u, 0u, 0u};
      bool any = false;  // a vector without kept entries has probability 0 everywhere and cannot win the argmax
#pragma unroll
      for (int q = 0; q < 4; ++q)
        any = any || bf16_key((uint16_t)(w[q] & 0xffffu)) >= thr_key || bf16_key((uint16_t)(w[q] >> 16)) >= thr_key;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (!any) break;
        uint32_t pb2[2];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const uint32_t bits = hf ? (w[q] >> 16) : (w[q] & 0xffffu);
          const bool kept = bf16_key((uint16_t)bits) >= thr_key && bits != 0xff80u;
          const float pv = kept ? expf(bits_f(bits) - gmax) / total : 0.f;
          const float pr = rbf(pv);
          pb2[hf] = __float_as_uint(pr) >> 16;
          if (noise != nullptr) {
            const int idx = 8 * i + 2 * q + hf;
            const float r = rbf(pr / bits_f(hf ? (nw[q] >> 16) : (nw[q] & 0xffffu)));
            if (idx < V && (r > best || (r == best && idx < best_i))) { best = r; best_i = idx; }
          }
        }
        o[q] = pb2[0] | (pb2[1] << 16);
      }
      if (probs != nullptr) {
        if (8 * i + 8 <= V && (reinterpret_cast<uintptr_t>(probs) & 15) == 0) {
          reinterpret_cast<uint4*>(probs)[i] = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (8 * i + e < V) reinterpret_cast<uint16_t*>(probs)[8 * i + e] = (uint16_t)(o[e >> 1] >> (16 * (e & 1)));
        }
      }
    }
  }
  if (noise != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
    __syncthreads();
    if (lane == 0) { red[warp] = best; red_i[warp] = best_i; }
    __syncthreads();
    if (warp == 0) {
      best = red[lane]; best_i = red_i[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
      }
      if (lane == 0) *token = (long long)best_i;
    }
  }
}

}  // namespace b2l

using namespace b2l;

static int launch_topk(const void* logits, float temperature, int top_k, void* probs, const void* noise, void* token, int V,
                       b2l_stream_t stream, const char* who) {
  B2L_CHECK_ARG(logits && V > 0 && temperature > 0.f && top_k >= 0, "%s: bad argument", who);
  B2L_CHECK_ARG(((uintptr_t)logits % 16) == 0, "%s: logits must be 16-byte aligned", who);
  B2L_CHECK_ARG(noise == nullptr || ((uintptr_t)noise % 16) == 0, "%s: noise must be 16-byte aligned", who);
  const size_t Vp = ((size_t)V + 7) & ~(size_t)7;
  const size_t smem = Vp * 2 * (noise != nullptr ? 2 : 1);
  B2L_CHECK_SUPPORTED(smem <= 200 * 1024, "%s: vocabulary %d too large for one CTA", who, V);
  static DynSmemCache smem_cache;
  if (smem > 48 * 1024)
    if (int rc = ensure_dyn_smem(topk_softmax_kernel, smem, smem_cache)) return rc;
  topk_softmax_kernel<<<1, SAMP_THREADS, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)logits, 1.0f / temperature, top_k,
                                                                      (__nv_bfloat16*)probs, (const __nv_bfloat16*)noise,
                                                                      (long long*)token, V);
  B2L_LAUNCH_CHECK("topk_softmax_kernel");
  return 0;
}

extern "C" int b2l_topk_softmax(const void* logits, float temperature, int top_k, void* probs, int V, b2l_stream_t stream) {
  B2L_CHECK_ARG(probs != nullptr, "b2l_topk_softmax: null probs");
  return launch_topk(logits, temperature, top_k, probs, nullptr, nullptr, V, stream, "b2l_topk_softmax");
}

extern "C" int b2l_topk_softmax_sample(const void* logits, float temperature, int top_k, const void* noise, void* probs,
                                       int64_t* token, int V, b2l_stream_t stream) {
  B2L_CHECK_ARG(noise != nullptr && token != nullptr, "b2l_topk_softmax_sample: null noise / token");
  return launch_topk(logits, temperature, top_k, probs, noise, token, V, stream, "b2l_topk_softmax_sample");
}
