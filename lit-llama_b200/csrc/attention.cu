// CausalSelfAttention.forward between the two linears (model.py:197-232):
// RoPE (model.py:306-323), KV-cache append with the roll branch as a device-side ring
// (model.py:211-221), and masked softmax(q k^T / sqrt(hs)) v (model.py:230) as a
// split-S streaming kernel that only touches the valid slots 0..pos.
//
// All positions are read on the device; the host never synchronises (the reference
// does once per layer per token, model.py:214).
#include "b2l_common.cuh"

namespace b2l {

__global__ void ring_advance_kernel(const int64_t* __restrict__ input_pos, int T, int32_t* ring_start, int S) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (input_pos[T - 1] >= (int64_t)S) *ring_start = (*ring_start + 1) % S;
  }
}

// grid (B*T, n_head), block hs/2 threads (one per rotated pair).
// q is rotated in place inside qkv; k (rotated) and v go to the cache (or, without a
// cache, k is rotated in place as well).
__global__ void rope_append_kernel(__nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ k_cache,
                                   __nv_bfloat16* __restrict__ v_cache, const float* __restrict__ rope,
                                   const int64_t* __restrict__ input_pos, const int32_t* __restrict__ ring_start,
                                   int T, int n_head, int hs, int S, int block_size, int rope_rows) {
  const int bt = blockIdx.x, h = blockIdx.y, b = bt / T, t = bt % T;
  const int C = n_head * hs;
  long long p = input_pos ? input_pos[t] : (long long)t;
  // rope_rows: `rope` already holds the T selected rows (reference call convention, model.py:93)
  const long long prow = rope_rows ? (long long)t : (p < block_size ? p : (long long)block_size - 1);
  __nv_bfloat16* q = qkv + (size_t)bt * 3 * C + h * hs;
  __nv_bfloat16* k = q + C;
  const __nv_bfloat16* v = q + 2 * C;
  __nv_bfloat16 *kd = k, *vd = nullptr;
  if (k_cache != nullptr) {
    const int w = (int)(p < S ? p : (long long)S - 1);
    const int phys = (w + *ring_start) % S;
    const size_t off = (((size_t)b * n_head + h) * S + phys) * hs;
    kd = k_cache + off;
    vd = v_cache + off;
  }
  for (int i = threadIdx.x; i < hs / 2; i += blockDim.x) {
    const float c = rope[((size_t)prow * (hs / 2) + i) * 2 + 0];
    const float s = rope[((size_t)prow * (hs / 2) + i) * 2 + 1];
    const float q0 = bf2f(q[2 * i]), q1 = bf2f(q[2 * i + 1]);
    const float k0 = bf2f(k[2 * i]), k1 = bf2f(k[2 * i + 1]);
    // model.py:315-318 (separate multiplies and add/sub in fp32, then type_as(x))
    q[2 * i] = f2bf(__fsub_rn(__fmul_rn(q0, c), __fmul_rn(q1, s)));
    q[2 * i + 1] = f2bf(__fadd_rn(__fmul_rn(q1, c), __fmul_rn(q0, s)));
    kd[2 * i] = f2bf(__fsub_rn(__fmul_rn(k0, c), __fmul_rn(k1, s)));
    kd[2 * i + 1] = f2bf(__fadd_rn(__fmul_rn(k1, c), __fmul_rn(k0, s)));
    if (vd != nullptr) {
      vd[2 * i] = v[2 * i];
      vd[2 * i + 1] = v[2 * i + 1];
    }
  }
}

struct KvView {
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  size_t b_stride, h_stride, s_stride;  // elements
  int S;                                // ring modulus (0 = no ring)
};

constexpr int ATT_WARPS = 4;
constexpr int ATT_MAX_EPL = 8;  // head_size <= 256

// grid (B*n_head, T, n_split).  Each CTA streams its chunk of the valid keys of query
// (b, t, h) with an online softmax per warp, merges its warps, and writes one partial
// (max, sum, acc[hs]) to `work`.
template <int EPL>  // elements per lane: head_size == 32*EPL when VEC, else generic
__global__ void __launch_bounds__(ATT_WARPS * 32)
    attn_partial_kernel(const __nv_bfloat16* __restrict__ qkv, KvView kv, const int64_t* __restrict__ input_pos,
                        const int32_t* __restrict__ ring_start, float* __restrict__ work, int T, int n_head,
                        int hs, int n_split, int chunk) {
  const int bh = blockIdx.x, b = bh / n_head, h = bh % n_head, t = blockIdx.y, sp = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = n_head * hs;
  long long p = input_pos ? input_pos[t] : (long long)t;
  const int cap = kv.S > 0 ? kv.S : T;
  const int L = (int)(p < cap ? p : (long long)cap - 1) + 1;  // valid logical slots 0..L-1
  const int ring = (kv.S > 0 && ring_start) ? *ring_start : 0;
  const int j0 = sp * chunk, j1 = min(L, j0 + chunk);
  float* out = work + (((size_t)bh * T + t) * n_split + sp) * (hs + 2);
  if (j0 >= j1) {  // empty split: neutral partial
    if (threadIdx.x == 0) { out[0] = -INFINITY; out[1] = 0.f; }
    for (int d = threadIdx.x; d < hs; d += blockDim.x) out[2 + d] = 0.f;
    return;
  }
  const float scale = rsqrtf((float)hs);
  float qv[ATT_MAX_EPL];
  const __nv_bfloat16* q = qkv + ((size_t)b * T + t) * 3 * C + h * hs;
#pragma unroll
  for (int e = 0; e < ATT_MAX_EPL; ++e) {
    int d = (EPL > 0) ? lane * EPL + e : lane + 32 * e;
    qv[e] = (e < (EPL > 0 ? EPL : ATT_MAX_EPL) && d < hs) ? bf2f(q[d]) * scale : 0.f;
  }
  float m = -INFINITY, l = 0.f, acc[ATT_MAX_EPL];
#pragma unroll
  for (int e = 0; e < ATT_MAX_EPL; ++e) acc[e] = 0.f;
  const __nv_bfloat16* kb = kv.k + b * kv.b_stride + h * kv.h_stride;
  const __nv_bfloat16* vb = kv.v + b * kv.b_stride + h * kv.h_stride;
  for (int j = j0 + warp; j < j1; j += ATT_WARPS) {
    int phys = j;
    if (kv.S > 0) { phys = j + ring; if (phys >= kv.S) phys -= kv.S; }
    const __nv_bfloat16* kr = kb + (size_t)phys * kv.s_stride;
    const __nv_bfloat16* vr = vb + (size_t)phys * kv.s_stride;
    float kx[ATT_MAX_EPL], vx[ATT_MAX_EPL];
    if constexpr (EPL == 4) {
      uint2 ku = *reinterpret_cast<const uint2*>(kr + lane * 4);
      uint2 vu = *reinterpret_cast<const uint2*>(vr + lane * 4);
      kx[0] = __uint_as_float(ku.x << 16); kx[1] = __uint_as_float(ku.x & 0xffff0000u);
      kx[2] = __uint_as_float(ku.y << 16); kx[3] = __uint_as_float(ku.y & 0xffff0000u);
      vx[0] = __uint_as_float(vu.x << 16); vx[1] = __uint_as_float(vu.x & 0xffff0000u);
      vx[2] = __uint_as_float(vu.y << 16); vx[3] = __uint_as_float(vu.y & 0xffff0000u);
    } else {
#pragma unroll
      for (int e = 0; e < ATT_MAX_EPL; ++e) {
        int d = lane + 32 * e;
        kx[e] = d < hs ? bf2f(kr[d]) : 0.f;
        vx[e] = d < hs ? bf2f(vr[d]) : 0.f;
      }
    }
    float sc = 0.f;
#pragma unroll
    for (int e = 0; e < (EPL > 0 ? EPL : ATT_MAX_EPL); ++e) sc = fmaf(qv[e], kx[e], sc);
    sc = warp_sum(sc);
    const float mn = fmaxf(m, sc);
    const float corr = __expf(m - mn), pj = __expf(sc - mn);
    l = l * corr + pj;
#pragma unroll
    for (int e = 0; e < (EPL > 0 ? EPL : ATT_MAX_EPL); ++e) acc[e] = fmaf(pj, vx[e], acc[e] * corr);
    m = mn;
  }
  // merge the warps
  __shared__ float sm_m[ATT_WARPS], sm_l[ATT_WARPS];
  __shared__ float sm_acc[ATT_WARPS][32 * ATT_MAX_EPL];
  if (lane == 0) { sm_m[warp] = m; sm_l[warp] = l; }
#pragma unroll
  for (int e = 0; e < (EPL > 0 ? EPL : ATT_MAX_EPL); ++e) {
    int d = (EPL > 0) ? lane * EPL + e : lane + 32 * e;
    sm_acc[warp][d] = acc[e];
  }
  __syncthreads();
  float M = -INFINITY;
#pragma unroll
  for (int w = 0; w < ATT_WARPS; ++w) M = fmaxf(M, sm_m[w]);
  float Ls = 0.f;
  float wgt[ATT_WARPS];
#pragma unroll
  for (int w = 0; w < ATT_WARPS; ++w) {
    wgt[w] = (sm_m[w] == -INFINITY) ? 0.f : __expf(sm_m[w] - M);
    Ls += sm_l[w] * wgt[w];
  }
  if (threadIdx.x == 0) { out[0] = M; out[1] = Ls; }
  for (int d = threadIdx.x; d < hs; d += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < ATT_WARPS; ++w) a += sm_acc[w][d] * wgt[w];
    out[2 + d] = a;
  }
}

// grid (B*n_head, T): merge the split partials and write y[b][t][h*hs + d] in bf16.
__global__ void attn_combine_kernel(const float* __restrict__ work, __nv_bfloat16* __restrict__ y, int T,
                                    int n_head, int hs, int n_split) {
  // the kernel after this one (attn.c_proj) may start streaming its weights now
  pdl_launch_dependents();
  const int bh = blockIdx.x, b = bh / n_head, h = bh % n_head, t = blockIdx.y;
  const float* base = work + ((size_t)bh * T + t) * n_split * (hs + 2);
  float M = -INFINITY;
  for (int s = 0; s < n_split; ++s) M = fmaxf(M, base[(size_t)s * (hs + 2)]);
  float Ls = 0.f;
  for (int s = 0; s < n_split; ++s) {
    float ms = base[(size_t)s * (hs + 2)];
    if (ms != -INFINITY) Ls += base[(size_t)s * (hs + 2) + 1] * __expf(ms - M);
  }
  const float inv = 1.0f / Ls;
  __nv_bfloat16* yr = y + ((size_t)b * T + t) * (n_head * hs) + h * hs;
  for (int d = threadIdx.x; d < hs; d += blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < n_split; ++s) {
      float ms = base[(size_t)s * (hs + 2)];
      if (ms != -INFINITY) a += base[(size_t)s * (hs + 2) + 2 + d] * __expf(ms - M);
    }
    yr[d] = f2bf(a * inv);
  }
}

__global__ void kv_unroll_kernel(const __nv_bfloat16* __restrict__ cache, const int32_t* __restrict__ ring_start,
                                 __nv_bfloat16* __restrict__ out, int S, int hs) {
  // grid (B*n_head, S): logical slot blockIdx.y <- physical (slot + ring) % S
  const int ring = *ring_start;
  const int phys = (blockIdx.y + ring) % S;
  const __nv_bfloat16* src = cache + ((size_t)blockIdx.x * S + phys) * hs;
  __nv_bfloat16* dst = out + ((size_t)blockIdx.x * S + blockIdx.y) * hs;
  for (int d = threadIdx.x; d < hs; d += blockDim.x) dst[d] = src[d];
}

static inline void split_plan(int T, int S, int* n_split, int* chunk) {
  if (T > 1) { *n_split = 1; *chunk = S; return; }
  *chunk = 64;
  *n_split = (S + 63) / 64;
}

static int launch_attn(const __nv_bfloat16* qkv, KvView kv, const int64_t* input_pos, const int32_t* ring_start,
                       float* work, __nv_bfloat16* y, int B, int T, int n_head, int hs, int cap, cudaStream_t st) {
  int n_split, chunk;
  split_plan(T, cap, &n_split, &chunk);
  dim3 grid(B * n_head, T, n_split), block(ATT_WARPS * 32);
  if (hs == 128)
    attn_partial_kernel<4><<<grid, block, 0, st>>>(qkv, kv, input_pos, ring_start, work, T, n_head, hs, n_split, chunk);
  else
    attn_partial_kernel<0><<<grid, block, 0, st>>>(qkv, kv, input_pos, ring_start, work, T, n_head, hs, n_split, chunk);
  B2L_LAUNCH_CHECK("attn_partial_kernel");
  attn_combine_kernel<<<dim3(B * n_head, T), 128, 0, st>>>(work, y, T, n_head, hs, n_split);
  B2L_LAUNCH_CHECK("attn_combine_kernel");
  return 0;
}

}  // namespace b2l

using namespace b2l;

extern "C" size_t b2l_attn_workspace_bytes(int B, int n_head, int head_size, int T, int S) {
  int n_split, chunk;
  split_plan(T, S, &n_split, &chunk);
  return (size_t)B * n_head * T * n_split * (head_size + 2) * sizeof(float);
}

extern "C" int b2l_ring_advance(const int64_t* input_pos, int T, int32_t* ring_start, int S, b2l_stream_t stream) {
  B2L_CHECK_ARG(input_pos && ring_start && T > 0 && S > 0, "b2l_ring_advance: bad argument");
  ring_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(input_pos, T, ring_start, S);
  B2L_LAUNCH_CHECK("ring_advance_kernel");
  return 0;
}

extern "C" int b2l_attention(void* qkv, void* k_cache, void* v_cache, const void* rope, const int64_t* input_pos,
                             const int32_t* ring_start, void* y, void* work, int B, int T, int n_head,
                             int head_size, int S, int block_size, int flags, b2l_stream_t stream) {
  B2L_CHECK_ARG(qkv && k_cache && v_cache && rope && input_pos && ring_start && y && work,
                "b2l_attention: null pointer");
  B2L_CHECK_ARG(B > 0 && T > 0 && n_head > 0 && S > 0 && T <= S && block_size > 0, "b2l_attention: bad shape");
  B2L_CHECK_SUPPORTED(head_size % 2 == 0 && head_size >= 2 && head_size <= 32 * ATT_MAX_EPL,
                      "b2l_attention: head_size %d unsupported (even, <= %d)", head_size, 32 * ATT_MAX_EPL);
  cudaStream_t st = (cudaStream_t)stream;
  int rt = head_size / 2 < 32 ? 32 : head_size / 2;
  rope_append_kernel<<<dim3(B * T, n_head), rt, 0, st>>>((__nv_bfloat16*)qkv, (__nv_bfloat16*)k_cache,
                                                        (__nv_bfloat16*)v_cache, (const float*)rope, input_pos,
                                                        ring_start, T, n_head, head_size, S, block_size, (flags & B2L_F_ROPE_ROWS) ? 1 : 0);
  B2L_LAUNCH_CHECK("rope_append_kernel");
  KvView kv{(const __nv_bfloat16*)k_cache, (const __nv_bfloat16*)v_cache, (size_t)n_head * S * head_size,
            (size_t)S * head_size, (size_t)head_size, S};
  return launch_attn((const __nv_bfloat16*)qkv, kv, input_pos, ring_start, (float*)work, (__nv_bfloat16*)y, B, T,
                     n_head, head_size, S, st);
}

extern "C" int b2l_attention_nocache(void* qkv, const void* rope, void* y, void* work, int B, int T, int n_head,
                                     int head_size, int block_size, b2l_stream_t stream) {
  B2L_CHECK_ARG(qkv && rope && y && work && B > 0 && T > 0 && n_head > 0 && T <= block_size,
                "b2l_attention_nocache: bad argument");
  B2L_CHECK_SUPPORTED(head_size % 2 == 0 && head_size >= 2 && head_size <= 32 * ATT_MAX_EPL,
                      "b2l_attention_nocache: head_size %d unsupported", head_size);
  cudaStream_t st = (cudaStream_t)stream;
  int rt = head_size / 2 < 32 ? 32 : head_size / 2;
  rope_append_kernel<<<dim3(B * T, n_head), rt, 0, st>>>((__nv_bfloat16*)qkv, nullptr, nullptr, (const float*)rope,
                                                        nullptr, nullptr, T, n_head, head_size, 0, block_size, 0);
  B2L_LAUNCH_CHECK("rope_append_kernel");
  const int C = n_head * head_size;
  const __nv_bfloat16* base = (const __nv_bfloat16*)qkv;
  KvView kv{base + C, base + 2 * C, (size_t)T * 3 * C, (size_t)head_size, (size_t)3 * C, 0};
  return launch_attn(base, kv, nullptr, nullptr, (float*)work, (__nv_bfloat16*)y, B, T, n_head, head_size, T, st);
}

extern "C" int b2l_kv_unroll(const void* cache, const int32_t* ring_start, void* out, int B, int n_head, int S,
                             int head_size, b2l_stream_t stream) {
  B2L_CHECK_ARG(cache && ring_start && out && B > 0 && n_head > 0 && S > 0 && head_size > 0,
                "b2l_kv_unroll: bad argument");
  kv_unroll_kernel<<<dim3(B * n_head, S), 64, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)cache, ring_start,
                                                                         (__nv_bfloat16*)out, S, head_size);
  B2L_LAUNCH_CHECK("kv_unroll_kernel");
  return 0;
}
