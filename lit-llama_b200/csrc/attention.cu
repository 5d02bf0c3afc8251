// CausalSelfAttention.forward between the two linears (model.py:197-232):
// RoPE (model.py:306-323), KV-cache append with the roll branch as a device-side ring
// (model.py:211-221), and masked softmax(q k^T / sqrt(hs)) v (model.py:230) as a
// split-S streaming kernel that only touches the valid slots 0..pos.
//
// All positions are read on the device; the host never synchronises (the reference
// does once per layer per token, model.py:214).
#include <cstdlib>

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

// ----------------------------------------------------------------------------------
// Fused single-token attention for head_size 128 (every LLaMA size): one kernel does
// RoPE(q), RoPE(k) + in-place KV append, split-S online-softmax attention over the valid
// slots, and the cross-split merge (last CTA of a head, atomic ticket).
//
// grid (B*n_head, ceil(S / 64)), 8 warps; CTAs beyond the position-dependent split count exit at once.  A CTA owns
// 64..256 keys of one head (chosen from the position so that ~400 CTAs work) and streams them as 64-key
// sub-tiles (K 16 KB + V 16 KB) through a two-deep shared-memory ring with TMA bulk copies: the first two
// sub-tiles are requested BEFORE griddepcontrol.wait (old cache rows do not depend on the current token), the
// next one as soon as a buffer has been consumed.  A warp handles 4 keys per round: 8 lanes per key, 16 head
// dims (32 B) per lane, a score needs 3 shuffles.  The new token's key / value never touch the tile: they are
// rotated, appended to the cache and scored from registers.
// Round 1 used one 128-key tile per CTA (64 KB, 3 CTAs per SM): at position 2047 its 512 CTAs needed a second
// wave and 16 partials per head had to be merged (profiles/r01_ncu_attn_decode_fused_kernel.txt: 0.19 of HBM
// peak); 256 keys per CTA keep every position a single wave (<= 8 x n_head CTAs) with half the partials.
// ----------------------------------------------------------------------------------
constexpr int FD_CHUNK = 256;   // most keys per CTA
constexpr int FD_SUB = 64;      // keys per sub-tile = smallest number of keys per CTA
constexpr int FD_WARPS = 8;
constexpr int FD_TARGET_CTAS = 400;  // working CTAs aimed at (148 SMs x 3 resident CTAs = 444 slots: one wave)
constexpr int WS_CHUNK = FD_SUB;     // workspace sizing granularity (finest split of any kernel that uses it)

__device__ __forceinline__ void bf16x8_to_f32(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

__device__ __forceinline__ uint32_t fd_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fd_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Dynamic shared memory: 2 buffers x (K sub-tile 16 KB + V sub-tile 16 KB), merge scratch, 2 mbarriers.
constexpr int FD_SUB_BYTES = FD_SUB * 128 * 2;
constexpr int FD_SMEM_BYTES = 4 * FD_SUB_BYTES + FD_WARPS * 128 * 4 + 2 * FD_WARPS * 4 + 16;

__global__ void __launch_bounds__(FD_WARPS * 32)
    attn_decode_fused_kernel(const __nv_bfloat16* qkv, __nv_bfloat16* __restrict__ k_cache,
                             __nv_bfloat16* __restrict__ v_cache, const float* __restrict__ rope,
                             const int64_t* __restrict__ input_pos, const int32_t* __restrict__ ring_start,
                             __nv_bfloat16* __restrict__ y, float* __restrict__ work, int* __restrict__ tickets,
                             int n_head, int S, int block_size, int n_split, unsigned long long* tl, int pre_tiles,
                             int smem_merge) {
  constexpr int HS = 128;
  extern __shared__ __align__(128) uint8_t fsm[];
  float* sm_acc = reinterpret_cast<float*>(fsm + 4 * FD_SUB_BYTES);                // [FD_WARPS][HS]
  float* sm_m = sm_acc + FD_WARPS * HS;                                            // [FD_WARPS]
  float* sm_l = sm_m + FD_WARPS;                                                   // [FD_WARPS]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm_l + FD_WARPS);  // 8-byte aligned by construction
  __shared__ int sm_last;

  if (threadIdx.x == 0) tl_min(tl, 0);
  const int bh = blockIdx.x, b = bh / n_head, h = bh % n_head, sp = blockIdx.y;
  // debug: raw stamps of CTA (0, 0) in slots 8.. (start, wait done, q ready, per sub-tile: data seen / done, merged, end)
  unsigned long long* tl0 = (tl != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) ? tl + 8 : nullptr;
  int tli = 0;
  auto stamp = [&]() { if (tl0 != nullptr && tli < 24) tl0[tli++] = globaltimer_ns(); };
  stamp();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = n_head * HS;
  const size_t head_base = ((size_t)b * n_head + h) * S * HS;

  // input_pos and ring_start are inputs of the step (written by the host side long before), not
  // products of the previous kernel: they may be read before the dependency is resolved.
  const long long p = input_pos[0];
  const int w_slot = (int)(p < S ? p : (long long)S - 1);  // logical slot of the new token
  const int L = w_slot + 1;                                 // valid logical slots 0..L-1
  // keys per CTA: a multiple of 64 in [64, 256], chosen (identically by every CTA) so that about FD_TARGET_CTAS CTAs
  // have work: few long chunks would serialise sub-tiles inside a CTA, many short ones would need a second wave
  // (measured, tools/diag.py bench_ctx: a sub-tile costs a CTA ~0.7 us, the cross-CTA merge ~3 us -- up to 256
  // keys stay in ONE CTA per head, with no merge at all)
  const int want_splits = max(1, FD_TARGET_CTAS / (int)gridDim.x);
  const int chunk = L <= FD_CHUNK ? FD_CHUNK
                                  : min(FD_CHUNK, max(FD_SUB, FD_SUB * ((L + FD_SUB * want_splits - 1) / (FD_SUB * want_splits))));
  const int n_active = (L + chunk - 1) / chunk;
  if (sp >= n_active) return;
  const int ring = *ring_start;
  const int j0 = sp * chunk, j1 = min(L, j0 + chunk);
  const int n_old = min(j1, L - 1) - j0;  // rows written by earlier steps (slot L-1 is written by this one)
  const int n_sub = (n_old + FD_SUB - 1) / FD_SUB;
  const bool has_new = (w_slot >= j0 && w_slot < j1);

  // sub-tile i -> buffer i & 1: old K rows then old V rows, each possibly in two pieces (the ring wraps)
  const uint32_t bar0 = fd_smem_u32(bars);
  auto request = [&](int i) {
    const int buf = i & 1;
    const int cnt = min(FD_SUB, n_old - i * FD_SUB);
    int phys0 = j0 + i * FD_SUB + ring; if (phys0 >= S) phys0 -= S;
    const int first = min(cnt, S - phys0);  // rows before the ring wraps
    const uint32_t bar = bar0 + buf * 8;
    const uint32_t kd = fd_smem_u32(fsm) + buf * 2 * FD_SUB_BYTES, vd = kd + FD_SUB_BYTES;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)cnt * HS * 2 * 2) : "memory");
    fd_bulk(kd, k_cache + head_base + (size_t)phys0 * HS, (uint32_t)first * HS * 2, bar);
    fd_bulk(vd, v_cache + head_base + (size_t)phys0 * HS, (uint32_t)first * HS * 2, bar);
    if (first < cnt) {  // wrapped part starts at physical row 0
      fd_bulk(kd + first * HS * 2, k_cache + head_base, (uint32_t)(cnt - first) * HS * 2, bar);
      fd_bulk(vd + first * HS * 2, v_cache + head_base, (uint32_t)(cnt - first) * HS * 2, bar);
    }
  };
  // ---- before the dependency: the first sub-tile (both when one CTA per head is all there is: nothing queues then)
  const int pre = (n_active == 1) ? 2 : pre_tiles;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (n_sub > 0) request(0);
    if (pre > 1 && n_sub > 1) request(1);
  }
  // the RoPE row of this position is a constant table entry: fetch it before the dependency too
  const long long prow = p < block_size ? p : (long long)block_size - 1;
  const int grp = lane >> 3, sub = lane & 7, d0 = sub * 16;
  float cs[16];
  {
    const float4* rp = reinterpret_cast<const float4*>(rope + ((size_t)prow * (HS / 2) + d0 / 2) * 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 t = rp[i];
      cs[4 * i] = t.x; cs[4 * i + 1] = t.y; cs[4 * i + 2] = t.z; cs[4 * i + 3] = t.w;
    }
  }
  __syncthreads();  // the barriers are initialised before anyone polls them
  pdl_wait();
  if (threadIdx.x == 0) tl_max(tl, 1);
  stamp();
  pdl_launch_dependents();  // attn.c_proj may start streaming its weights

  const __nv_bfloat16* qrow = qkv + (size_t)b * 3 * C + h * HS + d0;
  float q[16];
  {
    float raw[16];
    // qkv is the output of the kernel this one was launched behind (PDL): coherent loads only
    const uint4 qa = ld_coherent_u4(qrow), qb = ld_coherent_u4(qrow + 8);
    // The second sub-tile is requested BEHIND the q loads: a reply to this SM queues behind the bulk data already
    // on its way (measured, tools/diag.py timeline: with two 32 KB sub-tiles per CTA in flight the 32-byte q load
    // came back after 2.4 us at position 1033, 0.5 us with one), and q is what the first score needs.
    if (threadIdx.x == 0 && pre <= 1 && n_sub > 1) request(1);
    bf16x8_to_f32(qa, raw);
    bf16x8_to_f32(qb, raw + 8);
    const float scale = rsqrtf((float)HS);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float c = cs[2 * i], s_ = cs[2 * i + 1];
      const float e = rbf(__fsub_rn(__fmul_rn(raw[2 * i], c), __fmul_rn(raw[2 * i + 1], s_)));
      const float o = rbf(__fadd_rn(__fmul_rn(raw[2 * i + 1], c), __fmul_rn(raw[2 * i], s_)));
      q[2 * i] = e * scale;
      q[2 * i + 1] = o * scale;
    }
  }

  float m = -INFINITY, l = 0.f, acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  if (tl0 != nullptr && q[0] != 12345.678f) stamp();   // q ready (the comparison keeps the stamp behind the loads)
  // ---- the new token FIRST (its k / v loads travel with the q loads, while the old tiles are still landing):
  // rotate k, append k and v to the cache, score from registers (warp 0, key group 0)
  if (has_new && warp == 0 && grp == 0) {
    int phys = w_slot + ring; if (phys >= S) phys -= S;
    float raw[16], kf[16], vf[16];
    bf16x8_to_f32(ld_coherent_u4(qrow + C), raw);
    bf16x8_to_f32(ld_coherent_u4(qrow + C + 8), raw + 8);
    uint32_t out[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float c = cs[2 * i], s_ = cs[2 * i + 1];
      kf[2 * i] = rbf(__fsub_rn(__fmul_rn(raw[2 * i], c), __fmul_rn(raw[2 * i + 1], s_)));
      kf[2 * i + 1] = rbf(__fadd_rn(__fmul_rn(raw[2 * i + 1], c), __fmul_rn(raw[2 * i], s_)));
      out[i] = (__float_as_uint(kf[2 * i]) >> 16) | (__float_as_uint(kf[2 * i + 1]) & 0xffff0000u);
    }
    const uint4 va = ld_coherent_u4(qrow + 2 * C), vb = ld_coherent_u4(qrow + 2 * C + 8);
    uint4* kd = reinterpret_cast<uint4*>(k_cache + head_base + (size_t)phys * HS + d0);
    uint4* vd = reinterpret_cast<uint4*>(v_cache + head_base + (size_t)phys * HS + d0);
    kd[0] = make_uint4(out[0], out[1], out[2], out[3]); kd[1] = make_uint4(out[4], out[5], out[6], out[7]);
    vd[0] = va; vd[1] = vb;
    bf16x8_to_f32(va, vf); bf16x8_to_f32(vb, vf + 8);
    float sc = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) sc = fmaf(q[e], kf[e], sc);
    sc += __shfl_xor_sync(0x000000ffu, sc, 1);
    sc += __shfl_xor_sync(0x000000ffu, sc, 2);
    sc += __shfl_xor_sync(0x000000ffu, sc, 4);
    const float mn = fmaxf(m, sc);
    const float corr = __expf(m - mn), pj = __expf(sc - mn);
    l = l * corr + pj;
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = fmaf(pj, vf[e], acc[e] * corr);
    m = mn;
  }
  // ---- old rows, sub-tile by sub-tile: warp w takes rows 8 w .. 8 w + 7 (4 keys per round, 8 lanes per key)
  for (int i = 0; i < n_sub; ++i) {
    const int buf = i & 1;
    const int cnt = min(FD_SUB, n_old - i * FD_SUB);
    {
      uint32_t ok;
      const uint32_t par = (uint32_t)(i >> 1) & 1u;
      do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar0 + buf * 8), "r"(par) : "memory");
      } while (!ok);
    }
    const __nv_bfloat16* kt = reinterpret_cast<const __nv_bfloat16*>(fsm + buf * 2 * FD_SUB_BYTES);
    const __nv_bfloat16* vt = kt + FD_SUB * HS;
    stamp();
    // the warp's 8 keys of this sub-tile as two groups of 4 (8 lanes per key): both scores first (independent chains),
    // ONE running-max update, then both value rows -- half the dependent (m, l, acc) updates of a key-by-key loop
    float sc[2];
    bool valid[2];
    const uint4* vr[2];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int r = warp * 8 + it * 4 + grp;
      valid[it] = r < cnt;
      const int rc = valid[it] ? r : 0;
      const uint4* kr = reinterpret_cast<const uint4*>(kt + (size_t)rc * HS + d0);
      vr[it] = reinterpret_cast<const uint4*>(vt + (size_t)rc * HS + d0);
      float kf[16];
      bf16x8_to_f32(kr[0], kf); bf16x8_to_f32(kr[1], kf + 8);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) { a0 = fmaf(q[e], kf[e], a0); a1 = fmaf(q[8 + e], kf[8 + e], a1); }
      sc[it] = a0 + a1;
    }
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1) {
      sc[0] += __shfl_xor_sync(0xffffffffu, sc[0], o);
      sc[1] += __shfl_xor_sync(0xffffffffu, sc[1], o);
    }
    {
      const float s0 = valid[0] ? sc[0] : -INFINITY, s1 = valid[1] ? sc[1] : -INFINITY;
      const float mn = fmaxf(m, fmaxf(s0, s1));
      if (mn != -INFINITY) {   // at least one key so far in this lane group
        const float corr = __expf(m - mn), p0 = __expf(s0 - mn), p1 = __expf(s1 - mn);   // exp(-inf) = 0
        float v0[16], v1[16];
        bf16x8_to_f32(vr[0][0], v0); bf16x8_to_f32(vr[0][1], v0 + 8);
        bf16x8_to_f32(vr[1][0], v1); bf16x8_to_f32(vr[1][1], v1 + 8);
        l = l * corr + p0 + p1;
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = fmaf(p1, v1[e], fmaf(p0, v0[e], acc[e] * corr));
        m = mn;
      }
    }
    __syncthreads();   // every warp is done with this buffer
    stamp();
    if (threadIdx.x == 0 && i + 2 < n_sub) request(i + 2);
  }
  if (threadIdx.x == 0) tl_max(tl, 2);
  __syncwarp();
  if (threadIdx.x == 0) tl_max(tl, 3);
  // ---- merge the 32 (warp, key group) partials of the CTA.  smem_merge: every group leaves its row in shared memory (the
  // K / V ring is idle now: every requested sub-tile has been consumed behind a __syncthreads), one warp turns the 32
  // running maxima into weights, 128 threads add.  Otherwise: the 4 key groups of a warp merge with shuffles first
  // (36 per warp), then the 8 warps through shared memory -- the default: measured 1.1 % faster per token at position
  // 1030 on the same box (tools/env_sweep.sh "B2L_ATTN_SMEM_MERGE=0 B2L_ATTN_SMEM_MERGE=1").
  float* part = reinterpret_cast<float*>(fsm);                 // [32][HS]
  float* pm = part + 32 * HS;                                  // [32] running max, [32] sum, [32] weight, M, Ls
  float* pl = pm + 32, *pw = pl + 32;
  const int n_part = smem_merge ? 32 : FD_WARPS;
  if (smem_merge) {
    const int g = warp * 4 + grp;
    float4* dst = reinterpret_cast<float4*>(part + g * HS + d0);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
    if (sub == 0) { pm[g] = m; pl[g] = l; }
  } else {
#pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {   // lanes with the same `sub` hold the same dims
      const float mo = __shfl_xor_sync(0xffffffffu, m, off);
      const float lo = __shfl_xor_sync(0xffffffffu, l, off);
      const float mn = fmaxf(m, mo);
      const float ca = (m == -INFINITY) ? 0.f : __expf(m - mn);
      const float cb = (mo == -INFINITY) ? 0.f : __expf(mo - mn);
      l = l * ca + lo * cb;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float ao = __shfl_xor_sync(0xffffffffu, acc[i], off);
        acc[i] = acc[i] * ca + ao * cb;
      }
      m = mn;
    }
    if (lane == 0) { pm[warp] = m; pl[warp] = l; }
    if (grp == 0) {
      float4* dst = reinterpret_cast<float4*>(part + warp * HS + d0);
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
    }
  }
  __syncthreads();
  if (warp == 0) {
    const float mi = lane < n_part ? pm[lane] : -INFINITY;
    const float Mx = warp_max(mi);
    const float wi = (mi == -INFINITY) ? 0.f : __expf(mi - Mx);
    const float Lx = warp_sum(lane < n_part ? pl[lane] * wi : 0.f);
    pw[lane] = wi;
    if (lane == 0) { pw[32] = Mx; pw[33] = Lx; }
  }
  __syncthreads();
  const float M = pw[32], Ls = pw[33];
  const int d = threadIdx.x & (HS - 1);
  const bool writer = threadIdx.x < HS;
  float a = 0.f;
  if (writer) {
#pragma unroll 8
    for (int g = 0; g < n_part; ++g) a = fmaf(part[g * HS + d], pw[g], a);
  }
  stamp();
  if (n_active == 1) {  // nothing to merge
    if (writer) y[(size_t)b * C + h * HS + d] = f2bf(a / Ls);
    if (threadIdx.x == 0) tl_max(tl, 4);
    stamp();
    return;
  }
  float* out = work + ((size_t)bh * n_split + sp) * (HS + 2);
  if (threadIdx.x == 0) { out[0] = M; out[1] = Ls; }
  if (writer) out[2 + d] = a;
  __syncthreads();  // all partial stores of this CTA are ordered before the ticket (cumulative release below)
  if (threadIdx.x == 0) {
    int t;
    asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(t) : "l"(tickets + bh) : "memory");
    sm_last = (t == n_active - 1);
    if (sm_last) tickets[bh] = 0;  // every contributor has arrived: safe to re-arm for the next step
  }
  __syncthreads();  // thread 0's acquire + this barrier order the other CTAs' partials before the loads below
  stamp();
  if (!sm_last) return;
  // merge: every load is issued before the first use (n_split <= 8 for S <= 2048; larger S loops in batches)
  const float* base = work + (size_t)bh * n_split * (HS + 2);
  float MM = -INFINITY, LL = 0.f, aa = 0.f;
  for (int s0 = 0; s0 < n_active; s0 += 8) {
    float ms[8], ls[8], as[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int s2 = s0 + i;
      const bool ok = s2 < n_active;
      const float* bp = base + (size_t)(ok ? s2 : s0) * (HS + 2);
      ms[i] = ok ? __ldcg(bp) : -INFINITY;
      ls[i] = ok ? __ldcg(bp + 1) : 0.f;
      as[i] = ok ? __ldcg(bp + 2 + d) : 0.f;
    }
    float bm = MM;
#pragma unroll
    for (int i = 0; i < 8; ++i) bm = fmaxf(bm, ms[i]);
    const float c0 = (MM == -INFINITY) ? 0.f : __expf(MM - bm);
    LL *= c0; aa *= c0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float wg = (ms[i] == -INFINITY) ? 0.f : __expf(ms[i] - bm);
      LL += ls[i] * wg;
      aa += as[i] * wg;
    }
    MM = bm;
  }
  if (writer) y[(size_t)b * C + h * HS + d] = f2bf(aa / LL);
  if (threadIdx.x == 0) tl_max(tl, 4);
}

// ----------------------------------------------------------------------------------
// Prefill / no-cache attention for head_size 128 and T > 1 (model.py:200-230 over many query positions): a tiled
// online-softmax kernel on the tensor cores (mma.sync.m16n8k16 bf16, fp32 accumulate).  Round 1 gave every query row
// its own CTA that re-read all its keys from L2 (13B, 8 x 512 tokens: 153 ms of a 315 ms prefill); here a CTA owns
// 64 query rows of one head (4 warps x 16 rows), streams 64-key K / V tiles through a double-buffered shared-memory
// ring (cp.async, rows gathered through the cache ring), computes S = Q K^T and O += P V with ldmatrix-fed MMAs.
// Scores are scaled and exponentiated in fp32; P enters the second MMA as bf16 hi + lo parts (2^-17 relative), so the
// result matches an fp32 softmax(q k^T / sqrt(hs)) v to the final bf16 rounding.
// q is read from qkv (already rotated by rope_append_kernel); the mask is "slot <= position of the query"
// (model.py:94-96 through the tril rows selected by input_pos).
// ----------------------------------------------------------------------------------
constexpr int PF_Q = 64, PF_K = 64, PF_LD = 136;               // padded row: 128 + 8 elements (272 B) -> conflict-free ldmatrix
constexpr int PF_TILE_BYTES = PF_Q * PF_LD * 2;                 // 17408
constexpr int PF_SMEM_BYTES = 5 * PF_TILE_BYTES;                // Q + 2 x (K, V)

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void pf_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// bf16 hi / lo split of two fp32 values: hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ void split_bf16x2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - __low2float(h), y - __high2float(h));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(128)
    attn_prefill_kernel(const __nv_bfloat16* __restrict__ qkv, KvView kv, const int64_t* __restrict__ input_pos,
                        const int32_t* __restrict__ ring_start, __nv_bfloat16* __restrict__ y, int T, int n_head) {
  constexpr int HS = 128;
  extern __shared__ __align__(128) uint8_t psm[];
  const uint32_t sq = fd_smem_u32(psm), sk0 = sq + PF_TILE_BYTES;   // K tile of buffer b at sk0 + 2 b TILE, V tile right behind it
  const int bh = blockIdx.x, b = bh / n_head, h = bh % n_head, t0 = blockIdx.y * PF_Q;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int C = n_head * HS;
  const int cap = kv.S > 0 ? kv.S : T;
  const int ring = (kv.S > 0 && ring_start) ? *ring_start : 0;
  const __nv_bfloat16* kb = kv.k + (size_t)b * kv.b_stride + (size_t)h * kv.h_stride;
  const __nv_bfloat16* vb = kv.v + (size_t)b * kv.b_stride + (size_t)h * kv.h_stride;

  // valid slots of this thread's two query rows (g and g + 8 of the warp's 16), and of the whole CTA
  int Lrow[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int t = min(t0 + warp * 16 + g + 8 * i, T - 1);
    const long long p = input_pos ? input_pos[t] : (long long)t;
    Lrow[i] = (int)(p < cap ? p : (long long)cap - 1) + 1;
  }
  __shared__ int s_lmax;
  if (tid == 0) s_lmax = 0;
  __syncthreads();
  atomicMax(&s_lmax, max(Lrow[0], Lrow[1]));
  // ---- Q tile (rows beyond T repeat row T - 1: never stored)
  for (int q = tid; q < PF_Q * 16; q += 128) {
    const int r = q >> 4, c = q & 15;
    const int t = min(t0 + r, T - 1);
    pf_cp16(sq + (r * PF_LD + c * 8) * 2, qkv + ((size_t)b * T + t) * 3 * C + h * HS + c * 8);
  }
  __syncthreads();
  const int Lmax = s_lmax;
  const int n_kb = (Lmax + PF_K - 1) / PF_K;
  auto load_kv = [&](int kbi, int buf) {
    const uint32_t dk = sk0 + buf * 2 * PF_TILE_BYTES, dv = dk + PF_TILE_BYTES;
    for (int q = tid; q < PF_K * 16; q += 128) {
      const int r = q >> 4, c = q & 15;
      int j = min(kbi * PF_K + r, Lmax - 1);     // rows beyond the last valid slot repeat it (masked below)
      if (kv.S > 0) { j += ring; if (j >= kv.S) j -= kv.S; }
      pf_cp16(dk + (r * PF_LD + c * 8) * 2, kb + (size_t)j * kv.s_stride + c * 8);
      pf_cp16(dv + (r * PF_LD + c * 8) * 2, vb + (size_t)j * kv.s_stride + c * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_kv(0, 0);   // group 0 also carries the Q tile
  uint32_t qf[8][4];
  float o[16][4];
#pragma unroll
  for (int n = 0; n < 16; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[n][i] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const float scale = rsqrtf((float)HS);

  for (int kbi = 0; kbi < n_kb; ++kbi) {
    const int buf = kbi & 1;
    if (kbi + 1 < n_kb) {
      load_kv(kbi + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (kbi == 0) {   // Q fragments: 8 k-steps of 16 dims
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int mi = lane >> 3;
        ldsm_x4(sq + ((warp * 16 + (mi & 1) * 8 + (lane & 7)) * PF_LD + ks * 16 + (mi >> 1) * 8) * 2, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    const uint32_t ksm = sk0 + buf * 2 * PF_TILE_BYTES, vsm = ksm + PF_TILE_BYTES;
    // ---- S = Q K^T for 64 keys: 8 n-tiles of 8 keys
    float sacc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) sacc[nt][i] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 8; ks += 2) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(ksm + ((nt * 8 + (lane & 7)) * PF_LD + ks * 16 + (lane >> 3) * 8) * 2, b0, b1, b2, b3);
        mma_bf16(sacc[nt], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
        mma_bf16(sacc[nt], qf[ks + 1][0], qf[ks + 1][1], qf[ks + 1][2], qf[ks + 1][3], b2, b3);
      }
    }
    // ---- scale, mask (slot < valid slots of the row), online softmax; lane (g, t4) holds rows g / g + 8, keys 8 nt + 2 t4 (+1)
    float mnew[2] = {mrow[0], mrow[1]};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int key = kbi * PF_K + nt * 8 + 2 * t4 + (i & 1);
        const float v = (key < Lrow[i >> 1]) ? sacc[nt][i] * scale : -INFINITY;
        sacc[nt][i] = v;
        mnew[i >> 1] = fmaxf(mnew[i >> 1], v);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mnew[r] = fmaxf(mnew[r], __shfl_xor_sync(0xffffffffu, mnew[r], 1));
      mnew[r] = fmaxf(mnew[r], __shfl_xor_sync(0xffffffffu, mnew[r], 2));
    }
    float corr[2], psum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) corr[r] = (mrow[r] == -INFINITY) ? 0.f : __expf(mrow[r] - mnew[r]);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float pj = (sacc[nt][i] == -INFINITY) ? 0.f : __expf(sacc[nt][i] - mnew[i >> 1]);
        sacc[nt][i] = pj;
        psum[i >> 1] += pj;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      lrow[r] = lrow[r] * corr[r] + psum[r];   // per-lane partial row sums; reduced across the quad at the end
      mrow[r] = mnew[r];
    }
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      o[n][0] *= corr[0]; o[n][1] *= corr[0]; o[n][2] *= corr[1]; o[n][3] *= corr[1];
    }
    // ---- O += P V: 4 k-steps of 16 keys, 16 n-tiles of 8 dims; P as bf16 hi + lo
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ah[4], al[4];
      split_bf16x2(sacc[2 * kk][0], sacc[2 * kk][1], ah[0], al[0]);          // row g,     keys 16 kk + 2 t4 (+1)
      split_bf16x2(sacc[2 * kk][2], sacc[2 * kk][3], ah[1], al[1]);          // row g + 8
      split_bf16x2(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1], ah[2], al[2]);  // row g,     keys 16 kk + 8 + 2 t4 (+1)
      split_bf16x2(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3], ah[3], al[3]);  // row g + 8
#pragma unroll
      for (int dn = 0; dn < 16; dn += 2) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(vsm + ((kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PF_LD + dn * 8 + (lane >> 4) * 8) * 2, b0, b1, b2, b3);
        mma_bf16(o[dn], ah[0], ah[1], ah[2], ah[3], b0, b1);
        mma_bf16(o[dn], al[0], al[1], al[2], al[3], b0, b1);
        mma_bf16(o[dn + 1], ah[0], ah[1], ah[2], ah[3], b2, b3);
        mma_bf16(o[dn + 1], al[0], al[1], al[2], al[3], b2, b3);
      }
    }
    __syncthreads();   // this buffer is refilled two iterations later
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int t = t0 + warp * 16 + g + 8 * r;
    if (t < T) {
      const float inv = 1.0f / lrow[r];
      __nv_bfloat16* dst = y + ((size_t)b * T + t) * C + h * HS + 2 * t4;
#pragma unroll
      for (int n = 0; n < 16; ++n)
        *reinterpret_cast<__nv_bfloat162*>(dst + n * 8) = __floats2bfloat162_rn(o[n][2 * r] * inv, o[n][2 * r + 1] * inv);
    }
  }
}

// debug timeline slot for the next fused-attention launch (set by b2l_decode_step; nullptr = off)
void* g_attn_timeline = nullptr;

static inline void split_plan(int T, int S, int* n_split, int* chunk) {
  if (T > 1) { *n_split = 1; *chunk = S; return; }
  *chunk = 64;
  *n_split = (S + 63) / 64;
}

static int launch_attn(const __nv_bfloat16* qkv, KvView kv, const int64_t* input_pos, const int32_t* ring_start,
                       float* work, __nv_bfloat16* y, int B, int T, int n_head, int hs, int cap, cudaStream_t st) {
  static const int env_pf = [] { const char* e = getenv("B2L_ATTN_PREFILL"); return e ? atoi(e) : 1; }();
  if (hs == 128 && T > 1 && env_pf) {   // tiled tensor-core kernel (B2L_ATTN_PREFILL=0: the per-query path below)
    static DynSmemCache smem_cache;
    if (int rc = ensure_dyn_smem(attn_prefill_kernel, PF_SMEM_BYTES, smem_cache)) return rc;
    attn_prefill_kernel<<<dim3(B * n_head, (T + PF_Q - 1) / PF_Q), 128, PF_SMEM_BYTES, st>>>(qkv, kv, input_pos, ring_start, y, T, n_head);
    B2L_LAUNCH_CHECK("attn_prefill_kernel");
    return 0;
  }
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

static inline size_t ws_partials_bytes(int B, int n_head, int head_size, int T, int S) {
  int n_split, chunk;
  split_plan(T, S, &n_split, &chunk);
  size_t a = (size_t)B * n_head * T * n_split * (head_size + 2) * sizeof(float);
  size_t f = (size_t)B * n_head * ((S + WS_CHUNK - 1) / WS_CHUNK) * (head_size + 2) * sizeof(float);
  return ((a > f ? a : f) + 15) & ~(size_t)15;
}

// [partials | int32 tickets[B*n_head]].  The caller zero-fills the buffer once when it
// allocates it; the fused decode kernel re-arms its tickets itself.
extern "C" size_t b2l_attn_workspace_bytes(int B, int n_head, int head_size, int T, int S) {
  return ws_partials_bytes(B, n_head, head_size, T, S) + (size_t)B * n_head * sizeof(int);
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
  if (T == 1 && head_size == 128 && !(flags & B2L_F_ROPE_ROWS) && !(flags & B2L_F_ATTN_UNFUSED)) {
    const int n_split = (S + FD_SUB - 1) / FD_SUB;
    int* tickets = reinterpret_cast<int*>(reinterpret_cast<char*>(work) + ws_partials_bytes(B, n_head, head_size, T, S));
    static DynSmemCache smem_cache;
    // B2L_ATTN_PRE (read once): sub-tiles requested before griddepcontrol.wait, 1 (default) or 2
    static const int env_pre = [] { const char* e = getenv("B2L_ATTN_PRE"); return e ? atoi(e) : 1; }();
    // B2L_ATTN_SMEM_MERGE (read once): 1 = all 32 key groups merge through shared memory, 0 = shuffles inside a warp first
    // (default 0: same-box A/B at position 1030, 1125.7 vs 1138.8 us per 7B token -- the extra block barrier costs more than the shuffles)
    static const int env_smem_merge = [] { const char* e = getenv("B2L_ATTN_SMEM_MERGE"); return e ? atoi(e) : 0; }();
    if (int rc = ensure_dyn_smem(attn_decode_fused_kernel, FD_SMEM_BYTES, smem_cache)) return rc;
    LaunchCfg lc(dim3(B * n_head, n_split), dim3(FD_WARPS * 32), FD_SMEM_BYTES, st, (flags & B2L_F_PDL) != 0);
    B2L_CUDA(cudaLaunchKernelEx(&lc.cfg, attn_decode_fused_kernel, (const __nv_bfloat16*)qkv, (__nv_bfloat16*)k_cache,
                                (__nv_bfloat16*)v_cache, (const float*)rope, input_pos, ring_start, (__nv_bfloat16*)y,
                                (float*)work, tickets, n_head, S, block_size, n_split, (unsigned long long*)g_attn_timeline, env_pre, env_smem_merge));
    return 0;
  }
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
