// LLM.int8() linear for one activation row (decode), fused with its activation quantisation.
//
// Replaces Linear8bitLt.forward (lit_llama/quantization.py:38-77 + the bitsandbytes forward it
// inherits: MatMul8bitLt with has_fp16_weights=False, threshold=6.0):
//   a   = fp16(x)
//   out = { k : |a_k| >= threshold }                      (outlier columns, shared by the batch)
//   SCA = max_{k not in out} |a_k| ;  CA_k = round(a_k * 127 / SCA), 0 on outlier columns
//   y   = fp16( (CA . CB[o]) * SCA * SCB[o] / 127^2 ) + fp16( sum_{k in out} a_k * fp16(CB[o][k] * SCB[o] / 127) )
// The int8 x int8 -> int32 contraction runs on the tensor cores (mma.sync.m16n8k32.s8) straight
// from the TMA-staged tile: int8 needs no unpacking at all.  Same skeleton as q4_gemv.cu:
// persistent CTAs over 16-row blocks and the full K, TMA bulk copies into an mbarrier ring
// issued before the PDL dependency, deterministic in-CTA reduction.
//
// Weight layout (b2l_q8_tile): [N/16 row blocks][K/128 k blocks][4 k32 chunks][32 lanes][16 B];
// the 16 bytes of lane (g, t) are registers a0..a3 of m16n8k32: rows g, g+8 x k = 32c + 4t..+3
// and 32c + 16 + 4t..+3.
//
// parity: the arithmetic restates the published LLM.int8() algorithm (bitsandbytes is not in
// the reference tree and not installed): parity with the reference is unpinned.
#include <cuda_fp16.h>

#include "b2l_common.cuh"

namespace b2l {
namespace q8mv {

constexpr int RB = 16;
constexpr int KB = 128;                 // k per k block (4 IMMAs of k32)
constexpr int KB_BYTES = 2048;          // one (row block, k block)
constexpr int NCW = 8;
constexpr int KBP_PER_STAGE = 8;        // one k block per consumer warp per stage
constexpr int STAGE_BYTES = KBP_PER_STAGE * KB_BYTES;  // 16 KB
constexpr int MAX_STAGES = 6;
constexpr int PRODUCER_WARP = NCW;
constexpr int NTHREADS = (NCW + 2) * 32;
constexpr int MAX_K = 12288;

struct Params {
  const __nv_bfloat16* x;     // one row, bf16 [K]
  const uint8_t* wt;          // tiled CB
  const int8_t* cb;           // reference layout CB (N, K) row-major, for the outlier columns
  const float* scb;           // [N]
  const uint32_t* mask_in;    // optional precomputed outlier mask (K bits), shared by a batch; nullptr = derive from x
  __nv_bfloat16* y;           // [N]
  int N, K, n_rb, nst;
  float threshold;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
template <int ID> __device__ __forceinline__ void bar_sync_c(int n) { asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(n) : "memory"); }
template <int ID> __device__ __forceinline__ void bar_arrive_c(int n) { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(n) : "memory"); }

__device__ __forceinline__ void imma_16832(int (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

struct SmemLayout {
  uint32_t ring, xf, ah, mask, scratch, red, bars, total;
};
__host__ __device__ inline SmemLayout smem_layout(int nst, int K) {
  SmemLayout L;
  uint32_t o = 0;
  L.ring = o;    o += (uint32_t)nst * STAGE_BYTES;
  L.xf = o;      o += (uint32_t)(K / KB) * 128;   // int8 B fragments: [k block][t (4)][32 B]
  L.ah = o;      o += (uint32_t)K * 2;            // fp16 activations (outlier term)
  L.mask = o;    o += (uint32_t)((K + 31) / 32) * 4;
  o = (o + 15u) & ~15u;
  L.scratch = o; o += 2 * NCW * RB * 4;           // [buf][warp][row] int32 partials
  L.red = o;     o += 64;
  o = (o + 7u) & ~7u;
  L.bars = o;    o += 2 * MAX_STAGES * 8;
  L.total = (o + 127u) & ~127u;
  return L;
}

__global__ void __launch_bounds__(NTHREADS, 2) q8_gemv_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const SmemLayout L = smem_layout(p.nst, p.K);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_kb = p.K / KB;
  const int stages_per_rb = (n_kb + KBP_PER_STAGE - 1) / KBP_PER_STAGE;
  const int rb_lo = (int)(((long long)blockIdx.x * p.n_rb) / gridDim.x);
  const int rb_hi = (int)(((long long)(blockIdx.x + 1) * p.n_rb) / gridDim.x);
  const int n_units = rb_hi - rb_lo;
  const int total_stages = n_units * stages_per_rb;
  const uint32_t bar_full = sbase + L.bars, bar_empty = bar_full + MAX_STAGES * 8;

  if (tid == 0) {
    for (int i = 0; i < p.nst; ++i) {
      mbar_init(bar_full + i * 8, 1);
      mbar_init(bar_empty + i * 8, NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == PRODUCER_WARP) {
    if (lane == 0) {
      int slot = 0, it = 0;
      uint32_t phase = 1;
      for (int u = 0; u < n_units; ++u) {
        const uint8_t* src = p.wt + (size_t)(rb_lo + u) * n_kb * KB_BYTES;
        for (int s = 0; s < stages_per_rb; ++s, ++it) {
          const int nkb = min(KBP_PER_STAGE, n_kb - s * KBP_PER_STAGE);
          const uint32_t bytes = (uint32_t)nkb * KB_BYTES;
          mbar_wait(bar_empty + slot * 8, phase);
          mbar_expect_tx(bar_full + slot * 8, bytes);
          tma_bulk_g2s(sbase + L.ring + slot * STAGE_BYTES, src + (size_t)s * STAGE_BYTES, bytes, bar_full + slot * 8);
          if (++slot == p.nst) { slot = 0; phase ^= 1; }
          if (it + 1 == min(total_stages, p.nst)) pdl_launch_dependents();
        }
      }
      if (total_stages == 0) pdl_launch_dependents();
    }
  } else if (warp < NCW) {
    // ===================== consumer warps =====================
    pdl_wait();
    float* red = reinterpret_cast<float*>(smem + L.red);
    __half* ah = reinterpret_cast<__half*>(smem + L.ah);
    uint32_t* mask = reinterpret_cast<uint32_t*>(smem + L.mask);
    constexpr int NT = NCW * 32;
    constexpr int MAXC = MAX_K / (NT * 8);  // 6
    // ---- activations: fp16 copy, outlier mask, row-wise absmax over inliers, int8 B fragments
    for (int i = tid; i < (p.K + 31) / 32; i += NT) mask[i] = p.mask_in ? p.mask_in[i] : 0u;
    bar_sync_c<1>(NT);
    float av[MAXC][8];
    float amax = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int k = (c * NT + tid) * 8;
      if (k < p.K) {
        const uint4 u = *reinterpret_cast<const uint4*>(p.x + k);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
        uint32_t outl = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          av[c][2 * q] = __half2float(__float2half_rn(__uint_as_float(w[q] << 16)));
          av[c][2 * q + 1] = __half2float(__float2half_rn(__uint_as_float(w[q] & 0xffff0000u)));
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          ah[k + e] = __float2half_rn(av[c][e]);
          if (!p.mask_in && fabsf(av[c][e]) >= p.threshold) outl |= 1u << e;
        }
        if (outl) atomicOr(&mask[k >> 5], outl << (k & 31));
      }
    }
    bar_sync_c<1>(NT);
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int k = (c * NT + tid) * 8;
      if (k < p.K) {
        const uint32_t mb = (mask[k >> 5] >> (k & 31)) & 0xFFu;
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (!((mb >> e) & 1u)) amax = fmaxf(amax, fabsf(av[c][e]));
      }
    }
    amax = warp_max(amax);
    if (lane == 0) red[warp] = amax;
    bar_sync_c<1>(NT);
    float sca = 0.f;
#pragma unroll
    for (int w = 0; w < NCW; ++w) sca = fmaxf(sca, red[w]);
    const float qs = sca > 0.f ? 127.0f / sca : 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int k = (c * NT + tid) * 8;
      if (k < p.K) {
        const uint32_t mb = (mask[k >> 5] >> (k & 31)) & 0xFFu;
        uint32_t pk[2] = {0u, 0u};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          int qv = ((mb >> e) & 1u) ? 0 : __float2int_rn(av[c][e] * qs);
          qv = max(-127, min(127, qv));
          pk[e >> 2] |= (uint32_t)(qv & 0xFF) << (8 * (e & 3));
        }
        // k..k+3 -> lane t = (k % 16) / 4, half = (k % 32) / 16, chunk c32 = (k % 128) / 32; k+4..k+7 -> t + 1
        const int kb = k >> 7, c32 = (k >> 5) & 3, half = (k >> 4) & 1, t0 = (k >> 2) & 3;
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem + L.xf + kb * 128);
        dst[t0 * 8 + c32 * 2 + half] = pk[0];
        dst[(t0 + 1) * 8 + c32 * 2 + half] = pk[1];
      }
    }
    if (tid == 0) red[8] = sca;
    bar_sync_c<3>(NT + 32);  // xf, ah, mask, SCA ready (epilogue warp included)

    // ---- weights: stage -> registers -> mma.sync s8 (no unpacking)
    const int t4 = lane & 3;
    int slot = 0;
    uint32_t phase = 0;
    int* scratch = reinterpret_cast<int*>(smem + L.scratch);
    const uint8_t* xf_lane = smem + L.xf + t4 * 32;
    for (int u = 0; u < n_units; ++u) {
      int acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
      for (int s = 0; s < stages_per_rb; ++s) {
        const int nkb = min(KBP_PER_STAGE, n_kb - s * KBP_PER_STAGE);
        mbar_wait(bar_full + slot * 8, phase);
        if (warp < nkb) {
          const uint8_t* wb = smem + L.ring + slot * STAGE_BYTES + warp * KB_BYTES + lane * 16;
          const uint4* xp = reinterpret_cast<const uint4*>(xf_lane + (s * KBP_PER_STAGE + warp) * 128);
          const uint4 xa = xp[0], xb = xp[1];
          const uint32_t bb[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 a = *reinterpret_cast<const uint4*>(wb + c * 512);
            imma_16832(acc[c & 1], a, bb[2 * c], bb[2 * c + 1]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + slot * 8);
        if (++slot == p.nst) { slot = 0; phase ^= 1; }
      }
      const int buf = u & 1;
      if (buf) bar_sync_c<5>(NCW * 32 + 32); else bar_sync_c<4>(NCW * 32 + 32);
      if (t4 == 0) {
        int* dst = scratch + (buf * NCW + warp) * RB + (lane >> 2);
        dst[0] = acc[0][0] + acc[1][0];
        dst[8] = acc[0][2] + acc[1][2];
      }
      __syncwarp();
      if (buf) bar_arrive_c<7>(NCW * 32 + 32); else bar_arrive_c<6>(NCW * 32 + 32);
    }
  } else {
    // ===================== epilogue warp: lanes 0..15 = rows of the block =====================
    pdl_wait();
    const float* red = reinterpret_cast<const float*>(smem + L.red);
    const int* scratch = reinterpret_cast<const int*>(smem + L.scratch);
    const __half* ah = reinterpret_cast<const __half*>(smem + L.ah);
    const uint32_t* mask = reinterpret_cast<const uint32_t*>(smem + L.mask);
    bar_sync_c<3>(NCW * 32 + 32);
    const float sca = red[8];
    if (n_units > 0) bar_arrive_c<4>(NCW * 32 + 32);
    if (n_units > 1) bar_arrive_c<5>(NCW * 32 + 32);
    const int nwords = (p.K + 31) / 32;
    for (int u = 0; u < n_units; ++u) {
      const int buf = u & 1;
      const int row = lane & 15;
      const int orow = (rb_lo + u) * RB + row;
      const int o = min(orow, p.N - 1);
      const float scb = p.scb[o];
      // outlier term first (global loads overlap the consumers' work): fp16 weights, fp32 accumulate, k ascending
      float term = 0.f;
      bool any = false;
      const float wsc = scb / 127.0f;
      for (int wi = 0; wi < nwords; ++wi) {
        uint32_t mb = mask[wi];
        while (mb) {
          const int e = __ffs(mb) - 1;
          mb &= mb - 1;
          const int k = wi * 32 + e;
          const float wv = __half2float(__float2half_rn((float)p.cb[(size_t)o * p.K + k] * wsc));
          term = fmaf(__half2float(ah[k]), wv, term);
          any = true;
        }
      }
      if (buf) bar_sync_c<7>(NCW * 32 + 32); else bar_sync_c<6>(NCW * 32 + 32);
      int t = 0;
#pragma unroll
      for (int w = 0; w < NCW; ++w) t += scratch[(buf * NCW + w) * RB + row];
      if (u + 2 < n_units) { if (buf) bar_arrive_c<5>(NCW * 32 + 32); else bar_arrive_c<4>(NCW * 32 + 32); }
      float v = __half2float(__float2half_rn((float)t * (sca * scb * (1.0f / (127.0f * 127.0f)))));
      if (any) v = __half2float(__float2half_rn(v + __half2float(__float2half_rn(term))));
      if (lane < 16 && orow < p.N) p.y[orow] = f2bf(v);
    }
  }
}

// ---- re-tiling: CB (N, K) row-major int8 -> fragment order
__global__ void q8_tile_kernel(const int8_t* __restrict__ cb, uint32_t* __restrict__ out, int N, int K) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one output word
  const int n_kb = K / KB, n_rb = (N + RB - 1) / RB;
  const size_t total = (size_t)n_rb * n_kb * 4 * 32 * 4;
  if (idx >= total) return;
  const int r = idx & 3, lane = (idx >> 2) & 31, c = (idx >> 7) & 3;
  const size_t rest = idx >> 9;
  const int kb = (int)(rest % n_kb), rb = (int)(rest / n_kb);
  const int g = lane >> 2, t = lane & 3;
  const int row = rb * RB + g + 8 * (r & 1);
  const int k0 = kb * KB + 32 * c + 16 * (r >> 1) + 4 * t;
  uint32_t w = 0;
  if (row < N) {
#pragma unroll
    for (int e = 0; e < 4; ++e) w |= (uint32_t)(uint8_t)cb[(size_t)row * K + k0 + e] << (8 * e);
  }
  out[idx] = w;
}

}  // namespace q8mv
}  // namespace b2l

using namespace b2l;
using namespace b2l::q8mv;

extern "C" size_t b2l_q8_tiled_bytes(int N, int K) {
  if (N <= 0 || K <= 0 || K % KB != 0) return 0;
  return (size_t)((N + RB - 1) / RB) * (K / KB) * KB_BYTES;
}

extern "C" int b2l_q8_tile(const void* cb, void* tiled, int N, int K, b2l_stream_t stream) {
  B2L_CHECK_ARG(cb && tiled && N > 0 && K > 0, "b2l_q8_tile: bad argument");
  B2L_CHECK_SUPPORTED(K % KB == 0, "b2l_q8_tile: in_features %d must be a multiple of %d", K, KB);
  const size_t total = b2l_q8_tiled_bytes(N, K) / 4;
  q8_tile_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const int8_t*)cb, (uint32_t*)tiled, N, K);
  B2L_LAUNCH_CHECK("q8_tile_kernel");
  return 0;
}

extern "C" int b2l_q8_gemv(const void* x, const void* w_tiled, const void* cb, const void* scb, const void* outlier_mask, void* y,
                           int N, int K, float threshold, int flags, b2l_stream_t stream) {
  B2L_CHECK_ARG(x && w_tiled && cb && scb && y && N > 0, "b2l_q8_gemv: bad argument");
  B2L_CHECK_SUPPORTED(K > 0 && K % KB == 0 && K <= MAX_K, "b2l_q8_gemv: K=%d must be a multiple of %d and <= %d", K, KB, MAX_K);
  B2L_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)w_tiled % 16 == 0), "b2l_q8_gemv: x / w_tiled must be 16-byte aligned");
  Params p;
  p.x = (const __nv_bfloat16*)x; p.wt = (const uint8_t*)w_tiled; p.cb = (const int8_t*)cb; p.scb = (const float*)scb;
  p.mask_in = (const uint32_t*)outlier_mask; p.y = (__nv_bfloat16*)y;
  p.N = N; p.K = K; p.n_rb = (N + RB - 1) / RB; p.threshold = threshold;
  const uint32_t fixed = smem_layout(0, K).total;
  int nst = (int)((110u * 1024u - fixed) / STAGE_BYTES);
  if (nst > MAX_STAGES) nst = MAX_STAGES;
  if (nst < 2) nst = 2;
  p.nst = nst;
  const SmemLayout L = smem_layout(nst, K);
  static DynSmemCache smem_cache;
  if (int rc = ensure_dyn_smem(q8_gemv_kernel, L.total, smem_cache)) return rc;
  int grid = 2 * sm_count();
  if (grid > p.n_rb) grid = p.n_rb;
  LaunchCfg lc(dim3(grid), dim3(NTHREADS), L.total, (cudaStream_t)stream, (flags & B2L_F_PDL) != 0, 1);
  B2L_CUDA(cudaLaunchKernelEx(&lc.cfg, q8_gemv_kernel, p));
  return 0;
}

// outlier columns of a batch: bit k set iff any row has |fp16(x[m][k])| >= threshold
__global__ void q8_outlier_mask_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int M, int K, float threshold, uint32_t* __restrict__ mask) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  bool o = false;
  if (k < K)
    for (int m = 0; m < M; ++m) o |= fabsf(__half2float(__float2half_rn(bf2f(x[(size_t)m * ldx + k])))) >= threshold;
  const uint32_t b = __ballot_sync(0xffffffffu, o);
  if ((threadIdx.x & 31) == 0 && k < K) mask[k >> 5] = b;
}

extern "C" int b2l_q8_outlier_mask(const void* x, int ldx, int M, int K, float threshold, void* mask, b2l_stream_t stream) {
  B2L_CHECK_ARG(x && mask && M > 0 && K > 0 && K % 32 == 0, "b2l_q8_outlier_mask: bad argument");
  q8_outlier_mask_kernel<<<(K + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ldx, M, K, threshold, (uint32_t*)mask);
  B2L_LAUNCH_CHECK("q8_outlier_mask_kernel");
  return 0;
}
