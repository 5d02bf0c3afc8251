// Batch-1 decode kernel: fused [RMSNorm ->] int4 weight-only GEMV [-> residual | SwiGLU], integer tensor pipe.
//
// Replaces, for gptq.int4 with one (scale, zero) per output row and a single activation row:
//   ColBlockQuantizedLinear.forward      lit_llama/quantization.py:413-423
//   linear_kernel_4bit_weight (Triton)   lit_llama/quantization.py:187-333
//   RMSNorm.forward                      lit_llama/model.py:270-277      (prologue)
//   x + h / silu(a) * b                  lit_llama/model.py:166-167, 252 (epilogue)
//
// The contraction y[o] = scale[o] * sum_k (level[o,k] - zero[o]) * x[k] is evaluated in EXACT integer arithmetic:
//   * the activation row (after the RMSNorm prologue, i.e. the bf16 values the reference feeds its linear) is
//     scaled by a power of two 2^sh chosen from max|x| and rounded to X[k] (|X| < 2^22: 14 bits of dynamic range
//     below the largest element before an 8-bit significand loses a bit, and a rounding unit of 2^-22 max|x| from
//     there on), then split into three balanced base-256 digits X = sum_j d_j 256^j, d_j in [-128, 127];
//   * mma.sync.m16n8k32 (u8 x s8 -> s32, SASS IMMA.16832.U8.S8) has 8 result columns and a single activation row
//     needs one: digit plane j is column j, so all digits cost ONE MMA per 16 x 32 weight tile;
//   * a packed byte holds two levels.  It is fed to the tensor core UNMASKED as the operand of row g (value
//     level[g] + 16 level[g+8]) and with the low nibbles masked off as the operand of row g + 8 (16 level[g+8]):
//     one LOP3 per word, and the epilogue recovers row g = D[g] - D[g+8], row g + 8 = D[g+8] / 16;
//   * int32 accumulators cannot overflow for K <= 65536 (255 * 127 * K < 2^31); digits are recombined in int64 and
//     the zero point is removed with the exact sum of X: y = scale * 2^-sh * (sum level X - zero * sum X).
// Measured on B200 (tools/diag.py imma_rate / hmma_rate, profiles/r02_micro_imma_rate.txt): IMMA.16832 issues every
// 8.1 cycles per SM sub-partition like HMMA.16816, but consumes 512 levels behind 2 LOP3 where the f16 form consumed
// 256 behind 5 ALU ops: 10-12 issue cycles per 512 levels instead of 30.  Results no longer depend on the order of
// the K split, carry no 1024-bias cancellation, and have no fp16 range limit on the activations.
//
// Why not tcgen05 here: one thread issues a 128x16x16 tcgen05.mma every ~55 cycles (4 issuing warps: 15 cycles,
// profiles/r02_micro_issuers_gridflag.txt), but its A operand must first be expanded to 16-bit lanes by the same
// ALUs (>= 4 LOP3 per packed word) and handed over through tensor memory (300-500 cycles per hand-off, DESIGN.md
// section 3); the IMMA form needs 1 LOP3 per packed word and no hand-off.  tcgen05 carries the M > 8 shapes.
//
// Data movement is unchanged from round 1: a persistent CTA owns 16-row blocks over the FULL K, a producer warp
// streams 16 KB stages with TMA bulk copies into an mbarrier ring (issued before griddepcontrol.wait, so the
// weights of this linear stream while the previous kernel drains), 8 consumer warps split K inside a stage.
//
// Weight layout (b2l_q4_tile_i8): [N/16 row blocks][K/64 k blocks][32 lanes][16 B].  Lane (g = lane/4, t = lane%4),
// word 2c + j (c = k32 chunk of the k block, j = 0/1), byte i: low nibble = level[16 rb + g][k], high nibble =
// level[16 rb + g + 8][k], k = 64 kb + 32 c + 8 t + 4 j + i.  (The k order inside a chunk is a free choice as long as
// the activation digits use the same one; this one gives every prologue thread, which owns 8 consecutive k, both
// B registers of one lane.)
#include <cstdlib>

#include "q4_mma_common.cuh"

namespace b2l {
namespace q4mv {

struct Params {
  const __nv_bfloat16* x;
  const uint8_t* qwt;
  const void* scales; const void* zeros; int szdt;
  __nv_bfloat16* y;
  int N, K;              // N rows (padded to a multiple of 16 in the tiled weight), K % 64 == 0
  int n_rb;              // row blocks
  int prologue; const __nv_bfloat16* norm_scale; float eps;
  int epilogue; const __nv_bfloat16* res;
  int nst;               // ring stages
  unsigned long long* tl;  // debug timeline (nullptr = off)
  int nocompute;           // debug: consumers release every stage untouched (pure TMA streaming rate)
  // L2 prefetch hint (b2l_q4_linear_args::pf_ptr): byte ranges later launches stream; CTA c asks for the c-th slice
  const uint8_t* pf_ptr[B2L_PF_SEGMENTS];
  uint32_t pf_bytes[B2L_PF_SEGMENTS];
  int pf_mode;             // 0 off, 1 bulk prefetch by the producer lane, 2 / 4 per-line prefetch (128 B / 32 B apart)
  // strided form (KV-cache rows of a following attention launch; b2l_q4_linear_args::pf_kv)
  const uint8_t* pf_kv[2]; const long long* pf_rows; int pf_rows_max, pf_nseg, pf_row_bytes; unsigned long long pf_seg_stride;
  int evict_first;         // demand loads carry an L2 evict_first policy (a weight byte is read once per token)
};

constexpr uint32_t PF_CHUNK = 16384;   // bytes per bulk L2 prefetch instruction

// digit-plane stride in bytes: one byte per k, padded so that planes n and n + 1 fall into different bank halves
__host__ __device__ inline uint32_t plane_stride(int K) { return (uint32_t)K + ((K % 128 == 0) ? 64u : 0u); }

// shared memory map
struct SmemLayout {
  uint32_t ring, xf, zero, scratch, red, sxp, bars, total;
};
__host__ __device__ inline SmemLayout smem_layout(int nst, int K, int ndig) {
  SmemLayout L;
  uint32_t o = 0;
  L.ring = o;    o += (uint32_t)nst * STAGE_BYTES;
  L.xf = o;      o += (uint32_t)ndig * plane_stride(K);   // digit planes: [digit][k block][t (4)][16 B]
  L.zero = o;    o += 16;                                 // the B operand of the unused MMA columns
  L.scratch = o; o += 2 * NCW * MAX_HALVES * RB * 16;     // [buf][warp][half][row][digit (4)] int32 partials
  L.red = o;     o += 192;                                // reductions: float[8] sumsq, float[8] max, int64[8] sum X, int sh, float[8] max|g|
  L.sxp = o;     o += NCW * 32 * 4;                       // per-thread partial sums of X
  L.bars = o;    o += 2 * MAX_STAGES * 8;
  L.total = (o + 127u) & ~127u;
  return L;
}

// ---- L2 prefetch of weights a later launch reads (no shared memory, no completion: fire and forget)
__device__ __forceinline__ void l2_prefetch_bulk(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void l2_prefetch_line(const void* src) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
}
// this CTA's slice [lo, hi) of a segment of `bytes` bytes, cut at 128-byte lines
__device__ __forceinline__ void pf_slice(uint32_t bytes, uint32_t& lo, uint32_t& hi) {
  const uint32_t lines = (bytes + 127u) >> 7;
  lo = (uint32_t)(((unsigned long long)blockIdx.x * lines) / gridDim.x) << 7;
  hi = min((uint32_t)(((unsigned long long)(blockIdx.x + 1) * lines) / gridDim.x) << 7, bytes);
}
__device__ __forceinline__ void tma_bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(mbar), "l"(policy)
      : "memory");
}

// X (|X| < 2^30) -> word whose bytes are its balanced base-256 digits (byte 3 = signed top digit)
__device__ __forceinline__ uint32_t balanced_digits(int X) { return ((uint32_t)X + 0x00808080u) ^ 0x00808080u; }

// One k-block position (64 k = two k32 chunks) of NH 16-row halves: LDS.128 per half, 1 LOP3 + 1 IMMA per packed word pair.
template <int NH>
__device__ __forceinline__ void kblock_imma(int (&acc)[MAX_HALVES][2][4], const uint8_t* wbase, const uint4& xb) {
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const uint4 wv = *reinterpret_cast<const uint4*>(wbase + h * HALF_STAGE_BYTES);
    mma_u8s8_16832(acc[h][0], wv.x, wv.x & 0xf0f0f0f0u, wv.y, wv.y & 0xf0f0f0f0u, xb.x, xb.y);
    mma_u8s8_16832(acc[h][1], wv.z, wv.z & 0xf0f0f0f0u, wv.w, wv.w & 0xf0f0f0f0u, xb.z, xb.w);
  }
}

// One tile (16 rows x 64 k) into one accumulator set.
__device__ __forceinline__ void single_imma(int (&a)[2][4], const uint8_t* tile, const uint4& xb) {
  const uint4 wv = *reinterpret_cast<const uint4*>(tile);
  mma_u8s8_16832(a[0], wv.x, wv.x & 0xf0f0f0f0u, wv.y, wv.y & 0xf0f0f0f0u, xb.x, xb.y);
  mma_u8s8_16832(a[1], wv.z, wv.z & 0xf0f0f0f0u, wv.w, wv.w & 0xf0f0f0f0u, xb.z, xb.w);
}

// MAXC = activation chunks (2048 elements each) a thread block caches in registers during the prologue:
// 6 covers K <= 12288 (every 7B/13B/30B layer), 12 covers K <= 24576 (65B mlp.c_proj, K = 22016).
// NDIG = base-256 digits of the scaled activations: 3 (|X| < 2^22).
template <int MAXC, int NDIG>
__global__ void __launch_bounds__(NTHREADS, 2) q4_gemv_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const SmemLayout L = smem_layout(p.nst, p.K, NDIG);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_kb = p.K / KB;                                              // k blocks per 16-row block
  // A 16 KB stage holds 32 tiles of 512 B: 16 k-block positions of BOTH 16-row blocks of a pair, or 32 positions
  // of a single block (the ring slot is always used in full).  The last stage of a unit may be short.
  const int spu2 = (n_kb + KBP_PER_STAGE - 1) / KBP_PER_STAGE, spu1 = (n_kb + 2 * KBP_PER_STAGE - 1) / (2 * KBP_PER_STAGE);
  // this CTA's contiguous range of 16-row blocks, processed as pairs and at most one single
  const int rb_lo = (int)(((long long)blockIdx.x * p.n_rb) / gridDim.x);
  const int rb_hi = (int)(((long long)(blockIdx.x + 1) * p.n_rb) / gridDim.x);
  const int n_units = (rb_hi - rb_lo + 1) / 2;
  const int total_stages = ((rb_hi - rb_lo) / 2) * spu2 + ((rb_hi - rb_lo) & 1) * spu1;
  const uint32_t bar_full = sbase + L.bars, bar_empty = bar_full + MAX_STAGES * 8;
  const uint32_t PS = plane_stride(p.K);

  if (tid == 0) tl_min(p.tl, 0);
  if (tid == 0) {
    for (int i = 0; i < p.nst; ++i) {
      mbar_init(bar_full + i * 8, 1);
      mbar_init(bar_empty + i * 8, NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 4) reinterpret_cast<uint32_t*>(smem + L.zero)[tid] = 0u;
  __syncthreads();

  if (warp == PRODUCER_WARP) {
    // ===================== TMA producer: the CTA's units, stage by stage =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 1;  // fresh barriers: waiting on parity 1 passes immediately
      int it = 0;
      uint64_t policy = 0;
      if (p.evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      for (int u = 0; u < n_units; ++u) {
        const int rb = rb_lo + 2 * u;
        const int halves = min(2, rb_hi - rb);
        const uint8_t* src = p.qwt + (size_t)rb * n_kb * KB_BYTES;
        const int per_stage = halves == 2 ? KBP_PER_STAGE : 2 * KBP_PER_STAGE;
        for (int kb0 = 0; kb0 < n_kb; kb0 += per_stage, ++it) {
          const int nkb = min(per_stage, n_kb - kb0);
          const uint32_t bytes = (uint32_t)nkb * KB_BYTES;
          mbar_wait(bar_empty + slot * 8, phase);
          mbar_expect_tx(bar_full + slot * 8, bytes * halves);
          for (int h = 0; h < halves; ++h) {
            const uint32_t dst = sbase + L.ring + slot * STAGE_BYTES + h * HALF_STAGE_BYTES;
            const uint8_t* from = src + ((size_t)h * n_kb + kb0) * KB_BYTES;
            if (p.evict_first) tma_bulk_g2s_hint(dst, from, bytes, bar_full + slot * 8, policy);
            else tma_bulk_g2s(dst, from, bytes, bar_full + slot * 8);
          }
          if (++slot == p.nst) { slot = 0; phase ^= 1; }
          if (it + 1 == min(total_stages, p.nst)) {
            pdl_launch_dependents();  // ring full: next kernel may prefetch
            if (p.pf_mode == 1) {
              // the ring is full and this lane would now wait for a free slot: queue the L2 prefetch of this CTA's
              // slice of the weights the following launches read (HBM keeps streaming through the activation waits)
#pragma unroll
              for (int sgi = 0; sgi < B2L_PF_SEGMENTS; ++sgi) {
                if (p.pf_bytes[sgi] == 0) continue;
                uint32_t lo, hi;
                pf_slice(p.pf_bytes[sgi], lo, hi);
#pragma unroll 1
                for (uint32_t o = lo; o < hi; o += PF_CHUNK) l2_prefetch_bulk(p.pf_ptr[sgi] + o, min(PF_CHUNK, hi - o));
              }
            }
          }
        }
      }
      if (total_stages == 0) pdl_launch_dependents();
      if (p.pf_kv[0] != nullptr) {
        // KV-cache rows of the attention launch that follows the next linear: requested once this CTA's own ring is
        // full (behind its demand loads), 16 KB per instruction, round-robin over the grid.  The cache rows of old
        // positions do not depend on the current token; the attention kernel can only start its own loads when the
        // linear before it frees shared memory, so without this the cache streams from HBM on the critical path.
        long long rows = *p.pf_rows;
        rows = rows < 0 ? 0 : (rows > p.pf_rows_max ? p.pf_rows_max : rows);
        const uint32_t seg_bytes = (uint32_t)rows * (uint32_t)p.pf_row_bytes;
        const uint32_t cps = (seg_bytes + PF_CHUNK - 1) / PF_CHUNK;
        const uint32_t total = 2u * (uint32_t)p.pf_nseg * cps;
#pragma unroll 1
        for (uint32_t id = blockIdx.x; id < total; id += gridDim.x) {
          const uint32_t sg = id / cps, o = (id - sg * cps) * PF_CHUNK;
          const uint32_t which = sg >= (uint32_t)p.pf_nseg ? 1u : 0u;
          const uint8_t* base = p.pf_kv[which] + (size_t)(sg - which * p.pf_nseg) * p.pf_seg_stride;
          l2_prefetch_bulk(base + o, min(PF_CHUNK, seg_bytes - o));
        }
      }
    }
  } else if (warp < NCW) {
    // ===================== consumer warps =====================
    float* red = reinterpret_cast<float*>(smem + L.red);   // [0..7] sum of squares, [8..15] max, then int64[8] sum X, int sh
    int* red_sh = reinterpret_cast<int*>(smem + L.red + 128);
    // ---- activations: [RMSNorm], power-of-two scaling, balanced digits in B-fragment order, exact sum(X).
    // Every CTA converts the whole row (K values, 296 times per launch), and the launch cannot start its main loop
    // before this is done: the conversion is written for instruction count.  Per pair of elements: packed bf16
    // max / multiplies, ONE fma that both scales and rounds to an integer (x * 2^sh + 1.5 * 2^23: the integer sits
    // in the mantissa, |X| < 2^22 -- no F2I, which runs at a quarter of the fp32 rate), the balanced digits straight
    // from those bits with one add and one xor.
    {
      const bool norm = (p.prologue == B2L_PRO_RMSNORM);
      constexpr int NT = NCW * 32;   // 256 threads, 8 elements each per pass
      constexpr uint32_t MAGIC_BITS = 0x4B400000u;   // 1.5 * 2^23
      uint4 xv[MAXC], gv[MAXC];
      // the RMSNorm scale is a weight: fetch it (and reduce its max) BEFORE waiting for the producing kernel
      __nv_bfloat162 gmax2 = __float2bfloat162_rn(0.f);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int k = (c * NT + tid) * 8;
        gv[c] = make_uint4(0, 0, 0, 0);
        if (norm && k < p.K) gv[c] = *reinterpret_cast<const uint4*>(p.norm_scale + k);
      }
      if (norm) {
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          const uint32_t g[4] = {gv[c].x, gv[c].y, gv[c].z, gv[c].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) gmax2 = __hmax2(gmax2, __habs2(*reinterpret_cast<const __nv_bfloat162*>(&g[q])));
        }
        const float gw = warp_max(fmaxf(__low2float(gmax2), __high2float(gmax2)));
        if (lane == 0) red[36 + warp] = gw;      // read after the pass-1 barrier below
      }
      pdl_wait();
      if (tid == 0) tl_max(p.tl, 1);
      // written by the previous kernel: coherent loads (p.x is neither const __restrict__ nor read through ld.global.nc)
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int k = (c * NT + tid) * 8;
        xv[c] = make_uint4(0, 0, 0, 0);
        if (k < p.K) xv[c] = ld_coherent_u4(p.x + k);
      }
      const int nchunk = (p.K + NT * 8 - 1) / (NT * 8);  // warp-uniform: chunks that hold data
      // pass 1: sum of bf16-rounded squares (RMSNorm, model.py:274: one HMUL2 is the exactly-rounded bf16 product the
      // reference computes) and max |x| (with max |scale| it bounds the normalised values)
      float ss = 0.f;
      __nv_bfloat162 amax2 = __float2bfloat162_rn(0.f);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < nchunk) {
          const uint32_t w[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
            amax2 = __hmax2(amax2, __habs2(v));
            if (norm) {
              const __nv_bfloat162 sq = __hmul2(v, v);
              const uint32_t su = *reinterpret_cast<const uint32_t*>(&sq);
              ss += __uint_as_float(su << 16) + __uint_as_float(su & 0xffff0000u);
            }
          }
        }
      }
      float mx = fmaxf(__low2float(amax2), __high2float(amax2));
      ss = warp_sum(ss);
      mx = warp_max(mx);
      if (lane == 0) { red[warp] = ss; red[8 + warp] = mx; }
      named_bar_sync(1, NT);
      float gm = 0.f;
      ss = 0.f; mx = 0.f;
#pragma unroll
      for (int w = 0; w < NCW; ++w) { ss += red[w]; mx = fmaxf(mx, red[8 + w]); if (norm) gm = fmaxf(gm, red[36 + w]); }
      float rinv = 1.f;
      if (norm) {
        rinv = rms_rinv(ss, p.K, p.eps);
        mx = mx * gm * rinv * 1.01f;     // |bf16(g * bf16(x * rinv))| <= max|g| max|x| rinv (1 + 2^-8)^2
      }
      // 2^sh: the largest power of two with max|v| * 2^sh < 2^(8 NDIG - 2)
      const int e = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 127;   // mx < 2^(e + 1)
      int sh = (8 * NDIG - 3) - e;
      sh = max(-126, min(126, sh));
      const float scale = __uint_as_float((uint32_t)(sh + 127) << 23);
      const float magic = __uint_as_float(MAGIC_BITS);
      const __nv_bfloat162 rinv2 = __float2bfloat162_rn(rinv);  // rinv is already a bf16 value
      uint32_t sxu = 0;     // sum of this thread's X (<= 48 values below 2^22), modulo 2^32
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int k = (c * NT + tid) * 8;
        if (c < nchunk && k < p.K) {
          uint32_t w[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w};
          if (norm) {
            const uint32_t g[4] = {gv[c].x, gv[c].y, gv[c].z, gv[c].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
              const __nv_bfloat162 gg = *reinterpret_cast<const __nv_bfloat162*>(&g[q]);
              const __nv_bfloat162 y2 = __hmul2(gg, __hmul2(v, rinv2));  // bf16(scale * bf16(x * rinv)), model.py:276-277
              w[q] = *reinterpret_cast<const uint32_t*>(&y2);
            }
          }
          uint32_t xd[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            // x * 2^sh is exact in fp32 (8-bit significand, power-of-two scale): the fma rounds once, to nearest even,
            // and leaves MAGIC_BITS + X in the result's bit pattern
            const uint32_t b0 = __float_as_uint(__fmaf_rn(__uint_as_float(w[q] << 16), scale, magic));
            const uint32_t b1 = __float_as_uint(__fmaf_rn(__uint_as_float(w[q] & 0xffff0000u), scale, magic));
            sxu += b0 + b1;                                          // the 2 MAGIC_BITS per pair are removed below
            xd[2 * q] = (b0 + (0x00808080u - MAGIC_BITS)) ^ 0x00808080u;      // balanced_digits(X); byte 3 is unused
            xd[2 * q + 1] = (b1 + (0x00808080u - MAGIC_BITS)) ^ 0x00808080u;
          }
          sxu -= 8u * MAGIC_BITS;
          // 4 x 3 byte transposes: word (j, n) = digit n of elements 4j .. 4j+3  (B register j of lane t, column n)
          uint32_t dj[2][3];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t lo01 = __byte_perm(xd[4 * j], xd[4 * j + 1], 0x5140), hi01 = __byte_perm(xd[4 * j], xd[4 * j + 1], 0x7362);
            const uint32_t lo23 = __byte_perm(xd[4 * j + 2], xd[4 * j + 3], 0x5140), hi23 = __byte_perm(xd[4 * j + 2], xd[4 * j + 3], 0x7362);
            dj[j][0] = __byte_perm(lo01, lo23, 0x5410);
            dj[j][1] = __byte_perm(lo01, lo23, 0x7632);
            dj[j][2] = __byte_perm(hi01, hi23, 0x5410);
          }
          // k = 64 kb + 32 c32 + 8 t + (0..7): plane n, k block kb, lane slot t, words 2 c32, 2 c32 + 1
          uint8_t* dst = smem + L.xf + (k >> 6) * 64 + ((k >> 3) & 3) * 16 + ((k >> 5) & 1) * 8;
#pragma unroll
          for (int n = 0; n < NDIG; ++n) *reinterpret_cast<uint2*>(dst + n * PS) = make_uint2(dj[0][n], dj[1][n]);
        }
      }
      // exact sum of X over the row: every thread leaves its int32 partial (<= 48 values below 2^22) in shared memory
      // and goes on to the main loop; the epilogue warp, idle until the first unit is done, adds the 256 partials
      // in int64 (integers: the order does not matter)
      reinterpret_cast<int*>(smem + L.sxp)[tid] = (int)sxu;
      if (tid == 0) *red_sh = sh;
      named_bar_sync(3, NT + 32);          // releases the epilogue warp too: digit planes, sum X and sh are ready
      if (tid == 0) tl_max(p.tl, 2);
    }

    // ---- weights: stage -> registers -> mma.sync.  Warp w takes k-block positions w and w + 8 of a stage
    // and, for each, both 16-row halves of the unit (the B fragments are loaded once per position).
    // lanes 0..15 = MMA columns 0..3 = digit planes; lanes 16..31 (columns 4..7) read the zero block
    const int ncol = lane >> 2, t4 = lane & 3;
    const uint8_t* xf_lane = (ncol < NDIG) ? smem + L.xf + ncol * PS + t4 * 16 : smem + L.zero;
    const int xf_step = (ncol < NDIG) ? 64 : 0;
    int slot = 0;
    uint32_t phase = 0;
    int* scratch = reinterpret_cast<int*>(smem + L.scratch);
    for (int u = 0; u < n_units; ++u) {
      const int halves = min(2, rb_hi - (rb_lo + 2 * u));
      int acc[MAX_HALVES][2][4];
#pragma unroll
      for (int h = 0; h < MAX_HALVES; ++h)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[h][c][i] = 0;
      const int per_stage = halves == 2 ? KBP_PER_STAGE : 2 * KBP_PER_STAGE;
      for (int kb0 = 0; kb0 < n_kb; kb0 += per_stage) {
        const int nkb = min(per_stage, n_kb - kb0);
        mbar_wait(bar_full + slot * 8, phase);
        // warp w owns tiles w, w + 8, w + 16, w + 24 of the stage (same addresses for pairs and singles):
        // a pair: k-block positions w, w + 8 of block 0 and of block 1 (the digit fragments are shared);
        // a single: positions w, w + 8, w + 16, w + 24 of the one block, accumulated in both accumulator sets
        const uint8_t* st_base = smem + L.ring + slot * STAGE_BYTES + lane * 16;
        const uint8_t* xq = xf_lane + kb0 * xf_step;
        if (!p.nocompute) {
          if (halves == MAX_HALVES) {
            if (nkb == KBP_PER_STAGE) {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const int kbl = i * NCW + warp;
                kblock_imma<MAX_HALVES>(acc, st_base + kbl * KB_BYTES, *reinterpret_cast<const uint4*>(xq + kbl * xf_step));
              }
            } else {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const int kbl = i * NCW + warp;
                if (kbl < nkb) kblock_imma<MAX_HALVES>(acc, st_base + kbl * KB_BYTES, *reinterpret_cast<const uint4*>(xq + kbl * xf_step));
              }
            }
          } else {
            if (nkb == 2 * KBP_PER_STAGE) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int kbl = i * NCW + warp;
                single_imma(acc[i & 1], st_base + kbl * KB_BYTES, *reinterpret_cast<const uint4*>(xq + kbl * xf_step));
              }
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int kbl = i * NCW + warp;
                if (kbl < nkb) single_imma(acc[i & 1], st_base + kbl * KB_BYTES, *reinterpret_cast<const uint4*>(xq + kbl * xf_step));
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + slot * 8);
        if (++slot == p.nst) { slot = 0; phase ^= 1; }
      }
      // 16 x 8 result: lane (g, t) holds rows g (c0, c1) and g + 8 (c2, c3) of columns 2t, 2t + 1 = digits 2t, 2t + 1.
      // row g = D[g] - D[g+8] (the unmasked byte carried 16 * level[g+8] as well), row g + 8 = D[g+8] / 16 (exact)
      const int buf = u & 1;
      named_bar_sync(4 + buf, NCW * 32 + 32);  // the epilogue warp has drained this scratch buffer (two units ago)
      if (t4 < 2) {
        if (halves != MAX_HALVES) {   // a single block: its two accumulator sets hold different k-block positions
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[0][c][i] += acc[1][c][i];
        }
#pragma unroll
        for (int h = 0; h < MAX_HALVES; ++h) {
          const int c0 = acc[h][0][0] + acc[h][1][0], c1 = acc[h][0][1] + acc[h][1][1];
          const int c2 = acc[h][0][2] + acc[h][1][2], c3 = acc[h][0][3] + acc[h][1][3];
          int* dst = scratch + (((buf * NCW + warp) * MAX_HALVES + h) * RB + (lane >> 2)) * 4 + 2 * t4;
          *reinterpret_cast<int2*>(dst) = make_int2(c0 - c2, c1 - c3);
          *reinterpret_cast<int2*>(dst + 8 * 4) = make_int2(c2 >> 4, c3 >> 4);
        }
      }
      __syncwarp();
      named_bar_arrive(6 + buf, NCW * 32 + 32);  // partials of this unit are in the scratch buffer
    }
    if (tid == 0) tl_max(p.tl, 3);
  } else {
    // ===================== epilogue warp: lane = row of the 32-row unit =====================
    if (p.pf_mode >= 2) {   // per-line variant of the L2 prefetch (this warp idles until the first unit is reduced)
      const uint32_t step = p.pf_mode == 2 ? 128u : 32u;
#pragma unroll
      for (int sgi = 0; sgi < B2L_PF_SEGMENTS; ++sgi) {
        if (p.pf_bytes[sgi] == 0) continue;
        uint32_t lo, hi;
        pf_slice(p.pf_bytes[sgi], lo, hi);
#pragma unroll 1
        for (uint32_t o = lo + lane * step; o < hi; o += 32 * step) l2_prefetch_line(p.pf_ptr[sgi] + o);
      }
    }
    pdl_wait();
    const int* red_sh = reinterpret_cast<const int*>(smem + L.red + 128);
    const int* scratch = reinterpret_cast<const int*>(smem + L.scratch);
    named_bar_sync(3, NCW * 32 + 32);
    long long sum_x = 0;
    {
      const int4* sp = reinterpret_cast<const int4*>(smem + L.sxp) + lane * 2;   // 8 partials per lane
      const int4 s0 = sp[0], s1 = sp[1];
      sum_x = (long long)s0.x + s0.y + s0.z + s0.w + s1.x + s1.y + s1.z + s1.w;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum_x += __shfl_xor_sync(0xffffffffu, sum_x, o);
    }
    const double dsum_x = (double)sum_x;
    const int sh = *red_sh;
    const double inv_scale = __longlong_as_double((long long)(1023 - sh) << 52);   // 2^-sh
    // both scratch buffers start free
    if (n_units > 0) named_bar_arrive(4, NCW * 32 + 32);
    if (n_units > 1) named_bar_arrive(5, NCW * 32 + 32);
    for (int u = 0; u < n_units; ++u) {
      const int rb = rb_lo + 2 * u;
      const int halves = min(2, rb_hi - rb);
      const int buf = u & 1;
      const int half = lane >> 4, row = lane & 15;
      const bool active = half < halves;
      const int orow = (rb + half) * RB + row;             // row of the (interleaved) weight matrix
      const int o = min(orow, p.N - 1);
      const float sc = load_sz(p.scales, p.szdt, o);
      const float zero = load_sz(p.zeros, p.szdt, o);
      float resv = 0.f;
      if (p.epilogue == B2L_EPI_RESIDUAL && active && orow < p.N) resv = bf2f(p.res[orow]);
      named_bar_sync(6 + buf, NCW * 32 + 32);
      int d0 = 0, d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
      for (int w = 0; w < NCW; ++w) {   // integer sums: exact, independent of the order
        const int4 v = *reinterpret_cast<const int4*>(scratch + (((buf * NCW + w) * MAX_HALVES + half) * RB + row) * 4);
        d0 += v.x; d1 += v.y; d2 += v.z; d3 += v.w;
      }
      if (u + 2 < n_units) named_bar_arrive(4 + buf, NCW * 32 + 32);                                   // scratch buffer free again
      // sum_k level X = d0 + 256 d1 + 65536 d2 + 2^24 d3 (exact in int64, < 2^53)
      const long long tq = (long long)d0 + ((long long)d1 << 8) + ((long long)d2 << 16) + ((long long)d3 << 24);
      const float tf = (float)(((double)tq - (double)zero * dsum_x) * inv_scale);   // sum (level - zero) x, one rounding
      const float v = rbf(sc * tf);
      if (p.epilogue == B2L_EPI_SWIGLU) {
        // rows 0..7 of a 16-row block are c_fc1[o..o+7], rows 8..15 are c_fc2[o..o+7]
        const float b = __shfl_down_sync(0xffffffffu, v, 8);
        if (active && row < 8) {
          const float sl = rbf(v / (1.0f + expf(-v)));
          p.y[(rb + half) * 8 + row] = f2bf(sl * b);
        }
      } else if (active && orow < p.N) {
        p.y[orow] = f2bf(p.epilogue == B2L_EPI_RESIDUAL ? v + resv : v);
      }
    }
    if (lane == 0) tl_max(p.tl, 4);
  }
}

// ---------------------------------------------------------------- re-tiling for the int8-MMA layout
__global__ void q4_tile_i8_kernel(const uint8_t* __restrict__ qw, uint32_t* __restrict__ out, int N, int K) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one output word
  const int n_kb = K / KB;
  const int n_rb = (N + RB - 1) / RB;
  const size_t total = (size_t)n_rb * n_kb * 32 * 4;
  if (idx >= total) return;
  const int wd = idx & 3, lane = (idx >> 2) & 31;
  const size_t rest = idx >> 7;
  const int kb = (int)(rest % n_kb), rb = (int)(rest / n_kb);
  const int g = lane >> 2, t = lane & 3, c = wd >> 1, j = wd & 1;
  uint32_t w = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = kb * KB + 32 * c + 8 * t + 4 * j + i;
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const int row = rb * RB + g + 8 * hi;
      if (row < N) {
        const uint8_t b = qw[(size_t)(k >> 1) * N + row];
        w |= (uint32_t)((b >> ((k & 1) * 4)) & 0xF) << (8 * i + 4 * hi);
      }
    }
  }
  out[idx] = w;
}

__global__ void q4_untile_i8_kernel(const uint32_t* __restrict__ tiled, uint8_t* __restrict__ qw, int N, int K) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one packed byte [j][o]
  const size_t total = (size_t)(K / 2) * N;
  if (idx >= total) return;
  const int o = (int)(idx % N), jp = (int)(idx / N);
  const int n_kb = K / KB;
  uint8_t b = 0;
#pragma unroll
  for (int nr = 0; nr < 2; ++nr) {
    const int k = 2 * jp + nr;
    const int kb = k / KB, kl = k % KB, c = kl >> 5, kk = kl & 31;
    const int t = kk >> 3, j = (kk >> 2) & 1, i = kk & 3;
    const int rb = o / RB, rl = o % RB, g = rl & 7, hi = rl >> 3;
    const uint32_t w = tiled[(((size_t)rb * n_kb + kb) * 32 + (g * 4 + t)) * 4 + 2 * c + j];
    b |= (uint8_t)(((w >> (8 * i + 4 * hi)) & 0xF) << (4 * nr));
  }
  qw[idx] = b;
}

}  // namespace q4mv
}  // namespace b2l

using namespace b2l;
using namespace b2l::q4mv;

extern "C" size_t b2l_q4_tiled_i8_bytes(int N, int K) {
  if (N <= 0 || K <= 0 || K % KB != 0) return 0;
  return (size_t)((N + RB - 1) / RB) * (K / KB) * KB_BYTES;
}

extern "C" int b2l_q4_tile_i8(const void* qw, void* qw_tiled, int N, int K, b2l_stream_t stream) {
  B2L_CHECK_ARG(qw && qw_tiled && N > 0 && K > 0, "b2l_q4_tile_i8: bad argument");
  B2L_CHECK_SUPPORTED(K % KB == 0, "b2l_q4_tile_i8: in_features %d must be a multiple of %d", K, KB);
  const size_t total = b2l_q4_tiled_i8_bytes(N, K) / 4;
  q4_tile_i8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)qw, (uint32_t*)qw_tiled, N, K);
  B2L_LAUNCH_CHECK("q4_tile_i8_kernel");
  return 0;
}

extern "C" int b2l_q4_untile_i8(const void* qw_tiled, void* qw, int N, int K, b2l_stream_t stream) {
  B2L_CHECK_ARG(qw && qw_tiled && N > 0 && K > 0, "b2l_q4_untile_i8: bad argument");
  B2L_CHECK_SUPPORTED(K % KB == 0, "b2l_q4_untile_i8: in_features %d must be a multiple of %d", K, KB);
  const size_t total = (size_t)(K / 2) * N;
  q4_untile_i8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint32_t*)qw_tiled, (uint8_t*)qw, N, K);
  B2L_LAUNCH_CHECK("q4_untile_i8_kernel");
  return 0;
}

namespace {
constexpr int MAX_K = 12 * NCW * 32 * 8;   // 24576

template <int MAXC, int NDIG>
int launch_gemv(const Params& p0, int ctas_per_sm, int grid_override, bool pdl, cudaStream_t stream) {
  Params p = p0;
  // ring: as deep as fits `ctas_per_sm` CTAs per SM
  // B2L_GEMV_SMEM_KB: shared-memory budget of a CTA (default: 110 KB for two CTAs per SM, 224 KB for one)
  static const int env_kb = [] { const char* e = getenv("B2L_GEMV_SMEM_KB"); return e ? atoi(e) : 0; }();
  const uint32_t budget = (env_kb > 0 ? (uint32_t)env_kb : (ctas_per_sm >= 2 ? 110u : 224u)) * 1024u;
  const uint32_t fixed = smem_layout(0, p.K, NDIG).total;
  int nst = fixed + 2 * STAGE_BYTES <= budget ? (int)((budget - fixed) / STAGE_BYTES) : 0;
  if (nst > MAX_STAGES) nst = MAX_STAGES;
  static const int env_nst = [] { const char* e = getenv("B2L_GEMV_STAGES"); return e ? atoi(e) : 0; }();
  if (env_nst > 0 && nst > env_nst) nst = env_nst;
  if (nst < 2) {
    set_error("b2l_q4_gemv: K=%d does not leave room for the weight ring", p.K);
    return B2L_E_UNSUPPORTED;
  }
  p.nst = nst;
  const SmemLayout L = smem_layout(nst, p.K, NDIG);
  static DynSmemCache smem_cache;
  if (int rc = ensure_dyn_smem(q4_gemv_kernel<MAXC, NDIG>, L.total, smem_cache)) return rc;
  int grid = grid_override > 0 ? grid_override : ctas_per_sm * sm_count();
  if (grid > p.n_rb) grid = p.n_rb;
  LaunchCfg lc(dim3(grid), dim3(NTHREADS), L.total, stream, pdl, 1);
  B2L_CUDA(cudaLaunchKernelEx(&lc.cfg, q4_gemv_kernel<MAXC, NDIG>, p));
  return 0;
}
}  // namespace

extern "C" int b2l_q4_gemv(const b2l_q4_linear_args* a, b2l_stream_t stream) {
  B2L_CHECK_ARG(a != nullptr, "b2l_q4_gemv: null args");
  B2L_CHECK_ARG(a->x && a->qw_tiled && a->scales && a->zeros && a->y, "b2l_q4_gemv: null pointer");
  B2L_CHECK_SUPPORTED(a->M == 1, "b2l_q4_gemv: M=%d (this kernel is the batch-1 path; use b2l_q4_gemv_batch / b2l_q4_linear_tc)", a->M);
  B2L_CHECK_SUPPORTED(a->K > 0 && a->K % KB == 0 && a->K <= MAX_K, "b2l_q4_gemv: K=%d must be a multiple of %d and <= %d", a->K, KB, MAX_K);
  B2L_CHECK_ARG(a->N > 0, "b2l_q4_gemv: bad N");
  B2L_CHECK_ARG(((uintptr_t)a->x % 16 == 0) && ((uintptr_t)a->qw_tiled % 16 == 0), "b2l_q4_gemv: x / qw_tiled must be 16-byte aligned");
  B2L_CHECK_ARG(a->sz_dtype == B2L_BF16 || a->sz_dtype == B2L_F32, "b2l_q4_gemv: bad sz_dtype");
  if (a->prologue == B2L_PRO_RMSNORM)
    B2L_CHECK_ARG(a->norm_scale && ((uintptr_t)a->norm_scale % 16 == 0), "b2l_q4_gemv: RMSNorm prologue needs a 16-byte aligned scale");
  else
    B2L_CHECK_ARG(a->prologue == B2L_PRO_NONE, "b2l_q4_gemv: bad prologue %d", a->prologue);
  if (a->epilogue == B2L_EPI_RESIDUAL) B2L_CHECK_ARG(a->res != nullptr, "b2l_q4_gemv: RESIDUAL epilogue needs res");
  else if (a->epilogue == B2L_EPI_SWIGLU) B2L_CHECK_SUPPORTED(a->N % RB == 0, "b2l_q4_gemv: SWIGLU needs N %% 16 == 0");
  else B2L_CHECK_ARG(a->epilogue == B2L_EPI_STORE, "b2l_q4_gemv: bad epilogue %d", a->epilogue);

  Params p;
  p.x = (const __nv_bfloat16*)a->x;
  p.qwt = (const uint8_t*)a->qw_tiled;
  p.scales = a->scales; p.zeros = a->zeros; p.szdt = a->sz_dtype;
  p.y = (__nv_bfloat16*)a->y;
  p.N = a->N; p.K = a->K;
  p.n_rb = (a->N + RB - 1) / RB;
  p.prologue = a->prologue; p.norm_scale = (const __nv_bfloat16*)a->norm_scale; p.eps = a->eps;
  p.epilogue = a->epilogue; p.res = (const __nv_bfloat16*)a->res;
  p.nst = 0;
  p.tl = (unsigned long long*)a->trace;
  p.nocompute = (a->flags & B2L_F_DEBUG_NOCOMPUTE) ? 1 : 0;
  // L2 prefetch hint.  B2L_PF_MODE (read once): 0 ignores the hint, 1 (default) bulk prefetch, 2 / 4 per-line prefetch;
  // B2L_PF_EVICT=1 marks the demand loads evict_first
  static const int env_pf_mode = [] { const char* e = getenv("B2L_PF_MODE"); return e ? atoi(e) : 1; }();
  static const int env_evict = [] { const char* e = getenv("B2L_PF_EVICT"); return e ? atoi(e) : 0; }();
  p.pf_mode = 0;
  p.evict_first = env_evict;
  p.pf_kv[0] = p.pf_kv[1] = nullptr; p.pf_rows = nullptr; p.pf_rows_max = p.pf_nseg = p.pf_row_bytes = 0; p.pf_seg_stride = 0;
  if (a->pf_kv[0] != nullptr) {
    B2L_CHECK_ARG(a->pf_kv[1] != nullptr && a->pf_rows != nullptr && a->pf_rows_max > 0 && a->pf_nseg > 0 && a->pf_row_bytes > 0 &&
                      a->pf_row_bytes % 16 == 0 && a->pf_seg_stride % 16 == 0 && ((uintptr_t)a->pf_kv[0] % 16 == 0) &&
                      ((uintptr_t)a->pf_kv[1] % 16 == 0) && (unsigned long long)a->pf_rows_max * a->pf_row_bytes < (1ull << 31),
                  "b2l_q4_gemv: bad strided prefetch hint");
    p.pf_kv[0] = (const uint8_t*)a->pf_kv[0]; p.pf_kv[1] = (const uint8_t*)a->pf_kv[1];
    p.pf_rows = a->pf_rows; p.pf_rows_max = a->pf_rows_max; p.pf_nseg = a->pf_nseg; p.pf_row_bytes = a->pf_row_bytes;
    p.pf_seg_stride = a->pf_seg_stride;
  }
  for (int i = 0; i < B2L_PF_SEGMENTS; ++i) {
    p.pf_ptr[i] = (const uint8_t*)a->pf_ptr[i];
    p.pf_bytes[i] = 0;
    if (a->pf_ptr[i] != nullptr && a->pf_bytes[i] != 0) {
      B2L_CHECK_ARG(((uintptr_t)a->pf_ptr[i] % 16 == 0) && (a->pf_bytes[i] % 16 == 0) && a->pf_bytes[i] < (1ull << 31),
                    "b2l_q4_gemv: prefetch segment %d must be 16-byte aligned, a multiple of 16 and < 2 GiB", i);
      p.pf_bytes[i] = (uint32_t)a->pf_bytes[i];
      p.pf_mode = env_pf_mode;
    }
  }
  // tuning knob (read once): B2L_GEMV_CTAS_PER_SM (default 2; K > 16384 runs one CTA per SM with a deeper ring)
  static const int env_cps = [] { const char* e = getenv("B2L_GEMV_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
  const bool pdl = (a->flags & B2L_F_PDL) != 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = a->split_k;  // split_k doubles as a grid override
  // three digits (|X| < 2^22) everywhere: the prologue's fma conversion needs the integer inside a float mantissa
  if (a->K <= 12288) return launch_gemv<6, 3>(p, env_cps > 0 ? env_cps : 2, grid, pdl, st);
  if (a->K <= 16384) return launch_gemv<12, 3>(p, env_cps > 0 ? env_cps : 2, grid, pdl, st);
  return launch_gemv<12, 3>(p, 1, grid, pdl, st);
}
