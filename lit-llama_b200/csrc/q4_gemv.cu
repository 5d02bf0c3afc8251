// Batch-1 decode kernel: fused [RMSNorm ->] int4 weight-only GEMV [-> residual | SwiGLU].
//
// Replaces, for gptq.int4 with one (scale, zero) per output row and a single activation row:
//   ColBlockQuantizedLinear.forward      lit_llama/quantization.py:413-423
//   linear_kernel_4bit_weight (Triton)   lit_llama/quantization.py:187-333
//   RMSNorm.forward                      lit_llama/model.py:270-277      (prologue)
//   x + h / silu(a) * b                  lit_llama/model.py:166-167, 252 (epilogue)
//
// Why not tcgen05 here: measured on B200 (tools/diag.py mma_rate / trace)
// a single thread issues a 128x16x16 tcgen05.mma every 45-80 cycles and every
// convert -> commit -> mbarrier round trip costs 300-500 cycles, so at one activation row the
// tensor-memory path is latency-bound at ~1/5 of HBM speed.  This kernel keeps the
// Blackwell data movement (TMA bulk copies into an mbarrier ring, PDL prefetch of the weights
// ahead of the dependency) and runs the tiny contraction warp-synchronously with
// mma.sync.m16n8k16 from registers: no TMEM hand-offs, no cross-CTA reduction.
//
// Work split: a persistent CTA owns 16-row blocks rb = cta, cta + grid, ... over the FULL K
// (so nothing is reduced across CTAs and the result is deterministic).  The 8 consumer warps
// split K inside a stage; their fp32 partials meet in shared memory, where the epilogue warp
// applies y = scale * (acc - (1024 + zero) * sum_lo(x) - (64 + zero) * sum_hi(x)) and the fused epilogue.
//
// Weight layout (b2l_q4_tile_mma): [N/16 row blocks][K/64 k blocks][32 lanes][16 B].  Word c of
// lane (g = lane/4, t = lane%4) holds the A fragment of k16 chunk c: nibble s (s < 4) is row
// g + 8*(s>>1), k = 64*kb + 16*c + 2*t + 8*(s&1); nibble s+4 is the same row at k+1.  The A
// registers of mma.m16n8k16 are built as fp16 pairs with one shift per word:
//   a0 = (w      & 0x000f000f) | 0x64006400   = 1024 + level       (row g,   k lower half)
//   a2 = (w      & 0x00f000f0) | 0x64006400   = 1024 + 16 * level  (row g,   k upper half)
//   a1, a3 = the same two masks on w >> 8                          (row g+8)
// and the activations of the upper half are fed as x / 16 (exact in fp16), so the accumulator holds
// sum(level * x) + 1024 * sum_lo(x) + 64 * sum_hi(x); the two sums are removed with the zero point.
// bf16 -> fp16 of the activations is exact inside the fp16 normal range; |x| > 65504 saturates.
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
};

// debug (tools/diag.py cta_times, only written when a trace buffer is attached): per CTA
// {x_ready ns, loop_done ns, %smid, stages streamed}
__device__ unsigned long long g_cta_dbg[1024 * 4];

// shared memory map
struct SmemLayout {
  uint32_t ring, xf, scratch, red, bars, total;
};
__host__ __device__ inline SmemLayout smem_layout(int nst, int K) {
  SmemLayout L;
  uint32_t o = 0;
  L.ring = o;    o += (uint32_t)nst * STAGE_BYTES;
  L.xf = o;      o += (uint32_t)(K / KB) * 128;       // B fragments: [k block][t (4)][32 B]
  L.scratch = o; o += 2 * NCW * RB * MAX_HALVES * 4;  // [buf][warp][half][row] fp32 partials
  L.red = o;     o += 128;                            // per-warp reduction scratch: sum of squares, sum(x) of either k class
  o = (o + 7u) & ~7u;
  L.bars = o;    o += 2 * MAX_STAGES * 8;
  L.total = (o + 127u) & ~127u;
  return L;
}

// MAXC = activation chunks (2048 elements each) a thread block caches in registers during the prologue:
// 6 covers K <= 12288 (every 7B/13B/30B layer), 12 covers K <= 24576 (65B mlp.c_proj, K = 22016).
template <int MAXC>
__global__ void __launch_bounds__(NTHREADS, 2) q4_gemv_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const SmemLayout L = smem_layout(p.nst, p.K);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_kb = p.K / KB;                                              // k blocks per 16-row block
  const int stages_per_unit = (n_kb + KBP_PER_STAGE - 1) / KBP_PER_STAGE;  // the last stage of a unit may be short
  // this CTA's contiguous range of 16-row blocks, processed as pairs and at most one single
  const int rb_lo = (int)(((long long)blockIdx.x * p.n_rb) / gridDim.x);
  const int rb_hi = (int)(((long long)(blockIdx.x + 1) * p.n_rb) / gridDim.x);
  const int n_units = (rb_hi - rb_lo + 1) / 2;
  const int total_stages = n_units * stages_per_unit;
  const uint32_t bar_full = sbase + L.bars, bar_empty = bar_full + MAX_STAGES * 8;

  if (tid == 0) tl_min(p.tl, 0);
  if (tid == 0) {
    for (int i = 0; i < p.nst; ++i) {
      mbar_init(bar_full + i * 8, 1);
      mbar_init(bar_empty + i * 8, NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == PRODUCER_WARP) {
    // ===================== TMA producer: the CTA's units, stage by stage =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 1;  // fresh barriers: waiting on parity 1 passes immediately
      int it = 0;
      for (int u = 0; u < n_units; ++u) {
        const int rb = rb_lo + 2 * u;
        const int halves = min(2, rb_hi - rb);
        const uint8_t* src = p.qwt + (size_t)rb * n_kb * KB_BYTES;
        for (int s = 0; s < stages_per_unit; ++s, ++it) {
          const int nkb = min(KBP_PER_STAGE, n_kb - s * KBP_PER_STAGE);
          const uint32_t bytes = (uint32_t)nkb * KB_BYTES;
          mbar_wait(bar_empty + slot * 8, phase);
          mbar_expect_tx(bar_full + slot * 8, bytes * halves);
          for (int h = 0; h < halves; ++h)
            tma_bulk_g2s(sbase + L.ring + slot * STAGE_BYTES + h * HALF_STAGE_BYTES,
                         src + ((size_t)h * n_kb + (size_t)s * KBP_PER_STAGE) * KB_BYTES, bytes, bar_full + slot * 8);
          if (p.tl != nullptr && blockIdx.x == 0 && it < 12) p.tl[40 + it] = globaltimer_ns();
          if (++slot == p.nst) { slot = 0; phase ^= 1; }
          if (it + 1 == min(total_stages, p.nst)) pdl_launch_dependents();  // ring full: next kernel may prefetch
        }
      }
      if (total_stages == 0) pdl_launch_dependents();
    }
  } else if (warp < NCW) {
    // ===================== consumer warps =====================
    float* red = reinterpret_cast<float*>(smem + L.red);  // [0..7] sum of squares, [8..15] sum of x, per warp
    // ---- activations: [RMSNorm], B-fragment order, sum(x)
    {
      const bool norm = (p.prologue == B2L_PRO_RMSNORM);
      constexpr int NT = NCW * 32;   // 256 threads, 8 elements each per pass
      uint4 xv[MAXC], gv[MAXC];
      // the RMSNorm scale is a weight: fetch it BEFORE waiting for the producing kernel (it comes from HBM,
      // behind the queued weight prefetch; after the wait it would sit on the critical path)
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int k = (c * NT + tid) * 8;
        gv[c] = make_uint4(0, 0, 0, 0);
        if (norm && k < p.K) gv[c] = *reinterpret_cast<const uint4*>(p.norm_scale + k);
      }
      pdl_wait();
      if (tid == 0) tl_max(p.tl, 1);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int k = (c * NT + tid) * 8;
        xv[c] = make_uint4(0, 0, 0, 0);
        if (k < p.K) xv[c] = *reinterpret_cast<const uint4*>(p.x + k);
      }
      const int nchunk = (p.K + NT * 8 - 1) / (NT * 8);  // warp-uniform: chunks that hold data
      float rinv = 1.f;
      if (p.tl != nullptr) {  // debug: when did the activation loads land?
        uint32_t sink = 0;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) sink |= xv[c].x;
        if (tid == 0 && sink != 0x12345678u) tl_max(p.tl, 5);
      }
      // bf16x2 arithmetic: one HMUL2 is the exactly-rounded bf16 product the reference computes
      // (bf16 * bf16 is exact in fp32, so rounding the fp32 product once == the packed multiply)
      if (norm) {
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          if (c < nchunk) {
            const uint32_t w[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
              const __nv_bfloat162 sq = __hmul2(v, v);
              const uint32_t su = *reinterpret_cast<const uint32_t*>(&sq);
              ss += __uint_as_float(su << 16) + __uint_as_float(su & 0xffff0000u);
            }
          }
        }
        ss = warp_sum(ss);
        if (lane == 0) red[warp] = ss;
        named_bar_sync(1, NT);
        if (tid == 0) tl_max(p.tl, 6);
        ss = 0.f;
#pragma unroll
        for (int w = 0; w < NCW; ++w) ss += red[w];
        rinv = rms_rinv(ss, p.K, p.eps);
      }
      const __nv_bfloat162 rinv2 = __float2bfloat162_rn(rinv);  // rinv is already a bf16 value
      float sx = 0.f;
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
          // 8 consecutive k = half of a k16 chunk: pair q (k = k0 + 2q, +1) is B register (half) of lane t = q
          // xf[k block][t][chunk c16 (4)][half (2)] u32.  The half is (tid & 1): a thread only ever sees one k class.
          // The tensor-core operand is fp16 (exact for a bf16 value in the fp16 range); the upper half of every
          // k16 chunk is pre-divided by 16 because its weights are unpacked as 1024 + 16*level (see the main loop).
          const int kb = k >> 6, c16 = (k >> 4) & 3, half = (k >> 3) & 1;
          const float pre = half ? 0.0625f : 1.0f;
          uint32_t* dst = reinterpret_cast<uint32_t*>(smem + L.xf + kb * 128);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float lo = __uint_as_float(w[q] << 16), hi = __uint_as_float(w[q] & 0xffff0000u);
            sx += lo + hi;
            dst[q * 8 + c16 * 2 + half] = pack_f16x2(lo * pre, hi * pre);
          }
        }
      }
      // even lanes hold lower-half (k % 16 < 8) sums, odd lanes upper-half sums
#pragma unroll
      for (int o = 16; o > 1; o >>= 1) sx += __shfl_xor_sync(0xffffffffu, sx, o);
      if (lane < 2) red[8 + 8 * lane + warp] = sx;   // the epilogue warp adds the 8 partials of each class in a fixed order
      named_bar_sync(3, NT + 32);          // releases the epilogue warp too: xf and the partial sums are ready
      if (tid == 0) {
        tl_max(p.tl, 2);
        if (p.tl != nullptr) {
          atomicMin(p.tl + 60, globaltimer_ns());
          if (blockIdx.x < 1024) g_cta_dbg[blockIdx.x * 4] = globaltimer_ns();
        }
      }
    }

    // ---- weights: stage -> registers -> mma.sync.  Warp w takes k-block positions w and w + 8 of a stage
    // and, for each, both 16-row halves of the unit (the B fragments are loaded once per position).
    const int t4 = lane & 3;
    // fp16 unpack: (w & 0x000f000f) | 0x64006400 = (1024 + level) pairs, (w & 0x00f000f0) | 0x64006400 =
    // (1024 + 16 * level) pairs -- two of the four A registers of a k16 chunk need no shift at all
    uint32_t kmask, kmask4, kmagic;
    asm volatile("mov.b32 %0, 0x000f000f;" : "=r"(kmask));
    asm volatile("mov.b32 %0, 0x00f000f0;" : "=r"(kmask4));
    asm volatile("mov.b32 %0, 0x64006400;" : "=r"(kmagic));
    int slot = 0;
    uint32_t phase = 0;
    float* scratch = reinterpret_cast<float*>(smem + L.scratch);
    const uint8_t* xf_lane = smem + L.xf + t4 * 32;
    int dbg_it = 0;
    for (int u = 0; u < n_units; ++u) {
      const int halves = min(2, rb_hi - (rb_lo + 2 * u));
      float acc[MAX_HALVES][2][4];
#pragma unroll
      for (int h = 0; h < MAX_HALVES; ++h)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[h][c][i] = 0.f;
      for (int s = 0; s < stages_per_unit; ++s) {
        const int nkb = min(KBP_PER_STAGE, n_kb - s * KBP_PER_STAGE);
        mbar_wait(bar_full + slot * 8, phase);
        if (p.tl != nullptr && blockIdx.x == 0 && tid == 0 && dbg_it < 12) p.tl[8 + 2 * dbg_it] = globaltimer_ns();
        const uint8_t* st_base = smem + L.ring + slot * STAGE_BYTES + lane * 16;
        if (nkb == KBP_PER_STAGE && !p.nocompute) {
          // the common case (full stage), branch-free for two halves and for one
          if (halves == MAX_HALVES) {
#pragma unroll
            for (int i = 0; i < KBP_PER_STAGE / NCW; ++i) {
              const int kbl = i * NCW + warp;
              const uint4* xp = reinterpret_cast<const uint4*>(xf_lane + (s * KBP_PER_STAGE + kbl) * 128);
              kblock_mma<MAX_HALVES>(acc, st_base + kbl * KB_BYTES, xp[0], xp[1], kmask, kmask4, kmagic);
            }
          } else {
#pragma unroll
            for (int i = 0; i < KBP_PER_STAGE / NCW; ++i) {
              const int kbl = i * NCW + warp;
              const uint4* xp = reinterpret_cast<const uint4*>(xf_lane + (s * KBP_PER_STAGE + kbl) * 128);
              kblock_mma<1>(acc, st_base + kbl * KB_BYTES, xp[0], xp[1], kmask, kmask4, kmagic);
            }
          }
        } else if (!p.nocompute) {
#pragma unroll
          for (int i = 0; i < KBP_PER_STAGE / NCW; ++i) {
            const int kbl = i * NCW + warp;  // k-block position inside the stage
            if (kbl < nkb) {
              const uint4* xp = reinterpret_cast<const uint4*>(xf_lane + (s * KBP_PER_STAGE + kbl) * 128);
              if (halves == MAX_HALVES) kblock_mma<MAX_HALVES>(acc, st_base + kbl * KB_BYTES, xp[0], xp[1], kmask, kmask4, kmagic);
              else kblock_mma<1>(acc, st_base + kbl * KB_BYTES, xp[0], xp[1], kmask, kmask4, kmagic);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + slot * 8);
        if (p.tl != nullptr && blockIdx.x == 0 && tid == 0 && dbg_it < 12) p.tl[9 + 2 * dbg_it] = globaltimer_ns();
        ++dbg_it;
        if (++slot == p.nst) { slot = 0; phase ^= 1; }
      }
      // column 0 of each 16x8 result: lanes with t == 0 hold rows g (acc[.][0]) and g + 8 (acc[.][2])
      const int buf = u & 1;
      named_bar_sync(4 + buf, NCW * 32 + 32);  // the epilogue warp has drained this scratch buffer (two units ago)
      if (t4 == 0) {
#pragma unroll
        for (int h = 0; h < MAX_HALVES; ++h) {
          float* dst = scratch + ((buf * NCW + warp) * MAX_HALVES + h) * RB + (lane >> 2);
          dst[0] = acc[h][0][0] + acc[h][1][0];
          dst[8] = acc[h][0][2] + acc[h][1][2];
        }
      }
      __syncwarp();
      named_bar_arrive(6 + buf, NCW * 32 + 32);  // partials of this unit are in the scratch buffer
    }
    if (tid == 0) {
      tl_max(p.tl, 3);
      if (p.tl != nullptr) {
        atomicMin(p.tl + 61, globaltimer_ns());
        if (blockIdx.x < 1024) {
          uint32_t smid;
          asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
          g_cta_dbg[blockIdx.x * 4 + 1] = globaltimer_ns();
          g_cta_dbg[blockIdx.x * 4 + 2] = smid;
          g_cta_dbg[blockIdx.x * 4 + 3] = (unsigned long long)total_stages;
        }
      }
    }
  } else {
    // ===================== epilogue warp: lane = row of the 32-row unit =====================
    pdl_wait();
    const float* red = reinterpret_cast<const float*>(smem + L.red);
    const float* scratch = reinterpret_cast<const float*>(smem + L.scratch);
    named_bar_sync(3, NCW * 32 + 32);
    // sum of the (normalised) activations over the lower (k % 16 < 8) and upper halves of the k16 chunks
    float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
    for (int w = 0; w < NCW; ++w) { sum_lo += red[8 + w]; sum_hi += red[16 + w]; }
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
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < NCW; ++w) t += scratch[((buf * NCW + w) * MAX_HALVES + half) * RB + row];  // fixed order: deterministic
      if (u + 2 < n_units) named_bar_arrive(4 + buf, NCW * 32 + 32);                                   // scratch buffer free again
      // t = sum q x + 1024 sum_lo + 64 sum_hi  (upper half: (1024 + 16 q) * x / 16)
      const float v = rbf(sc * ((t - (1024.0f + zero) * sum_lo) - (64.0f + zero) * sum_hi));
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

// ---------------------------------------------------------------- re-tiling for the mma.sync layout
__global__ void q4_tile_mma_kernel(const uint8_t* __restrict__ qw, uint32_t* __restrict__ out, int N, int K) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one output word
  const int n_kb = K / KB;
  const int n_rb = (N + RB - 1) / RB;
  const size_t total = (size_t)n_rb * n_kb * 32 * 4;
  if (idx >= total) return;
  const int c = idx & 3, lane = (idx >> 2) & 31;
  const size_t rest = idx >> 7;
  const int kb = (int)(rest % n_kb), rb = (int)(rest / n_kb);
  const int g = lane >> 2, t = lane & 3;
  uint32_t w = 0;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int ss = s & 3;
    const int row = rb * RB + g + 8 * (ss >> 1);
    const int k = kb * KB + 16 * c + 2 * t + 8 * (ss & 1) + (s >> 2);
    if (row < N) {
      const uint8_t b = qw[(size_t)(k >> 1) * N + row];
      w |= (uint32_t)((b >> ((k & 1) * 4)) & 0xF) << (4 * s);
    }
  }
  out[idx] = w;
}

__global__ void q4_untile_mma_kernel(const uint32_t* __restrict__ tiled, uint8_t* __restrict__ qw, int N, int K) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one packed byte [j][o]
  const size_t total = (size_t)(K / 2) * N;
  if (idx >= total) return;
  const int o = (int)(idx % N), j = (int)(idx / N);
  const int n_kb = K / KB;
  uint8_t b = 0;
#pragma unroll
  for (int nr = 0; nr < 2; ++nr) {
    const int k = 2 * j + nr;
    const int kb = k / KB, kl = k % KB, c = kl >> 4, k16 = kl & 15;
    const int hi8 = k16 >> 3, t = (k16 & 7) >> 1, odd = k16 & 1;
    const int rb = o / RB, rl = o % RB, g = rl & 7, r8 = rl >> 3;
    const int s = (r8 << 1 | hi8) + 4 * odd;
    const uint32_t w = tiled[(((size_t)rb * n_kb + kb) * 32 + (g * 4 + t)) * 4 + c];
    b |= (uint8_t)(((w >> (4 * s)) & 0xF) << (4 * nr));
  }
  qw[idx] = b;
}

}  // namespace q4mv
}  // namespace b2l

using namespace b2l;
using namespace b2l::q4mv;

extern "C" size_t b2l_q4_tiled_mma_bytes(int N, int K) {
  if (N <= 0 || K <= 0 || K % KB != 0) return 0;
  return (size_t)((N + RB - 1) / RB) * (K / KB) * KB_BYTES;
}

extern "C" int b2l_q4_tile_mma(const void* qw, void* qw_tiled, int N, int K, b2l_stream_t stream) {
  B2L_CHECK_ARG(qw && qw_tiled && N > 0 && K > 0, "b2l_q4_tile_mma: bad argument");
  B2L_CHECK_SUPPORTED(K % KB == 0, "b2l_q4_tile_mma: in_features %d must be a multiple of %d", K, KB);
  const size_t total = b2l_q4_tiled_mma_bytes(N, K) / 4;
  q4_tile_mma_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)qw, (uint32_t*)qw_tiled, N, K);
  B2L_LAUNCH_CHECK("q4_tile_mma_kernel");
  return 0;
}

extern "C" int b2l_q4_untile_mma(const void* qw_tiled, void* qw, int N, int K, b2l_stream_t stream) {
  B2L_CHECK_ARG(qw && qw_tiled && N > 0 && K > 0, "b2l_q4_untile_mma: bad argument");
  B2L_CHECK_SUPPORTED(K % KB == 0, "b2l_q4_untile_mma: in_features %d must be a multiple of %d", K, KB);
  const size_t total = (size_t)(K / 2) * N;
  q4_untile_mma_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint32_t*)qw_tiled, (uint8_t*)qw, N, K);
  B2L_LAUNCH_CHECK("q4_untile_mma_kernel");
  return 0;
}

extern "C" int b2l_debug_gemv_cta_times(void* out, int n_cta, b2l_stream_t stream) {
  B2L_CHECK_ARG(out && n_cta > 0 && n_cta <= 1024, "b2l_debug_gemv_cta_times: bad argument");
  B2L_CUDA(cudaMemcpyFromSymbolAsync(out, g_cta_dbg, (size_t)n_cta * 4 * sizeof(unsigned long long), 0, cudaMemcpyDeviceToDevice,
                                     (cudaStream_t)stream));
  return 0;
}

extern "C" int b2l_q4_gemv(const b2l_q4_linear_args* a, b2l_stream_t stream) {
  B2L_CHECK_ARG(a != nullptr, "b2l_q4_gemv: null args");
  B2L_CHECK_ARG(a->x && a->qw_tiled && a->scales && a->zeros && a->y, "b2l_q4_gemv: null pointer");
  B2L_CHECK_SUPPORTED(a->M == 1, "b2l_q4_gemv: M=%d (this kernel is the batch-1 path; use b2l_q4_linear_tc)", a->M);
  B2L_CHECK_SUPPORTED(a->K > 0 && a->K % KB == 0 && a->K <= 12 * NCW * 32 * 8, "b2l_q4_gemv: K=%d must be a multiple of %d and <= %d", a->K, KB, 12 * NCW * 32 * 8);
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
  p.tl = (unsigned long long*)a->trace;
  p.nocompute = (a->flags & B2L_F_DEBUG_NOCOMPUTE) ? 1 : 0;
  // tuning knobs (read once): B2L_GEMV_CTAS_PER_SM (default 2), B2L_GEMV_STAGES (ring depth cap)
  static const int env_cps = [] { const char* e = getenv("B2L_GEMV_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
  static const int env_nst = [] { const char* e = getenv("B2L_GEMV_STAGES"); return e ? atoi(e) : 0; }();
  // ring: as deep as fits two CTAs per SM
  const uint32_t fixed = smem_layout(0, a->K).total;
  int nst = (int)((110u * 1024u - fixed) / STAGE_BYTES);
  if (nst > MAX_STAGES) nst = MAX_STAGES;
  if (env_nst > 0 && nst > env_nst) nst = env_nst;
  if (nst < 2) nst = 2;
  p.nst = nst;
  const SmemLayout L = smem_layout(nst, a->K);
  const bool wide = a->K > 6 * NCW * 32 * 8;
  static DynSmemCache smem_cache[2];
  if (int rc = wide ? ensure_dyn_smem(q4_gemv_kernel<12>, L.total, smem_cache[1]) : ensure_dyn_smem(q4_gemv_kernel<6>, L.total, smem_cache[0]))
    return rc;
  int grid = a->split_k > 0 ? a->split_k : (env_cps > 0 ? env_cps : 2) * sm_count();  // split_k doubles as a grid override
  if (grid > p.n_rb) grid = p.n_rb;
  LaunchCfg lc(dim3(grid), dim3(NTHREADS), L.total, (cudaStream_t)stream, (a->flags & B2L_F_PDL) != 0, 1);
  if (wide) B2L_CUDA(cudaLaunchKernelEx(&lc.cfg, q4_gemv_kernel<12>, p));
  else B2L_CUDA(cudaLaunchKernelEx(&lc.cfg, q4_gemv_kernel<6>, p));
  return 0;
}
