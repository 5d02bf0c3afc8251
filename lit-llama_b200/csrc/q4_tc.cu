// Fused [RMSNorm ->] int4 weight-only linear [-> residual | SwiGLU] for M <= 16 rows
// (decode and short prefill) on the Blackwell tensor cores.
//
// Replaces, for gptq.int4 with one (scale, zero) per output row:
//   ColBlockQuantizedLinear.forward      lit_llama/quantization.py:413-423
//   linear_kernel_4bit_weight (Triton)   lit_llama/quantization.py:187-333
//   RMSNorm.forward                      lit_llama/model.py:270-277      (prologue)
//   x + h / silu(a) * b                  lit_llama/model.py:166-167, 252 (epilogue)
//
// Data flow per CTA (one 128-row output tile x one K range; the K ranges of a tile
// form a thread-block cluster and are reduced through distributed shared memory):
//
//   HBM --TMA bulk copy--> smem ring of packed slabs [128 rows][16 B = 32 nibbles]
//       --LDS.128, LOP3--> registers: bf16 pairs (128 + level), exact
//       --tcgen05.st-----> TMEM A operand (lane = output row, column = k pair)
//   x (bf16, RMSNorm'd on the fly) --> smem B operand (K-major core matrices)
//   tcgen05.mma.kind::f16  D[128 x 16] (TMEM, fp32) += A(TMEM) * B(smem)
//   tcgen05.ld --> y[o] = scale[o] * (acc - (128 + zero[o]) * sum_k x[k])
//
// The scale/zero are hoisted out of the K loop (exact algebra, fp32): the tensor core
// only ever sees the integers 128..143 and the bf16 activations.
//
// Weights do not depend on the previous kernel, so with programmatic dependent launch
// the TMA producer starts streaming before `griddepcontrol.wait`; only the x load waits.
#include "b2l_common.cuh"

namespace b2l {
namespace q4tc {

constexpr int TILE_N = 128;
constexpr int SLAB_K = 32;
constexpr int SLAB_BYTES = TILE_N * 16;  // 2048
constexpr int G = 2;                     // slabs per stage / per A buffer
constexpr int STAGE_BYTES = G * SLAB_BYTES;
constexpr int NAB = 3;                   // A buffers in TMEM
constexpr int A_COLS = G * 16;           // 32-bit columns per A buffer (K = 64 bf16)
constexpr int D_COL = NAB * A_COLS;      // accumulator columns start (96)
constexpr int TMEM_COLS = 128;
constexpr int NGROUPS = 2;               // convert warp groups, alternating stages
constexpr int NCONV = 128 * NGROUPS;     // convert threads (warps 0..7)
constexpr int PRODUCER_WARP = NCONV / 32;      // warp 8: TMA producer, TMEM alloc/dealloc
constexpr int MMA_WARP = PRODUCER_WARP + 1;    // warp 9: MMA issuer
constexpr int NTHREADS = NCONV + 64;
constexpr int MAX_M = 16;
constexpr int MAX_STAGES = 16;
constexpr int SMEM_BUDGET = 74 * 1024;   // three CTAs per SM

struct Params {
  const __nv_bfloat16* x; int ldx;
  const uint8_t* qwt;
  const void* scales; const void* zeros; int szdt;
  __nv_bfloat16* y; int ldy;
  int M, N, K;
  int prologue; const __nv_bfloat16* norm_scale; float eps;
  int epilogue; const __nv_bfloat16* res; int ldres;
  int S;          // cluster size (split-K)
  int nst_ring;   // ring stages
  int kcb;        // bytes per 8-k core-matrix column of the B operand (256, or 128 when rows 8..15 alias)
  int kseg_max;   // max K elements of one rank
  unsigned long long* trace;  // debug: clock64 stamps of CTA 0 (nullptr = off)
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(mbar)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor (sm_100 format):
//   core matrix = 8 rows x 16 bytes, contiguous (128 B)
//   LBO = byte distance between the two K halves of one K=16 MMA (next 8-k column)
//   SBO = byte distance between 8-row groups along N
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version for sm_100
  return d;                // layout_type = 0 (no swizzle), base_offset = 0
}
// kind::f16, A = B = bf16 (K-major), D = f32, M = 128, N = 16
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);

// 8 nibbles of one word -> 4 registers of bf16 pairs (128 + level):
// 0x4300 is bf16 128.0 whose ulp is 1, so OR-ing a 4-bit level into the mantissa is exact.
__device__ __forceinline__ void unpack_word(uint32_t w, uint32_t* out) {
  out[0] = (w & 0x000f000fu) | 0x43004300u;
  out[1] = ((w >> 4) & 0x000f000fu) | 0x43004300u;
  out[2] = ((w >> 8) & 0x000f000fu) | 0x43004300u;
  out[3] = ((w >> 12) & 0x000f000fu) | 0x43004300u;
}

#define B2L_TRACE(slot)                                                      \
  do {                                                                       \
    if (p.trace != nullptr && blockIdx.x == 0) p.trace[(slot)] = clock64(); \
  } while (0)

// ---------------------------------------------------------------- shared memory map
struct SmemLayout {
  uint32_t ring, xb, part, xsum, red, bars, tmem_slot, total;
};
__host__ __device__ inline SmemLayout smem_layout(int nst_ring, int kseg_max, int kcb, int M) {
  SmemLayout L;
  uint32_t o = 0;
  L.ring = o; o += (uint32_t)nst_ring * STAGE_BYTES;
  L.xb = o;   o += (uint32_t)(kseg_max / 8) * kcb;
  L.part = o; o += TILE_N * M * 4;  // [m][row] fp32 partials of this rank (DSMEM-read by rank 0)
  L.xsum = o; o += MAX_M * 4;
  L.red = o;  o += (NCONV / 32) * MAX_M * 4;  // per-warp partials
  o = (o + 7u) & ~7u;
  L.bars = o; o += (2 * MAX_STAGES + 2 * NAB + 2) * 8;
  L.tmem_slot = o; o += 8;
  L.total = (o + 127u) & ~127u;
  return L;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(NTHREADS, 3) q4_linear_tc_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const SmemLayout L = smem_layout(p.nst_ring, p.kseg_max, p.kcb, p.M);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.S;
  const int nt = blockIdx.x / S;
  const int rank = (S > 1) ? (int)cluster_ctarank() : 0;

  // K range of this rank, in slabs
  const int slabs_total = p.K / SLAB_K;
  const int sl_base = slabs_total / S, sl_rem = slabs_total % S;
  const int nslab = sl_base + (rank < sl_rem ? 1 : 0);
  const int slab0 = rank * sl_base + min(rank, sl_rem);
  const int nstages = (nslab + G - 1) / G;
  const int k0 = slab0 * SLAB_K, kseg = nslab * SLAB_K;

  const uint32_t bar_w_full = sbase + L.bars;
  const uint32_t bar_w_empty = bar_w_full + MAX_STAGES * 8;
  const uint32_t bar_a_full = bar_w_empty + MAX_STAGES * 8;
  const uint32_t bar_a_empty = bar_a_full + NAB * 8;
  const uint32_t bar_d_full = bar_a_empty + NAB * 8;
  const uint32_t bar_x_ready = bar_d_full + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + L.tmem_slot);

  if (tid == 0) B2L_TRACE(0);
  if (tid == 0) {
    for (int i = 0; i < p.nst_ring; ++i) {
      mbar_init(bar_w_full + i * 8, 1);
      mbar_init(bar_w_empty + i * 8, 4);
    }
    for (int i = 0; i < NAB; ++i) {
      mbar_init(bar_a_full + i * 8, 4);
      mbar_init(bar_a_empty + i * 8, 1);
    }
    mbar_init(bar_d_full, 1);
    mbar_init(bar_x_ready, 1);
    fence_barrier_init();
  }
  __syncthreads();

  // ===================== TMA producer (warp 8, lane 0): stream the packed slabs of this rank ==========
  // The first ring-full of stages is requested before anything else (TMEM allocation included):
  // the weights do not depend on the previous kernel, and under PDL this CTA may have to wait
  // for TMEM while the previous kernel still holds it.
  const uint8_t* w_src = p.qwt + ((size_t)nt * slabs_total + slab0) * SLAB_BYTES;
  const int n_first = min(nstages, p.nst_ring);
  if (warp == PRODUCER_WARP) {
    if (lane == 0) {
      for (int st = 0; st < n_first; ++st) {
        const int ns = min(G, nslab - st * G);
        const uint32_t bytes = (uint32_t)ns * SLAB_BYTES;
        mbar_expect_tx(bar_w_full + st * 8, bytes);
        tma_bulk_g2s(sbase + L.ring + st * STAGE_BYTES, w_src + (size_t)st * STAGE_BYTES, bytes, bar_w_full + st * 8);
        if (st < 20) B2L_TRACE(108 + st);
      }
      pdl_launch_dependents();  // the next kernel's CTAs may start prefetching their weights
    }
    __syncwarp();
  }
  if (warp == PRODUCER_WARP) tmem_alloc(sbase + L.tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) B2L_TRACE(1);

  if (warp == PRODUCER_WARP) {
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;  // second use of slot 0 waits for its first release
      for (int st = n_first; st < nstages; ++st) {
        mbar_wait(bar_w_empty + slot * 8, phase);
        const int ns = min(G, nslab - st * G);
        const uint32_t bytes = (uint32_t)ns * SLAB_BYTES;
        mbar_expect_tx(bar_w_full + slot * 8, bytes);
        tma_bulk_g2s(sbase + L.ring + slot * STAGE_BYTES, w_src + (size_t)st * STAGE_BYTES, bytes, bar_w_full + slot * 8);
        if (st < 20) B2L_TRACE(108 + st);
        if (++slot == p.nst_ring) { slot = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
    mbar_wait(bar_x_ready, 0);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + D_COL;
    const uint32_t kcb16 = (uint32_t)p.kcb >> 4;
    // descriptor of the first 8-k column; one K=16 MMA advances it by two columns
    uint64_t bdesc = make_b_desc(sbase + L.xb, p.kcb, p.kcb == 256 ? 128 : 0);
    uint32_t accumulate = 0;
    int ab = 0;
    uint32_t aphase = 0;
    for (int st = 0; st < nstages; ++st) {
      mbar_wait(bar_a_full + ab * 8, aphase);
      tc_fence_after();
      if (st < 20 && lane == 0) B2L_TRACE(64 + st);
      const int ns = min(G, nslab - st * G);
      if (elect_one()) {
        const uint32_t a_tmem = tmem_base + ab * A_COLS;
#pragma unroll
        for (int j = 0; j < 2 * G; ++j) {
          if (j < 2 * ns) {
            tc_mma_ts(d_tmem, a_tmem + j * 8, bdesc + (uint64_t)(2 * j * kcb16), IDESC, accumulate);
            accumulate = 1;
          }
        }
        tc_commit(bar_a_empty + ab * 8);  // arrives when the MMAs above have read A
      }
      __syncwarp();
      accumulate = 1;
      bdesc += (uint64_t)(4 * G * kcb16);
      if (st < 20 && lane == 0) B2L_TRACE(84 + st);
      if (++ab == NAB) { ab = 0; aphase ^= 1; }
    }
    if (elect_one()) tc_commit(bar_d_full);
    __syncwarp();
  } else if (warp < PRODUCER_WARP) {
    // ===================== convert warps (thread = output row of the tile; two groups alternate stages) =====
    const int group = warp >> 2;       // 0 or 1
    const int row = tid & (TILE_N - 1);
    // -- activations: wait for the producing kernel, normalise, lay out as the B operand
    pdl_wait();
    if (tid == 0) B2L_TRACE(2);
    float* xsum = reinterpret_cast<float*>(smem + L.xsum);
    float* red = reinterpret_cast<float*>(smem + L.red);  // [8 warps][MAX_M]
    {
      const int rows = (p.kcb == 256) ? 16 : 8;
      const int nkc = kseg / 8;
      if (p.M < rows) {
        for (int i = tid; i < nkc * rows; i += NCONV) {
          const int kc = i / rows, r = i % rows;
          if (r >= p.M)
            *reinterpret_cast<uint4*>(smem + L.xb + kc * p.kcb + (r >> 3) * 128 + (r & 7) * 16) = make_uint4(0, 0, 0, 0);
        }
      }
      for (int m = 0; m < p.M; ++m) {
        const __nv_bfloat16* xr = p.x + (size_t)m * p.ldx;
        const bool norm = (p.prologue == B2L_PRO_RMSNORM);
        // issue every global load of this row up front (one round trip): the whole row for the
        // sum of squares (4 chunks of 2048 elements in registers), this rank's segment, and the norm scale
        constexpr int MAXC = 4;
        uint4 full[MAXC];
        if (norm) {
#pragma unroll
          for (int c = 0; c < MAXC; ++c) {
            const int k = (c * NCONV + tid) * 8;
            full[c] = (k < p.K) ? *reinterpret_cast<const uint4*>(xr + k) : make_uint4(0, 0, 0, 0);
          }
        }
        // this rank's K segment: up to XC chunks of 8 elements per thread (kseg <= XC * NCONV * 8, host-checked)
        constexpr int XC = 2;
        uint4 u[XC], sc[XC];
#pragma unroll
        for (int c = 0; c < XC; ++c) {
          const int kk = (c * NCONV + tid) * 8;
          u[c] = make_uint4(0, 0, 0, 0);
          sc[c] = make_uint4(0, 0, 0, 0);
          if (kk < kseg) {
            u[c] = *reinterpret_cast<const uint4*>(xr + k0 + kk);
            if (norm) sc[c] = *reinterpret_cast<const uint4*>(p.norm_scale + k0 + kk);
          }
        }
        float rinv = 1.f;
        if (norm) {
          float ss = 0.f;
          if (p.K > MAXC * NCONV * 8) {  // very wide rows: remaining chunks the slow way
            for (int k = (MAXC * NCONV + tid) * 8; k < p.K; k += NCONV * 8) {
              uint4 t = *reinterpret_cast<const uint4*>(xr + k);
              const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float a = __uint_as_float(w[q] << 16), b = __uint_as_float(w[q] & 0xffff0000u);
                ss += rbf(a * a) + rbf(b * b);
              }
            }
          }
#pragma unroll
          for (int c = 0; c < MAXC; ++c) {
            const uint32_t w[4] = {full[c].x, full[c].y, full[c].z, full[c].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float a = __uint_as_float(w[q] << 16), b = __uint_as_float(w[q] & 0xffff0000u);
              ss += rbf(a * a) + rbf(b * b);
            }
          }
          ss = warp_sum(ss);
          if (lane == 0) red[warp * MAX_M + m] = ss;
          named_bar_sync(1, NCONV);
          ss = 0.f;
#pragma unroll
          for (int w = 0; w < NCONV / 32; ++w) ss += red[w * MAX_M + m];
          rinv = rms_rinv(ss, p.K, p.eps);
          named_bar_sync(1, NCONV);
        }
        float sx = 0.f;
#pragma unroll
        for (int c = 0; c < XC; ++c) {
          const int kk = (c * NCONV + tid) * 8;
          if (kk < kseg) {
            uint32_t w[4] = {u[c].x, u[c].y, u[c].z, u[c].w};
            if (norm) {
              const uint32_t g[4] = {sc[c].x, sc[c].y, sc[c].z, sc[c].w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float a = rms_apply(__uint_as_float(w[q] << 16), rinv, __uint_as_float(g[q] << 16));
                float b = rms_apply(__uint_as_float(w[q] & 0xffff0000u), rinv, __uint_as_float(g[q] & 0xffff0000u));
                sx += a + b;
                w[q] = (__float_as_uint(a) >> 16) | (__float_as_uint(b) & 0xffff0000u);
              }
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) sx += __uint_as_float(w[q] << 16) + __uint_as_float(w[q] & 0xffff0000u);
            }
            *reinterpret_cast<uint4*>(smem + L.xb + (kk / 8) * p.kcb + (m >> 3) * 128 + (m & 7) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        sx = warp_sum(sx);
        if (lane == 0) red[warp * MAX_M + m] = sx;
        named_bar_sync(1, NCONV);
        if (tid == 0) {
          float t = 0.f;
#pragma unroll
          for (int w = 0; w < NCONV / 32; ++w) t += red[w * MAX_M + m];
          xsum[m] = t;
        }
      }
      fence_proxy_async_smem();  // B operand written with generic stores, read by the tensor core
      named_bar_sync(1, NCONV);
      if (tid == 0) mbar_arrive(bar_x_ready);
      if (tid == 0) B2L_TRACE(3);
    }

    // -- weights: smem slab -> registers -> TMEM A operand.  Group g takes stages g, g+2, ...
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    int slot = group % p.nst_ring;
    uint32_t rphase = (group >= p.nst_ring) ? 1u : 0u;
    int ab = group;                   // NAB == 3: (st % 3) advances by 2 each step
    uint32_t aphase = 1;              // parity to wait on for "A buffer free"; flips when ab wraps
    for (int st = group; st < nstages; st += NGROUPS) {
      const int ns = min(G, nslab - st * G);
      mbar_wait(bar_w_full + slot * 8, rphase);
      if (tid == 0 && st < 20) B2L_TRACE(4 + st);
      uint4 wv[G];
#pragma unroll
      for (int s = 0; s < G; ++s)
        if (s < ns) wv[s] = *reinterpret_cast<const uint4*>(smem + L.ring + slot * STAGE_BYTES + s * SLAB_BYTES + row * 16);
      mbar_wait(bar_a_empty + ab * 8, aphase);
      tc_fence_after();
      if (tid == 0 && st < 20) B2L_TRACE(24 + st);
#pragma unroll
      for (int s = 0; s < G; ++s) {
        if (s < ns) {
          uint32_t r[16];
          unpack_word(wv[s].x, r + 0);
          unpack_word(wv[s].y, r + 4);
          unpack_word(wv[s].z, r + 8);
          unpack_word(wv[s].w, r + 12);
          tmem_st16(tmem_base + lane_base + ab * A_COLS + s * 16, r);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_w_empty + slot * 8);  // slab bytes are in registers/TMEM: slot may be refilled
        mbar_arrive(bar_a_full + ab * 8);
      }
      if (tid == 0 && st < 20) B2L_TRACE(44 + st);
      slot += NGROUPS;
      while (slot >= p.nst_ring) { slot -= p.nst_ring; rphase ^= 1; }
      ab += NGROUPS;
      if (ab >= NAB) { ab -= NAB; aphase ^= 1; }
    }

    // -- epilogue part 1 (group 0): accumulator -> scaled partial of this rank
    if (group == 0) {
      mbar_wait(bar_d_full, 0);
      tc_fence_after();
      if (tid == 0) B2L_TRACE(104);
      uint32_t acc[16];
      tmem_ld16(tmem_base + lane_base + D_COL, acc);
      const int o = min(nt * TILE_N + row, p.N - 1);  // padded rows of the last tile are never stored
      const float sc = load_sz(p.scales, p.szdt, o);
      const float zz = 128.0f + load_sz(p.zeros, p.szdt, o);
      float* part = reinterpret_cast<float*>(smem + L.part);
#pragma unroll
      for (int m = 0; m < MAX_M; ++m)
        if (m < p.M) part[m * TILE_N + row] = sc * (__uint_as_float(acc[m]) - zz * xsum[m]);
    }
  }

  // ===================== cross-rank reduction + epilogue (rank 0, group 0) =====================
  tc_fence_before();
  if (S > 1) cluster_sync_all(); else __syncthreads();
  if (tid == 0) B2L_TRACE(105);
  if (rank == 0 && warp < 4) {
    float tot[MAX_M];
    float* part = reinterpret_cast<float*>(smem + L.part);
    const uint32_t part_addr = sbase + L.part + tid * 4;
#pragma unroll
    for (int m = 0; m < MAX_M; ++m) {
      if (m < p.M) {
        float t = part[m * TILE_N + tid];
        for (int r = 1; r < S; ++r) t += ld_dsmem_f32(part_addr + m * TILE_N * 4, (uint32_t)r);  // fixed order
        tot[m] = t;
      }
    }
    const int o = nt * TILE_N + tid;
    if (p.epilogue == B2L_EPI_SWIGLU) {
      // rank 0's own partials were only read by the owning thread: reuse them for the exchange
      float* fin = part;
#pragma unroll
      for (int m = 0; m < MAX_M; ++m)
        if (m < p.M) fin[m * TILE_N + tid] = rbf(tot[m]);
      named_bar_sync(2, 128);
      if (tid < 64) {
        const int oo = nt * 64 + tid;
#pragma unroll
        for (int m = 0; m < MAX_M; ++m) {
          if (m < p.M) {
            const float a = fin[m * TILE_N + tid], b = fin[m * TILE_N + tid + 64];
            const float sl = rbf(a / (1.0f + expf(-a)));
            p.y[(size_t)m * p.ldy + oo] = f2bf(sl * b);
          }
        }
      }
    } else {
#pragma unroll
      for (int m = 0; m < MAX_M; ++m) {
        if (m < p.M && o < p.N) {
          float v = rbf(tot[m]);
          if (p.epilogue == B2L_EPI_RESIDUAL) v = v + bf2f(p.res[(size_t)m * p.ldres + o]);
          p.y[(size_t)m * p.ldy + o] = f2bf(v);
        }
      }
    }
  }
  if (tid == 0) B2L_TRACE(106);
  if (S > 1) cluster_sync_all(); else __syncthreads();
  if (warp == PRODUCER_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
  if (tid == 0) B2L_TRACE(107);
}

// ---------------------------------------------------------------- re-tiling
// nibble position s of word i  <->  k = 32*slab + 8*i + (s < 4 ? 2*s : 2*(s-4)+1)
__global__ void q4_tile_kernel(const uint8_t* __restrict__ qw, uint32_t* __restrict__ out, int N, int K) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int KS = K / SLAB_K;
  const int ntiles = (N + TILE_N - 1) / TILE_N;
  const size_t total = (size_t)ntiles * KS * TILE_N * 4;
  if (idx >= total) return;
  const int i = idx & 3;
  const int r = (idx >> 2) & (TILE_N - 1);
  const size_t rest = idx >> 9;
  const int ks = (int)(rest % KS), nt = (int)(rest / KS);
  const int o = nt * TILE_N + r;
  uint32_t w = 0;
  if (o < N) {
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int k = ks * SLAB_K + 8 * i + (s < 4 ? 2 * s : 2 * (s - 4) + 1);
      const uint8_t b = qw[(size_t)(k >> 1) * N + o];
      w |= (uint32_t)((b >> ((k & 1) * 4)) & 0xF) << (4 * s);
    }
  }
  out[idx] = w;
}

__global__ void q4_untile_kernel(const uint32_t* __restrict__ tiled, uint8_t* __restrict__ qw, int N, int K) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one packed byte [j][o]
  const size_t total = (size_t)(K / 2) * N;
  if (idx >= total) return;
  const int o = (int)(idx % N), j = (int)(idx / N);
  const int KS = K / SLAB_K;
  uint8_t b = 0;
#pragma unroll
  for (int nr = 0; nr < 2; ++nr) {
    const int k = 2 * j + nr;
    const int ks = k / SLAB_K, kl = k % SLAB_K, i = kl / 8, e = kl % 8;
    const int s = (e & 1) ? 4 + (e >> 1) : (e >> 1);
    const uint32_t w = tiled[(((size_t)(o / TILE_N) * KS + ks) * TILE_N + (o % TILE_N)) * 4 + i];
    b |= (uint8_t)(((w >> (4 * s)) & 0xF) << (4 * nr));
  }
  qw[idx] = b;
}

}  // namespace q4tc
}  // namespace b2l

using namespace b2l;
using namespace b2l::q4tc;

extern "C" size_t b2l_q4_tiled_bytes(int N, int K) {
  if (N <= 0 || K <= 0 || K % SLAB_K != 0) return 0;
  return (size_t)((N + TILE_N - 1) / TILE_N) * (K / SLAB_K) * SLAB_BYTES;
}

extern "C" int b2l_q4_tile(const void* qw, void* qw_tiled, int N, int K, b2l_stream_t stream) {
  B2L_CHECK_ARG(qw && qw_tiled && N > 0 && K > 0, "b2l_q4_tile: bad argument");
  B2L_CHECK_SUPPORTED(K % SLAB_K == 0, "b2l_q4_tile: in_features %d must be a multiple of %d", K, SLAB_K);
  const size_t total = b2l_q4_tiled_bytes(N, K) / 4;
  q4_tile_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)qw, (uint32_t*)qw_tiled, N, K);
  B2L_LAUNCH_CHECK("q4_tile_kernel");
  return 0;
}

extern "C" int b2l_q4_untile(const void* qw_tiled, void* qw, int N, int K, b2l_stream_t stream) {
  B2L_CHECK_ARG(qw && qw_tiled && N > 0 && K > 0, "b2l_q4_untile: bad argument");
  B2L_CHECK_SUPPORTED(K % SLAB_K == 0, "b2l_q4_untile: in_features %d must be a multiple of %d", K, SLAB_K);
  const size_t total = (size_t)(K / 2) * N;
  q4_untile_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint32_t*)qw_tiled, (uint8_t*)qw, N, K);
  B2L_LAUNCH_CHECK("q4_untile_kernel");
  return 0;
}

namespace b2l {
// Split-K (cluster size) choice: one wave of at most 3 CTAs per SM if possible; cost model =
// waves x (slabs per CTA + a fixed per-CTA cost worth ~24 slabs of prologue/epilogue latency).
int q4_pick_split(int n_tiles, int slabs_total) {
  const int slots = 3 * sm_count();
  int best = 1;
  long best_cost = -1;
  for (int S = 1; S <= 8; ++S) {
    if (slabs_total / S < 2) break;
    const int per = (slabs_total + S - 1) / S;
    if (per * q4tc::SLAB_K > 2 * q4tc::NCONV * 8) continue;  // two activation chunks per convert thread
    const long waves = ((long)n_tiles * S + slots - 1) / slots;
    const long cost = waves * (per + 24);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = S; }
  }
  return best_cost < 0 ? 8 : best;
}
}  // namespace b2l

extern "C" int b2l_q4_linear_tc(const b2l_q4_linear_args* a, b2l_stream_t stream) {
  B2L_CHECK_ARG(a != nullptr, "b2l_q4_linear_tc: null args");
  B2L_CHECK_ARG(a->x && a->qw_tiled && a->scales && a->zeros && a->y, "b2l_q4_linear_tc: null pointer");
  B2L_CHECK_SUPPORTED(a->M >= 1 && a->M <= MAX_M, "b2l_q4_linear_tc: M=%d outside 1..%d", a->M, MAX_M);
  B2L_CHECK_SUPPORTED(a->K > 0 && a->K % SLAB_K == 0, "b2l_q4_linear_tc: K=%d must be a multiple of %d", a->K, SLAB_K);
  B2L_CHECK_ARG(a->N > 0 && a->ldx >= a->K, "b2l_q4_linear_tc: bad N/ldx");
  B2L_CHECK_ARG(((uintptr_t)a->x % 16 == 0) && (a->ldx % 8 == 0) && ((uintptr_t)a->qw_tiled % 16 == 0),
                "b2l_q4_linear_tc: x / qw_tiled must be 16-byte aligned, ldx a multiple of 8");
  B2L_CHECK_ARG(a->sz_dtype == B2L_BF16 || a->sz_dtype == B2L_F32, "b2l_q4_linear_tc: bad sz_dtype");
  if (a->prologue == B2L_PRO_RMSNORM)
    B2L_CHECK_ARG(a->norm_scale && ((uintptr_t)a->norm_scale % 16 == 0), "b2l_q4_linear_tc: RMSNorm prologue needs a 16-byte aligned scale");
  else
    B2L_CHECK_ARG(a->prologue == B2L_PRO_NONE, "b2l_q4_linear_tc: bad prologue %d", a->prologue);
  if (a->epilogue == B2L_EPI_RESIDUAL) B2L_CHECK_ARG(a->res && a->ldres >= a->N, "b2l_q4_linear_tc: RESIDUAL epilogue needs res");
  else if (a->epilogue == B2L_EPI_SWIGLU) B2L_CHECK_SUPPORTED(a->N % TILE_N == 0, "b2l_q4_linear_tc: SWIGLU needs N %% 128 == 0");
  else B2L_CHECK_ARG(a->epilogue == B2L_EPI_STORE, "b2l_q4_linear_tc: bad epilogue %d", a->epilogue);

  const int n_tiles = (a->N + TILE_N - 1) / TILE_N;
  const int slabs_total = a->K / SLAB_K;
  int S = a->split_k > 0 ? a->split_k : q4_pick_split(n_tiles, slabs_total);
  B2L_CHECK_SUPPORTED(S >= 1 && S <= 8, "b2l_q4_linear_tc: split_k=%d must be in 1..8", S);
  while (S > 1 && slabs_total < 2 * S) --S;

  Params p;
  p.x = (const __nv_bfloat16*)a->x; p.ldx = a->ldx;
  p.qwt = (const uint8_t*)a->qw_tiled;
  p.scales = a->scales; p.zeros = a->zeros; p.szdt = a->sz_dtype;
  p.y = (__nv_bfloat16*)a->y; p.ldy = a->ldy;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.prologue = a->prologue; p.norm_scale = (const __nv_bfloat16*)a->norm_scale; p.eps = a->eps;
  p.epilogue = a->epilogue; p.res = (const __nv_bfloat16*)a->res; p.ldres = a->ldres;
  p.S = S;
  p.trace = (unsigned long long*)a->trace;
  const int max_slabs = (slabs_total + S - 1) / S;
  p.kseg_max = max_slabs * SLAB_K;
  p.kcb = (!(a->flags & B2L_F_NO_ALIAS_N) && a->M <= 8) ? 128 : 256;  // rows 8..15 of B alias rows 0..7
  B2L_CHECK_SUPPORTED(p.kseg_max <= 2 * NCONV * 8, "b2l_q4_linear_tc: K/split_k = %d > %d: raise split_k (K=%d, split_k=%d)",
                      p.kseg_max, 2 * NCONV * 8, a->K, S);
  const int stages_needed = (max_slabs + G - 1) / G;
  // ring depth: the whole K range in flight when it fits the per-CTA budget (3 CTAs per SM), >= 4 stages
  const uint32_t fixed = smem_layout(0, p.kseg_max, p.kcb, p.M).total;
  int ring = fixed < (uint32_t)SMEM_BUDGET ? (int)((SMEM_BUDGET - fixed) / STAGE_BYTES) : 0;
  if (ring < 4) ring = 4;
  if (ring > MAX_STAGES) ring = MAX_STAGES;
  p.nst_ring = stages_needed < ring ? stages_needed : ring;
  if (p.nst_ring < 1) p.nst_ring = 1;
  const SmemLayout L = smem_layout(p.nst_ring, p.kseg_max, p.kcb, p.M);
  B2L_CHECK_SUPPORTED(L.total <= 200 * 1024, "b2l_q4_linear_tc: shared memory %u B too large (K=%d, split_k=%d)", L.total, a->K, S);

  static DynSmemCache smem_cache;
  if (int rc = ensure_dyn_smem(q4_linear_tc_kernel, L.total, smem_cache)) return rc;
  LaunchCfg lc(dim3(n_tiles * S), dim3(NTHREADS), L.total, (cudaStream_t)stream, (a->flags & B2L_F_PDL) != 0, S);
  B2L_CUDA(cudaLaunchKernelEx(&lc.cfg, q4_linear_tc_kernel, p));
  return 0;
}
