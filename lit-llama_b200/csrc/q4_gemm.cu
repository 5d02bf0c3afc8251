// Prefill-shaped int4 weight-only GEMM on the 5th-generation tensor cores: y[M, N] = x[M, K] . dequant(W)[N, K]^T
// for M > 16 (prompt processing, no-cache forward, evaluation; BASELINE.json configs[3] prefill 8 x 512).
//
// Replaces ColBlockQuantizedLinear.forward (lit_llama/quantization.py:413-423) and the Triton kernel
// linear_kernel_4bit_weight (quantization.py:187-333: tiles up to 256 x 256, dequantise inside the tile loop) for
// 4-bit weights with one (scale, zero) per output row.  Round 1 sent this shape to a dequantise + library GEMM.
//
// One CTA computes a 256 (tokens) x 256 (output features) tile of y over the full K:
//   warp 0, one elected lane: tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 -> fp32), M = 128, N = 256, K = 16,
//       both operands from shared memory (K-major canonical core-matrix layout, no swizzle), two accumulators
//       (token rows 0..127 and 128..255) of 256 columns each = all 512 columns of tensor memory;
//       tcgen05.commit releases a stage to the producers and finally hands the accumulators to the epilogue;
//   warps 1..8 (256 threads = the 256 weight rows of the tile), per 64-wide k stage:
//       * the activation tile [256 tokens][64 k]: ONE tensor-map TMA copy (cp.async.bulk.tensor.3d) per stage.  x is
//         described to the TMA unit as a 3-D tensor (8 elements = 16 B | M rows, stride ldx | K/8 chunks, stride 16 B),
//         so a box of 8 x 256 x 8 lands in shared memory as [k chunk][row][16 B] -- exactly the no-swizzle
//         K-major core-matrix order tcgen05 reads; rows beyond M are zero-filled by the TMA unit;
//       * two LDG.128 of packed levels per thread (load-time tiling b2l_q4_tile: a row's 32 levels of a k slab in
//         16 bytes, nibble order chosen so that (w >> 4s) & 0x000f000f | 0x43004300 IS the bf16 pair
//         (128 + level[k], 128 + level[k+1])), dequantised with the reference's own rounding:
//         level = v - 128 (exact), level - zero (bf16), * scale (bf16) -- bit-identical to get_weight
//         (quantization.py:392-411), so the GEMM sees exactly the matrix the reference's dense branch multiplies;
//       * 16-byte st.shared of 8 consecutive k of a row = one row of a core matrix; fence.proxy.async; mbarrier arrive;
//   afterwards the same 8 warps are the epilogue: tcgen05.ld 32x32b (a warp reaches the TMEM lanes of its
//   warp_id % 4 quarter), fp32 -> bf16, 32-byte stores.
// The dequantisation costs the 8 producer warps ~400 issue cycles per stage against 1024 tensor-pipe cycles
// (8 MMAs of 128 x 256 x 16), so the tensor pipe is the bound; a 4-deep ring (64 KB per stage) hides the loads.
#include <cuda.h>   // CUtensorMap and its enums only: the encoder is fetched with cudaGetDriverEntryPoint (no libcuda link)

#include <cstdlib>

#include "b2l_common.cuh"

namespace b2l {
namespace q4gm {

// NACC = accumulators = 128-token row groups per CTA: 2 (256 x 256 tile) for large problems; 1 (128 x 256) when the
// 256-row tiling would leave the last wave of CTAs mostly empty (e.g. N = 5120, M = 4096: 320 tiles on 148 SMs)
constexpr int BN = 256, BK = 64;
constexpr int B_BYTES = BN * BK * 2;           // 32 KB: [8 k-columns][256 rows][16 B]
constexpr int NPROD = 256;                     // producer threads = weight rows of the tile
constexpr int NTHREADS = 32 + NPROD;
constexpr int LBO_B = BN * 16;                 // bytes between adjacent 8-k columns of the weight operand
constexpr int SBO = 128;                       // bytes between adjacent 8-row groups
template <int NACC> struct Cfg {
  static constexpr int BM = 128 * NACC;
  static constexpr int A_BYTES = BM * BK * 2;  // [8 k-columns][BM rows][16 B]
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NSTAGE = NACC == 2 ? 3 : 4;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 256;
  static constexpr int LBO_A = BM * 16;
};

struct Params {
  CUtensorMap xmap;        // 3-D view of x (see above); must stay the first member (64-byte alignment)
  const __nv_bfloat16* x; int ldx;
  const uint8_t* qwt;      // b2l_q4_tile layout: [N/128 tiles][K/32 slabs][128 rows][16 B]
  const void* scales; const void* zeros; int szdt;
  __nv_bfloat16* y; int ldy;
  int M, N, K;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major, no-swizzle shared-memory matrix descriptor (sm_100): core matrix = 8 rows x 16 B, contiguous
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::f16: D = f32, A = B = bf16 (K-major), M = 128, N = 256
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t mbar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

template <int NACC>
__global__ void __launch_bounds__(NTHREADS, 1) q4_gemm_kernel(const __grid_constant__ Params p) {
  constexpr int BM = Cfg<NACC>::BM, A_BYTES = Cfg<NACC>::A_BYTES, STAGE_BYTES = Cfg<NACC>::STAGE_BYTES, NSTAGE = Cfg<NACC>::NSTAGE;
  constexpr int LBO_A = Cfg<NACC>::LBO_A;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + NSTAGE * STAGE_BYTES;      // full[NSTAGE], empty[NSTAGE], accum
  const uint32_t bar_full = bars, bar_empty = bars + NSTAGE * 8, bar_acc = bars + 2 * NSTAGE * 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + NSTAGE * STAGE_BYTES + (2 * NSTAGE + 1) * 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int n_kt = p.K / BK;

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(bar_full + i * 8, NPROD / 32 + 1);   // one elected arrival per producer warp + the TMA issuer's expect_tx
      mbar_init(bar_empty + i * 8, 1);           // tcgen05.commit
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256 * NACC) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      for (int kt = 0; kt < n_kt; ++kt) {
        const int st = kt % NSTAGE;
        mbar_wait(bar_full + st * 8, (uint32_t)(kt / NSTAGE) & 1u);
        tc_fence_after();
        const uint32_t a_base = sbase + st * STAGE_BYTES, b_base = a_base + A_BYTES;
#pragma unroll
        for (int j = 0; j < BK / 16; ++j) {
          const uint64_t bd = make_desc(b_base + j * 2 * LBO_B, LBO_B, SBO);
          // token rows 0..127 -> accumulator columns 0..255, rows 128..255 -> columns 256..511
#pragma unroll
          for (int h = 0; h < NACC; ++h)
            tc_mma_ss(tmem + 256 * h, make_desc(a_base + h * 128 * 16 + j * 2 * LBO_A, LBO_A, SBO), bd, IDESC, (kt > 0 || j > 0) ? 1u : 0u);
        }
        tc_commit(bar_empty + st * 8);   // the stage may be refilled once these MMAs have read it
      }
      tc_commit(bar_acc);                // accumulators complete
    }
  } else {
    // ===================== producers: activations (cp.async) + dequantised weights =====================
    const int pt = tid - 32;                       // 0..255 = weight row of the tile
    const int row_n = n0 + pt;
    const int n_slab = p.K / 32;
    const bool row_ok = row_n < ((p.N + 127) / 128) * 128;     // rows inside the (128-padded) tiling
    const uint8_t* wrow = p.qwt + ((size_t)(row_n >> 7) * n_slab * 128 + (row_n & 127)) * 16;   // + slab * 2048
    float sc_f = 0.f, z_f = 0.f;
    if (row_n < p.N) { sc_f = load_sz(p.scales, p.szdt, row_n); z_f = load_sz(p.zeros, p.szdt, row_n); }
    const __nv_bfloat162 sc2 = __float2bfloat162_rn(sc_f), z2 = __float2bfloat162_rn(z_f);
    const __nv_bfloat162 c128 = __float2bfloat162_rn(128.f);
    uint4 wq[2];
    auto load_w = [&](int kt) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        wq[h] = make_uint4(0, 0, 0, 0);
        if (row_ok) wq[h] = __ldg(reinterpret_cast<const uint4*>(wrow + (size_t)(kt * 2 + h) * 2048));
      }
    };
    load_w(0);
    for (int kt = 0; kt < n_kt; ++kt) {
      const int st = kt % NSTAGE;
      if (kt >= NSTAGE) mbar_wait(bar_empty + st * 8, (uint32_t)(kt / NSTAGE - 1) & 1u);
      const uint32_t a_base = sbase + st * STAGE_BYTES, b_base = a_base + A_BYTES;
      // ---- activations: one tensor-map TMA copy of the whole [8 k chunks][256 rows][16 B] tile
      if (pt == 0) {
        mbar_expect_tx(bar_full + st * 8, A_BYTES);
        tma_load_3d(a_base, &p.xmap, 0, m0, kt * (BK / 8), bar_full + st * 8);
      }
      // ---- weights of this stage (loaded one stage ahead), then the next stage's loads
      const uint4 w0 = wq[0], w1 = wq[1];
      if (kt + 1 < n_kt) load_w(kt + 1);
      const uint32_t ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int c = 0; c < 8; ++c) {       // word c = 8 consecutive k = one 16-byte core-matrix row
        uint32_t o[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const uint32_t v = ((ww[c] >> (4 * s)) & 0x000f000fu) | 0x43004300u;          // (128 + level) pair, exact
          __nv_bfloat162 t = __hsub2(*reinterpret_cast<const __nv_bfloat162*>(&v), c128);   // level, exact
          t = __hmul2(__hsub2(t, z2), sc2);      // (level - zero) * scale with the reference's bf16 roundings
          o[s] = *reinterpret_cast<const uint32_t*>(&t);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(b_base + c * LBO_B + pt * 16), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + st * 8);
    }
    // ===================== epilogue: TMEM -> bf16 -> y =====================
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int quarter = warp & 3;                 // TMEM lanes 32 * quarter .. + 31 are reachable from this warp
    // two accumulators: warps 1..4 drain accumulator 0 (tokens 0..127), warps 5..8 accumulator 1, all 256 columns each;
    // one accumulator: warps 1..4 drain columns 0..127, warps 5..8 columns 128..255
    const int grp = (warp - 1) >> 2;
    const int chalf = NACC == 2 ? grp : 0;
    const int cb_lo = NACC == 2 ? 0 : grp * (BN / 32), cb_hi = NACC == 2 ? BN / 16 : cb_lo + BN / 32;
    const int m = m0 + chalf * 128 + quarter * 32 + lane;
#pragma unroll 1
    for (int cb = cb_lo; cb < cb_hi; ++cb) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(chalf * 256 + cb * 16), r);
      if (m < p.M) {
        const int n = n0 + cb * 16;
        __nv_bfloat16* dst = p.y + (size_t)m * p.ldy + n;
        if (n + 16 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const __nv_bfloat162 t = __floats2bfloat162_rn(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            o[i] = *reinterpret_cast<const uint32_t*>(&t);
          }
          reinterpret_cast<uint4*>(dst)[0] = make_uint4(o[0], o[1], o[2], o[3]);
          reinterpret_cast<uint4*>(dst)[1] = make_uint4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (n + i < p.N) dst[i] = f2bf(__uint_as_float(r[i]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256 * NACC) : "memory");
}

}  // namespace q4gm
}  // namespace b2l

using namespace b2l;
using namespace b2l::q4gm;

extern "C" int b2l_q4_gemm(const b2l_q4_linear_args* a, b2l_stream_t stream) {
  B2L_CHECK_ARG(a != nullptr, "b2l_q4_gemm: null args");
  B2L_CHECK_ARG(a->x && a->qw_tiled && a->scales && a->zeros && a->y, "b2l_q4_gemm: null pointer");
  B2L_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "b2l_q4_gemm: bad shape");
  B2L_CHECK_SUPPORTED(a->K % BK == 0, "b2l_q4_gemm: K=%d must be a multiple of %d", a->K, BK);
  B2L_CHECK_ARG(a->ldx >= a->K && a->ldx % 8 == 0 && a->ldy >= a->N, "b2l_q4_gemm: bad leading dimension (ldx %% 8 == 0)");
  B2L_CHECK_ARG(((uintptr_t)a->x % 16 == 0) && ((uintptr_t)a->qw_tiled % 16 == 0), "b2l_q4_gemm: x / qw_tiled must be 16-byte aligned");
  B2L_CHECK_ARG(a->sz_dtype == B2L_BF16 || a->sz_dtype == B2L_F32, "b2l_q4_gemm: bad sz_dtype");
  B2L_CHECK_SUPPORTED(a->prologue == B2L_PRO_NONE && a->epilogue == B2L_EPI_STORE, "b2l_q4_gemm: plain linear only (no fused prologue / epilogue)");
  // cuTensorMapEncodeTiled through the runtime (resolved once)
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
    return (EncodeFn)fn;
  }();
  if (encode == nullptr) {
    set_error("b2l_q4_gemm: cuTensorMapEncodeTiled is not available from this driver");
    return B2L_E_STATE;
  }
  // tile height: 256 token rows unless that tiling would fill the last wave of CTAs less than half (then 128)
  const long tiles256 = (long)((a->N + BN - 1) / BN) * ((a->M + 255) / 256);
  const int sms = sm_count();
  static const int env_nacc = [] { const char* e = getenv("B2L_GEMM_NACC"); return e ? atoi(e) : 0; }();
  int nacc = (a->M <= 128 || (tiles256 % sms != 0 && tiles256 % sms < sms / 2 && tiles256 < 4 * sms)) ? 1 : 2;
  if (env_nacc == 1 || env_nacc == 2) nacc = env_nacc;
  Params p;
  {
    // x[M, K] (leading dimension ldx) as (8 elements | M rows | K/8 chunks): a box of 8 x 256 x 8 is one stage's tile
    const cuuint64_t dims[3] = {8, (cuuint64_t)a->M, (cuuint64_t)(a->K / 8)};
    const cuuint64_t strides[2] = {(cuuint64_t)a->ldx * 2, 16};
    const cuuint32_t box[3] = {8, (cuuint32_t)(128 * nacc), (cuuint32_t)(BK / 8)};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult cr = encode(&p.xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(a->x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      set_error("b2l_q4_gemm: cuTensorMapEncodeTiled failed (%d) for M=%d K=%d ldx=%d", (int)cr, a->M, a->K, a->ldx);
      return B2L_E_ARG;
    }
  }
  p.x = (const __nv_bfloat16*)a->x; p.ldx = a->ldx;
  p.qwt = (const uint8_t*)a->qw_tiled;
  p.scales = a->scales; p.zeros = a->zeros; p.szdt = a->sz_dtype;
  p.y = (__nv_bfloat16*)a->y; p.ldy = a->ldy;
  p.M = a->M; p.N = a->N; p.K = a->K;
  static DynSmemCache smem_cache[2];
  if (nacc == 2) {
    if (int rc = ensure_dyn_smem(q4_gemm_kernel<2>, Cfg<2>::SMEM_BYTES, smem_cache[1])) return rc;
    dim3 grid((a->N + BN - 1) / BN, (a->M + 255) / 256);
    q4_gemm_kernel<2><<<grid, NTHREADS, Cfg<2>::SMEM_BYTES, (cudaStream_t)stream>>>(p);
  } else {
    if (int rc = ensure_dyn_smem(q4_gemm_kernel<1>, Cfg<1>::SMEM_BYTES, smem_cache[0])) return rc;
    dim3 grid((a->N + BN - 1) / BN, (a->M + 127) / 128);
    q4_gemm_kernel<1><<<grid, NTHREADS, Cfg<1>::SMEM_BYTES, (cudaStream_t)stream>>>(p);
  }
  B2L_LAUNCH_CHECK("q4_gemm_kernel");
  return 0;
}
