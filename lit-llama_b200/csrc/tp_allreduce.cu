// One-shot all-reduce (sum) of a bf16 activation row over the GPUs of one node, through peer memory (NVLink).
//
// Tensor-parallel decode (lit_llama_b200/tp.py; SURVEY.md section 8e -- the reference has no multi-GPU inference, the
// split dims are those of scripts/convert_checkpoint.py:56-64) needs two sums per Block: after attn.c_proj and after
// mlp.c_proj (the row-parallel linears), 8..16 KB each.  At that size a collective is pure latency, so this kernel is
// built like NCCL's LL protocol but for exactly this case:
//   * every rank owns an exchange buffer that all peers have mapped (torch symmetric memory: plumbing only);
//   * a rank PUSHES its partial row into its slot of every peer's buffer as 8-byte words {2 bf16 values, epoch}:
//     an aligned 8-byte store is one NVLink transaction, so a word whose upper half shows the current epoch carries
//     valid data -- no separate flag, no fence, one NVLink one-way latency;
//   * it then polls its own buffer until the words of all peers show the epoch, adds the world's rows in fp32 in
//     RANK ORDER (every rank computes the same bits) and rounds once to bf16;
//   * slots are double-buffered by epoch parity: a peer can be at most one all-reduce ahead (it needs this rank's
//     next push to get further), and that one writes the other half;
//   * the epoch lives in device memory and is advanced by the kernel, so a captured CUDA graph replays correctly.
// CTAs of 128 threads x 4 words, each with its own epoch word (no cross-CTA step): small enough in registers to sit
// on an SM NEXT TO the two resident CTAs of the linears, so the next linear's CTAs still start early everywhere and
// stream their weights while the sum is in flight.  Launched with programmatic dependent launch between the
// row-parallel linear and the next fused [RMSNorm + linear]; waits are bounded (~2 s of polling); a timeout sets
// comm->status (sticky: later calls do not wait at all) and falls through instead of hanging the GPU.
#include "b2l_common.cuh"

namespace b2l {
namespace tp {

constexpr int THREADS = 128;
constexpr int PER_THREAD = 4;
constexpr int MAX_CTAS = 16;          // 16 x 512 words = 16384 bf16 values
constexpr int MAX_WORLD = 8;

struct Comm {
  unsigned long long* buf[MAX_WORLD];   // rank r's exchange buffer as mapped here; [parity][sender][max_pairs]
  int rank, world, max_pairs;
  unsigned int* epoch;
  int* status;
};

__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(THREADS, 4) tp_allreduce_kernel(const Comm c, const uint32_t* partial, uint32_t* out, int n_pairs) {
  pdl_wait();                 // `partial` is the previous kernel's output; the epoch word was written by the previous all-reduce
  pdl_launch_dependents();    // the next linear may start streaming its weights
  volatile unsigned int* my_epoch = reinterpret_cast<volatile unsigned int*>(c.epoch) + blockIdx.x;
  const unsigned int epoch = *my_epoch + 1u;
  const int parity = (int)(epoch & 1u);
  const size_t slot = (size_t)c.max_pairs;
  uint32_t mine[PER_THREAD];
  // ---- push: this rank's row into slot [parity][rank] of every peer
#pragma unroll
  for (int j = 0; j < PER_THREAD; ++j) {
    const int i = (blockIdx.x * PER_THREAD + j) * THREADS + threadIdx.x;
    mine[j] = 0u;
    if (i < n_pairs) {
      asm volatile("ld.global.u32 %0, [%1];" : "=r"(mine[j]) : "l"(partial + i) : "memory");   // coherent (PDL)
      const unsigned long long w = ((unsigned long long)epoch << 32) | mine[j];
#pragma unroll
      for (int r = 0; r < MAX_WORLD; ++r)
        if (r < c.world && r != c.rank) st_relaxed_sys_u64(c.buf[r] + ((size_t)parity * c.world + c.rank) * slot + i, w);
    }
  }
  // ---- gather + sum in rank order
  const unsigned long long* local = c.buf[c.rank] + (size_t)parity * c.world * slot;
  const long long t0 = clock64();
  bool timed_out = *reinterpret_cast<volatile int*>(c.status) != 0;   // sticky: after one timeout nobody waits again
#pragma unroll
  for (int j = 0; j < PER_THREAD; ++j) {
    const int i = (blockIdx.x * PER_THREAD + j) * THREADS + threadIdx.x;
    if (i >= n_pairs) continue;
    float lo = 0.f, hi = 0.f;
    for (int r = 0; r < c.world; ++r) {
      uint32_t v = mine[j];
      if (r != c.rank) {
        unsigned long long w = ld_relaxed_sys_u64(local + (size_t)r * slot + i);
        while ((unsigned int)(w >> 32) != epoch) {
          if (timed_out || clock64() - t0 > 4000000000LL) { timed_out = true; break; }
          w = ld_relaxed_sys_u64(local + (size_t)r * slot + i);
        }
        v = (uint32_t)w;
      }
      lo += __uint_as_float(v << 16);
      hi += __uint_as_float(v & 0xffff0000u);
    }
    const __nv_bfloat162 o = __floats2bfloat162_rn(lo, hi);
    out[i] = *reinterpret_cast<const uint32_t*>(&o);
  }
  if (timed_out) *c.status = 1;
  __syncthreads();
  if (threadIdx.x == 0) *my_epoch = epoch;
}

}  // namespace tp
}  // namespace b2l

using namespace b2l;

extern "C" size_t b2l_tp_buffer_bytes(int world, int max_elems) {
  if (world < 1 || world > tp::MAX_WORLD || max_elems <= 0 || max_elems % 2 != 0) return 0;
  return (size_t)2 * world * (max_elems / 2) * 8;
}

extern "C" int b2l_tp_allreduce(const b2l_tp_comm* comm, const void* partial, void* out, int n, int flags, b2l_stream_t stream) {
  B2L_CHECK_ARG(comm != nullptr && partial != nullptr && out != nullptr, "b2l_tp_allreduce: null pointer");
  B2L_CHECK_ARG(comm->world >= 1 && comm->world <= tp::MAX_WORLD && comm->rank >= 0 && comm->rank < comm->world,
                "b2l_tp_allreduce: bad rank %d / world %d", comm->rank, comm->world);
  B2L_CHECK_ARG(n > 0 && n % 2 == 0 && n <= comm->max_elems, "b2l_tp_allreduce: n=%d must be even and <= max_elems=%d", n, comm->max_elems);
  B2L_CHECK_SUPPORTED(n <= 2 * tp::MAX_CTAS * tp::PER_THREAD * tp::THREADS, "b2l_tp_allreduce: n=%d > %d", n,
                      2 * tp::MAX_CTAS * tp::PER_THREAD * tp::THREADS);
  B2L_CHECK_ARG(comm->epoch != nullptr && comm->status != nullptr, "b2l_tp_allreduce: null epoch / status word");
  tp::Comm c;
  for (int r = 0; r < tp::MAX_WORLD; ++r) {
    c.buf[r] = r < comm->world ? (unsigned long long*)comm->peer_buf[r] : nullptr;
    if (r < comm->world) B2L_CHECK_ARG(c.buf[r] != nullptr && ((uintptr_t)c.buf[r] % 8 == 0), "b2l_tp_allreduce: bad peer buffer %d", r);
  }
  c.rank = comm->rank; c.world = comm->world; c.max_pairs = comm->max_elems / 2;
  c.epoch = comm->epoch; c.status = comm->status;
  const int n_pairs = n / 2;
  const int grid = (n_pairs + tp::PER_THREAD * tp::THREADS - 1) / (tp::PER_THREAD * tp::THREADS);
  LaunchCfg lc(dim3(grid), dim3(tp::THREADS), 0, (cudaStream_t)stream, (flags & B2L_F_PDL) != 0);
  B2L_CUDA(cudaLaunchKernelEx(&lc.cfg, tp::tp_allreduce_kernel, c, (const uint32_t*)partial, (uint32_t*)out, n_pairs));
  return 0;
}
