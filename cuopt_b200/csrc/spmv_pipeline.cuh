// TMA-staged, software-pipelined CSR row-block SpMV (sm_100a).
//
// A persistent CTA walks the row blocks assigned to it (static round robin).  Per block:
//   * the column indices and row offsets are pulled into shared memory by the bulk-copy engine
//     (cp.async.bulk global->shared completing on an mbarrier), STAGES blocks ahead;
//   * every thread issues 8 random gathers of the multiplied vector plus 8 coalesced streaming loads of
//     the matrix values for block i+1 and only THEN runs the row phase of block i (one thread per row adds
//     that row's products left to right from shared memory and executes the fused row epilogue), so the
//     gathers of the next block are in flight while the current block is reduced;
//   * products of block i+1 are parked in the other half of a double-buffered, padded shared array;
//     one __syncthreads per block.
//
// Why this shape (measured on B200, scripts/microbench_gather.cu + profiles/): "12 B/nnz stream + one
// random 8 B gather per nnz from an L2-resident vector" is bound by the L2 sector rate (a 32 B sector per
// gather, ~400 G sectors/s), not by HBM: 8M nnz cannot go below ~21-25 us.  Reaching that needs ~3000
// gathers in flight per SM at all times (Little's law at ~2000 cycles loaded latency); a CTA that
// alternates "gather" and "reduce" phases only averages a quarter of its peak, hence the overlap above.
#pragma once

#include "device_utils.cuh"

#include <cstdint>

namespace cuopt_b200 {

constexpr int SPMV_THREADS = 256;
constexpr int SPMV_NNZ     = 2048;          // nonzeros per row block (8 per thread)
constexpr int SPMV_ROWS    = SPMV_THREADS;  // max rows per row block (1 per thread)
constexpr int SPMV_KN      = SPMV_NNZ / SPMV_THREADS;
constexpr int SPMV_PADDED  = SPMV_NNZ + (SPMV_NNZ >> 3);
__host__ __device__ constexpr int spmv_pad(int e) { return e + (e >> 3); }
// slack the 16-byte granular bulk copies may read past the end of idx / off: see upload_csr
constexpr int SPMV_TAIL_SLACK = 8;

struct csr_view_t {
  int rows;
  const int* off;
  const int* idx;
  const double* val;
  int n_blocks;
  const int4* blk;  // {first row, one-past-last row, first nnz, one-past-last nnz}
};

// ---- mbarrier / bulk-copy PTX -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
    "selp.u32 %0, 1, 0, p;\n"
    "}\n"
    : "=r"(ok)
    : "r"(smem_u32(bar)), "r"(parity)
    : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  while (!mbar_try_wait(bar, parity)) {}
}
// 1-D bulk copy global -> shared, bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- shared-memory layout ---------------------------------------------------------------------------
struct __align__(16) spmv_stage_t {
  int idx[SPMV_NNZ + 8];   // window starts at (lo & ~3)
  int off[SPMV_ROWS + 8];  // window starts at (r0 & ~3)
  int4 desc;               // {r0, r1, lo, hi}
};

template <int NV, int STAGES>
struct spmv_smem_t {
  spmv_stage_t stage[STAGES];
  double prod[2][NV * SPMV_PADDED];  // double buffered: block i is reduced while block i+1 is produced
  double red[32];
  uint64_t full[STAGES];
};

__device__ __forceinline__ bool spmv_is_regular(const int4& d) { return d.w - d.z <= SPMV_NNZ && d.w > d.z; }

// One elected thread: publish block descriptor `d` and launch the bulk copies of its stage.
__device__ __forceinline__ void spmv_issue_stage(const csr_view_t& A, int4 d, spmv_stage_t* st, uint64_t* bar)
{
  st->desc = d;
  if (!spmv_is_regular(d)) {
    // long single row (read straight from global memory by the consumers) or only empty rows: nothing to stage
    mbar_arrive(bar);
    return;
  }
  const int r0 = d.x, r1 = d.y, lo = d.z, hi = d.w;
  const int lo_i = lo & ~3, r0_a = r0 & ~3;
  const uint32_t bi = (uint32_t)(((hi - lo_i) * 4 + 15) & ~15);
  const uint32_t bo = (uint32_t)(((r1 + 1 - r0_a) * 4 + 15) & ~15);
  mbar_arrive_expect_tx(bar, bi + bo);
  bulk_g2s(st->idx, A.idx + lo_i, bi, bar);
  bulk_g2s(st->off, A.off + r0_a, bo, bar);
}

// Runs the whole pipeline for this CTA.
//   pre_op(row)                -> payload P   (issued one block ahead of its use)
//   row_op(row, sums[NV], P)                  (exactly once per row, by one thread)
template <int NV, int STAGES, typename P, typename PreOp, typename RowOp>
__device__ __forceinline__ void spmv_pipeline(const csr_view_t& A,
                                              const double* const* x,
                                              spmv_smem_t<NV, STAGES>& sm,
                                              PreOp& pre_op,
                                              RowOp& row_op)
{
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&sm.full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int first = blockIdx.x, step = gridDim.x;
  if (first >= A.n_blocks) return;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      const int b = first + s * step;
      if (b < A.n_blocks) spmv_issue_stage(A, __ldg(A.blk + b), &sm.stage[s], &sm.full[s]);
    }
  }

  // registers of the block "in preparation": gathered vector entries, matrix values, epilogue operands
  double g[NV][SPMV_KN], a[SPMV_KN];
  P pl_next, pl_cur;
  int4 d_next = make_int4(0, 0, 0, 0), d_cur;

  // Issue every load block `d` needs (stage `st` has landed): gathers, matrix values, epilogue operands.
  auto issue_loads = [&](const int4& d, const spmv_stage_t& st) {
    const int r0 = d.x, r1 = d.y, lo = d.z, hi = d.w;
    if (spmv_is_regular(d)) {
      const int cnt = hi - lo;
      const int di  = lo - (lo & ~3);
#pragma unroll
      for (int k = 0; k < SPMV_KN; ++k) {
        const int e = tid + k * SPMV_THREADS;
        if (e < cnt) {
          const int c = st.idx[di + e];
#pragma unroll
          for (int v = 0; v < NV; ++v) g[v][k] = __ldg(x[v] + c);
        }
      }
#pragma unroll
      for (int k = 0; k < SPMV_KN; ++k) {
        const int e = tid + k * SPMV_THREADS;
        if (e < cnt) a[k] = ld_stream(A.val + lo + e);
      }
    }
    if (hi - lo <= SPMV_NNZ) {
      const int r = r0 + tid;
      if (r < r1) pl_next = pre_op(r);
    }
  };
  // Park the products of block `d` (whose loads were issued by issue_loads) in product buffer `buf`.
  auto store_products = [&](const int4& d, int buf) {
    if (!spmv_is_regular(d)) return;
    const int cnt = d.w - d.z;
#pragma unroll
    for (int k = 0; k < SPMV_KN; ++k) {
      const int e = tid + k * SPMV_THREADS;
      if (e < cnt) {
        const int p = spmv_pad(e);
#pragma unroll
        for (int v = 0; v < NV; ++v) sm.prod[buf][v * SPMV_PADDED + p] = a[k] * g[v][k];
      }
    }
  };

  // prologue: first block
  mbar_wait(&sm.full[0], 0);
  d_next = sm.stage[0].desc;
  issue_loads(d_next, sm.stage[0]);
  store_products(d_next, 0);
  __syncthreads();

  int it = 0;
  for (int b = first; b < A.n_blocks; b += step, ++it) {
    const int s      = it % STAGES;
    spmv_stage_t& st = sm.stage[s];
    d_cur            = d_next;
    pl_cur           = pl_next;
    const int r0 = d_cur.x, r1 = d_cur.y, lo = d_cur.z, hi = d_cur.w;

    // (A) next block: wait for its stage, put its gathers / values / epilogue operands in flight
    const bool has_next = (b + step) < A.n_blocks;
    // descriptor of the block that will refill the current stage (thread 0; latency hidden behind this block)
    const int rb      = b + STAGES * step;
    const bool refill = (tid == 0) && (rb < A.n_blocks);
    int4 d_refill     = make_int4(0, 0, 0, 0);
    if (refill) d_refill = __ldg(A.blk + rb);
    if (has_next) {
      const int sn = (it + 1) % STAGES;
      mbar_wait(&sm.full[sn], (uint32_t)(((it + 1) / STAGES) & 1));
      d_next = sm.stage[sn].desc;
      issue_loads(d_next, sm.stage[sn]);
    }

    // (B) current block: row phase
    if (hi - lo > SPMV_NNZ) {
      // one long row: strided partial sums straight from global memory, then the fixed block tree
      P pl = pre_op(r0);
      double acc[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] = 0.0;
      for (int e = lo + tid; e < hi; e += SPMV_THREADS) {
        const int c     = ld_stream(A.idx + e);
        const double av = ld_stream(A.val + e);
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] += av * __ldg(x[v] + c);
      }
      double tot[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) tot[v] = block_reduce(acc[v], sm.red);
      if (tid == 0) row_op(r0, tot, pl);
    } else {
      const int r = r0 + tid;
      if (r < r1) {
        const int cnt  = hi - lo;
        const int doff = r0 - (r0 & ~3);
        // a block made only of empty rows stages nothing (cnt == 0): every extent is empty
        const int rs = cnt ? st.off[doff + tid] - lo : 0;
        const int re = cnt ? st.off[doff + tid + 1] - lo : 0;
        const double* pr = sm.prod[it & 1];
        double sum[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) sum[v] = 0.0;
        for (int p = rs; p < re; ++p) {
          const int q = spmv_pad(p);
#pragma unroll
          for (int v = 0; v < NV; ++v) sum[v] += pr[v * SPMV_PADDED + q];
        }
        row_op(r, sum, pl_cur);
      }
    }

    // (C) products of the next block into the other buffer (waits for its gathers)
    if (has_next) store_products(d_next, (it + 1) & 1);
    __syncthreads();  // next products visible; current stage and product buffer fully consumed
    if (refill) spmv_issue_stage(A, d_refill, &st, &sm.full[s]);
  }
}

}  // namespace cuopt_b200
