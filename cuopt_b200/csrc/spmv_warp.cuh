// Warp-synchronous CSR row blocks: the SpMV core of the two hot PDHG kernels (sm_100a).
//
// The host cuts the CSR rows into "warp blocks": consecutive rows with at most WARP_NNZ (256) nonzeros and
// at most 32 rows; a longer row is a block of its own.  One warp owns one block at a time:
//   1. coalesced evict-first loads of the block's 256 (col,val) entries, 8 per lane, all issued up front;
//   2. 8 independent random gathers per lane of the multiplied vector (L1::no_allocate: the vector lives in
//      L2, an L1 allocation per gathered sector would only thrash; L2 evict-last policy, device_utils.cuh);
//   3. products parked in the warp's own 2 KB of shared memory (XOR-swizzled: conflict-free for the
//      coalesced stores and for the strided per-row reads);  __syncwarp, never __syncthreads;
//   4. lane r adds the products of row r left to right (bit-identical to a sequential CPU row sum) and runs
//      the fused row epilogue with operands it prefetched in step 1.
// Warps drift apart freely, so gathers of some warps overlap the row phase of others.
//
// Measured on B200 (scripts/spmv_warp_rows.cu, spmv_variants.cu, microbench_gather.cu; profiles/): for
// 8 nnz/row with uniformly random columns the pattern "12 B/nnz stream + one 8 B gather per nnz" is bound by
// the per-SM rate of uncoalesced L1TEX requests (one 128 B-line wavefront per gathered element, ~0.9-1.4 per
// clock per SM), not by HBM: the rowless upper bound is 23.6 us per 8M nnz, row-structured kernels reach
// ~40 us.  Block-synchronous products (41-43 us), TMA-staged pipelines (70-79 us at the occupancy their
// shared-memory footprint allows) and cp.async gathers (93-197 us) were measured and lost.
#pragma once

#include "device_utils.cuh"

namespace cuopt_b200 {

constexpr int WARP_THREADS = 256;                // CTA size of the warp-block kernels
constexpr int WARP_PER_CTA = WARP_THREADS / 32;
constexpr int WARP_NNZ     = 256;                // nonzeros per warp block (8 per lane)
constexpr int WARP_KN      = WARP_NNZ / 32;
constexpr int WARP_WIDE_RPL = 8;                 // rows per lane of the wide schedule (blocks of <= 256 rows)
__host__ __device__ constexpr int warp_swz(int e) { return e ^ ((e >> 4) & 7); }

struct csr_warp_view_t {
  const int* off;
  const int* idx;
  const double* val;
  int n_wb;
  const int2* wdesc;  // n_wb + 1 entries {first row, first nnz}; entry n_wb = {rows, nnz}
};

// Walks this warp's blocks (static round robin over all warps of the grid).
//   pre_op(row)              -> payload P, issued before the matrix loads of the block
//   row_op(row, sum, P)      exactly once per row, by one lane
// `pw` = this warp's WARP_NNZ doubles of shared memory.
// RPL = rows per lane: 1 for the usual blocks of <= 32 rows; 8 for the "wide" schedule (<= 256 rows per block) that the
// host cuts for very sparse matrices (< 4 nonzeros per row: the transposed row shard A_g^T of a many-GPU solve has
// n rows but only nnz/G nonzeros), where blocks of 32 rows would leave 7 of the 8 gather slots of every lane idle.
// INIT: the row sum starts from P::init (a partial sum of the same row over earlier column blocks, see the gather
// blocking in pdlp_kernels.cuh) instead of 0, so that block after block the additions stay strictly left to right.
template <typename P, int RPL = 1, bool INIT = false, typename PreOp, typename RowOp>
__device__ __forceinline__ void spmv_warp_rows(const csr_warp_view_t& A,
                                               const double* __restrict__ x,
                                               double* pw,
                                               PreOp& pre_op,
                                               RowOp& row_op,
                                               unsigned long long gather_policy)
{
  const int lane   = threadIdx.x & 31;
  const int gwarp  = blockIdx.x * WARP_PER_CTA + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * WARP_PER_CTA;
  for (int wb = gwarp; wb < A.n_wb; wb += nwarps) {
    const int2 d0 = __ldg(A.wdesc + wb), d1 = __ldg(A.wdesc + wb + 1);
    const int r0 = d0.x, lo = d0.y, r1 = d1.x, hi = d1.y;
    if (hi - lo > WARP_NNZ) {
      // one long row: lanes stride over it, fixed xor tree at the end
      P pl;
      if (lane == 0) pl = pre_op(r0);
      double acc = 0.0;
      for (int e = lo + lane; e < hi; e += 32) acc += ld_stream(A.val + e) * ld_l2(x + ld_stream(A.idx + e), gather_policy);
      acc = warp_sum(acc);
      if (lane == 0) {
        if constexpr (INIT) acc = pl.init + acc;  // long rows are tree sums anyway
        row_op(r0, acc, pl);
      }
      continue;
    }
    int rs[RPL], re[RPL];
    P pl[RPL];
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
      const int r = r0 + lane + 32 * q;
      rs[q] = re[q] = 0;
      if (r < r1) {
        rs[q] = __ldg(A.off + r) - lo;
        re[q] = __ldg(A.off + r + 1) - lo;
        pl[q] = pre_op(r);
      }
    }
    int c[WARP_KN];
    double a[WARP_KN];
#pragma unroll
    for (int k = 0; k < WARP_KN; ++k) {
      const int e = lo + lane + 32 * k;
      c[k]        = e < hi ? ld_stream(A.idx + e) : -1;
    }
#pragma unroll
    for (int k = 0; k < WARP_KN; ++k) {
      const int e = lo + lane + 32 * k;
      a[k]        = e < hi ? ld_stream(A.val + e) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < WARP_KN; ++k)
      if (c[k] >= 0) pw[warp_swz(lane + 32 * k)] = a[k] * ld_l2(x + c[k], gather_policy);
    __syncwarp();
    if constexpr (RPL == 1) {
      if (r0 + lane < r1) {
        double s = 0.0;
        if constexpr (INIT) s = pl[0].init;
        for (int p = rs[0]; p < re[0]; p += 8) {
          double v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (p + j < re[0]) ? pw[warp_swz(p + j)] : 0.0;
#pragma unroll
          for (int j = 0; j < 8; ++j) s += v[j];
        }
        row_op(r0 + lane, s, pl[0]);
      }
    } else {
#pragma unroll
      for (int q = 0; q < RPL; ++q) {
        const int r = r0 + lane + 32 * q;
        if (r < r1) {
          double s = 0.0;
          if constexpr (INIT) s = pl[q].init;
          for (int p = rs[q]; p < re[q]; ++p) s += pw[warp_swz(p)];
          row_op(r, s, pl[q]);
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace cuopt_b200
