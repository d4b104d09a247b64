// Device CSR transpose (setup only, once per solve): A (rows x cols) -> A^T as CSR with row indices ascending inside
// each transposed row — the same ordering cusparseCsr2cscEx2 hands the reference
// (cpp/src/mip/problem/problem.cu:277-309 via raft::sparse::linalg::csr_transpose).
//
// Stable by construction: an LSD radix sort (cub::DeviceRadixSort, stable) of the entry ids by column index keeps
// entries of one column in their original row-major order, so the result is deterministic and identical to a
// sequential counting-sort transpose.  CUB is used here like the reference uses cuSPARSE: a library call in the
// one-time setup, not in the iteration loop.
#include "device_utils.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace cuopt_b200 {

namespace {

__global__ void k_expand_rows_and_count(int rows, const int* __restrict__ off, const int* __restrict__ idx,
                                        int* __restrict__ entry_row, int* __restrict__ entry_id, int* __restrict__ col_count)
{
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    for (int p = off[r] + lane; p < off[r + 1]; p += 32) {
      entry_row[p] = r;
      entry_id[p]  = p;
      atomicAdd(col_count + idx[p], 1);  // integer counts: order independent
    }
  }
}

__global__ void k_permute(int nnz, const int* __restrict__ sorted_entry, const int* __restrict__ entry_row,
                          const double* __restrict__ val, int* __restrict__ tidx, double* __restrict__ tval)
{
  const int stride = gridDim.x * blockDim.x;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += stride) {
    const int e = sorted_entry[q];
    tidx[q]     = entry_row[e];
    tval[q]     = val[e];
  }
}

}  // namespace

// toff must hold cols + 1 ints, tidx / tval nnz elements.  All pointers are device pointers.
void csr_transpose_device(int rows, int cols, int nnz, const int* off, const int* idx, const double* val, int* toff,
                          int* tidx, double* tval, cudaStream_t stream)
{
  dvec<int> entry_row(nnz), entry_id(nnz), keys_out(nnz), sorted_entry(nnz), col_count((size_t)cols + 1);
  col_count.zero(stream);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  k_expand_rows_and_count<<<std::max(1, std::min((rows + 7) / 8, sms * 16)), 256, 0, stream>>>(
    rows, off, idx, entry_row.data(), entry_id.data(), col_count.data());
  CUOPT_CUDA_TRY(cudaGetLastError());

  size_t tmp_scan = 0, tmp_sort = 0;
  CUOPT_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, col_count.data(), toff, cols + 1, stream));
  int end_bit = 1;
  while (end_bit < 31 && (1LL << end_bit) < (long long)cols) ++end_bit;
  CUOPT_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, idx, keys_out.data(), entry_id.data(),
                                                  sorted_entry.data(), nnz, 0, end_bit, stream));
  dvec<unsigned char> tmp(std::max(tmp_scan, tmp_sort) + 16);
  size_t bytes = tmp.size();
  CUOPT_CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.data(), bytes, col_count.data(), toff, cols + 1, stream));
  bytes = tmp.size();
  CUOPT_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp.data(), bytes, idx, keys_out.data(), entry_id.data(),
                                                  sorted_entry.data(), nnz, 0, end_bit, stream));
  k_permute<<<std::max(1, std::min((nnz + 255) / 256, sms * 8)), 256, 0, stream>>>(nnz, sorted_entry.data(),
                                                                                   entry_row.data(), val, tidx, tval);
  CUOPT_CUDA_TRY(cudaGetLastError());
  CUOPT_CUDA_TRY(cudaStreamSynchronize(stream));  // temporaries die with this scope
}

}  // namespace cuopt_b200
