// Multi-GPU plumbing for the row-sharded PDLP: one process per GPU, one NCCL communicator per solve group.
//
// Nothing like this exists in the reference ("There is no support for leveraging multiple GPUs to solve a single
// problem", docs/cuopt/source/faq.rst:53); SURVEY.md §8(e) defines the scheme.  NCCL is loaded at run time with
// dlopen("libnccl.so.2") so that single-GPU users need no NCCL at all; under torchrun the library torch already
// loaded (2.28.9) is the one found.
#pragma once

#include "lp_problem.hpp"

#include <cuda_runtime.h>

#include <cstddef>

namespace cuopt_b200 {

struct dist_context_t {
  int rank  = 0;
  int world = 1;
  void* comm = nullptr;  // ncclComm_t

  // in-place all-reduce of `count` doubles on `stream` (sum, or max when is_max)
  void allreduce(double* buf, size_t count, bool is_max, cudaStream_t stream) const;
};

constexpr int DIST_UNIQUE_ID_BYTES = 128;  // NCCL_UNIQUE_ID_BYTES

void dist_get_unique_id(char* id_out);  // rank 0; the caller broadcasts the bytes (torch.distributed, MPI, a file ...)
dist_context_t* dist_create(int rank, int world, const char* id);  // collective over all ranks; current CUDA device
void dist_destroy(dist_context_t* ctx);

}  // namespace cuopt_b200
