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
#include <vector>

namespace cuopt_b200 {

struct dist_context_t {
  int rank  = 0;
  int world = 1;
  void* comm = nullptr;  // ncclComm_t

  // in-place all-reduce of `count` doubles on `stream` (sum, or max when is_max)
  void allreduce(double* buf, size_t count, bool is_max, cudaStream_t stream) const;
  // in-place all-gather: rank r's `count` doubles live at buf + r * count
  void allgather(double* buf, size_t count, cudaStream_t stream) const;
  // recv[0, count) = sum over ranks of their send[rank * count, (rank + 1) * count)
  void reduce_scatter(const double* send, double* recv, size_t count, cudaStream_t stream) const;

  // Peer memory (NVLink P2P stores): collective.  Every rank passes the BASE pointer of one of its own cudaMalloc
  // allocations; on return peers[r] addresses rank r's allocation from this process (peers[rank] == local).
  // Mappings are closed by close_peers / dist_destroy.  Returns false (on every rank) when some rank could not map
  // some peer (no peer access between the two devices): the caller then stays on the NCCL path.
  bool open_peers(void* local, void** peers, cudaStream_t stream) const;
  void close_peers(void** peers) const;

  mutable std::vector<void*> opened;  // cudaIpcOpenMemHandle results still mapped
};

constexpr int DIST_UNIQUE_ID_BYTES = 128;  // NCCL_UNIQUE_ID_BYTES

void dist_get_unique_id(char* id_out);  // rank 0; the caller broadcasts the bytes (torch.distributed, MPI, a file ...)
dist_context_t* dist_create(int rank, int world, const char* id);  // collective over all ranks; current CUDA device
void dist_destroy(dist_context_t* ctx);

}  // namespace cuopt_b200
