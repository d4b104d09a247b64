#include "dist_comm.hpp"

#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>

namespace cuopt_b200 {

namespace {

// The handful of NCCL entry points used, with the ABI of nccl.h (2.x): ncclUniqueId is 128 opaque bytes,
// ncclFloat64 = 8, ncclSum = 0, ncclMax = 2.
struct nccl_id_t {
  char internal[DIST_UNIQUE_ID_BYTES];
};
using comm_t = void*;
struct nccl_api_t {
  int (*get_unique_id)(nccl_id_t*)                                                        = nullptr;
  int (*comm_init_rank)(comm_t*, int, nccl_id_t, int)                                     = nullptr;
  int (*comm_destroy)(comm_t)                                                             = nullptr;
  int (*all_reduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t)           = nullptr;
  int (*all_gather)(const void*, void*, size_t, int, comm_t, cudaStream_t)                = nullptr;
  int (*reduce_scatter)(const void*, void*, size_t, int, int, comm_t, cudaStream_t)       = nullptr;
  const char* (*get_error_string)(int)                                                    = nullptr;
  bool ok                                                                                 = false;
};

nccl_api_t& api()
{
  static nccl_api_t a;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) return;
    a.get_unique_id    = reinterpret_cast<decltype(a.get_unique_id)>(dlsym(h, "ncclGetUniqueId"));
    a.comm_init_rank   = reinterpret_cast<decltype(a.comm_init_rank)>(dlsym(h, "ncclCommInitRank"));
    a.comm_destroy     = reinterpret_cast<decltype(a.comm_destroy)>(dlsym(h, "ncclCommDestroy"));
    a.all_reduce       = reinterpret_cast<decltype(a.all_reduce)>(dlsym(h, "ncclAllReduce"));
    a.all_gather       = reinterpret_cast<decltype(a.all_gather)>(dlsym(h, "ncclAllGather"));
    a.reduce_scatter   = reinterpret_cast<decltype(a.reduce_scatter)>(dlsym(h, "ncclReduceScatter"));
    a.get_error_string = reinterpret_cast<decltype(a.get_error_string)>(dlsym(h, "ncclGetErrorString"));
    a.ok = a.get_unique_id && a.comm_init_rank && a.comm_destroy && a.all_reduce && a.all_gather && a.reduce_scatter &&
           a.get_error_string;
  });
  if (!a.ok) throw lp_error(error_type_t::RuntimeError, "NCCL (libnccl.so.2) could not be loaded: multi-GPU solve unavailable");
  return a;
}

void check(int rc, const char* what)
{
  if (rc != 0)
    throw lp_error(error_type_t::RuntimeError, std::string("NCCL error in ") + what + ": " + api().get_error_string(rc));
}

}  // namespace

void dist_get_unique_id(char* id_out)
{
  nccl_id_t id;
  check(api().get_unique_id(&id), "ncclGetUniqueId");
  std::memcpy(id_out, id.internal, DIST_UNIQUE_ID_BYTES);
}

dist_context_t* dist_create(int rank, int world, const char* id_bytes)
{
  if (world < 1 || rank < 0 || rank >= world) throw lp_error(error_type_t::InvalidArgument, "bad rank / world size");
  nccl_id_t id;
  std::memcpy(id.internal, id_bytes, DIST_UNIQUE_ID_BYTES);
  auto* ctx  = new dist_context_t;
  ctx->rank  = rank;
  ctx->world = world;
  comm_t c   = nullptr;
  try {
    check(api().comm_init_rank(&c, world, id, rank), "ncclCommInitRank");
  } catch (...) {
    delete ctx;
    throw;
  }
  ctx->comm = c;
  return ctx;
}

void dist_destroy(dist_context_t* ctx)
{
  if (!ctx) return;
  for (void* p : ctx->opened) cudaIpcCloseMemHandle(p);
  ctx->opened.clear();
  if (ctx->comm) api().comm_destroy(ctx->comm);
  delete ctx;
}

void dist_context_t::allreduce(double* buf, size_t count, bool is_max, cudaStream_t stream) const
{
  if (world <= 1 || count == 0) return;
  check(api().all_reduce(buf, buf, count, /*ncclFloat64*/ 8, is_max ? /*ncclMax*/ 2 : /*ncclSum*/ 0, comm, stream),
        "ncclAllReduce");
}

void dist_context_t::allgather(double* buf, size_t count, cudaStream_t stream) const
{
  if (world <= 1 || count == 0) return;
  check(api().all_gather(buf + (size_t)rank * count, buf, count, /*ncclFloat64*/ 8, comm, stream), "ncclAllGather");
}

void dist_context_t::reduce_scatter(const double* send, double* recv, size_t count, cudaStream_t stream) const
{
  if (count == 0) return;
  if (world <= 1) {
    cudaMemcpyAsync(recv, send, count * sizeof(double), cudaMemcpyDeviceToDevice, stream);
    return;
  }
  check(api().reduce_scatter(send, recv, count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm, stream), "ncclReduceScatter");
}

namespace {
void cuda_check(cudaError_t e, const char* what)
{
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw lp_error(error_type_t::RuntimeError, std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
  }
}
}  // namespace

// IPC handles (64 bytes each) travel through one NCCL all-gather of doubles; a final all-reduce(max) of the failure
// flags makes the outcome unanimous.
bool dist_context_t::open_peers(void* local, void** peers, cudaStream_t stream) const
{
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  constexpr size_t W = sizeof(cudaIpcMemHandle_t) / sizeof(double);
  for (int r = 0; r < world; ++r) peers[r] = nullptr;
  peers[rank] = local;
  if (world <= 1) return true;
  std::vector<cudaIpcMemHandle_t> handles(world);
  cuda_check(cudaIpcGetMemHandle(&handles[rank], local), "cudaIpcGetMemHandle");
  double* d = nullptr;
  cuda_check(cudaMalloc(&d, (world * W + 1) * sizeof(double)), "cudaMalloc");
  cuda_check(cudaMemcpyAsync(d + rank * W, &handles[rank], sizeof(cudaIpcMemHandle_t), cudaMemcpyHostToDevice, stream),
             "cudaMemcpyAsync");
  allgather(d, W, stream);
  cuda_check(cudaMemcpyAsync(handles.data(), d, world * sizeof(cudaIpcMemHandle_t), cudaMemcpyDeviceToHost, stream),
             "cudaMemcpyAsync");
  cuda_check(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  double failed = 0.0;
  std::vector<void*> mine;
  for (int r = 0; r < world && failed == 0.0; ++r) {
    if (r == rank) continue;
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, handles[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      failed = 1.0;
    } else {
      peers[r] = p;
      mine.push_back(p);
    }
  }
  cuda_check(cudaMemcpyAsync(d + world * W, &failed, sizeof(double), cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync");
  allreduce(d + world * W, 1, true, stream);
  cuda_check(cudaMemcpyAsync(&failed, d + world * W, sizeof(double), cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync");
  cuda_check(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  cudaFree(d);
  if (failed != 0.0) {
    for (void* p : mine) cudaIpcCloseMemHandle(p);
    for (int r = 0; r < world; ++r)
      if (r != rank) peers[r] = nullptr;
    return false;
  }
  opened.insert(opened.end(), mine.begin(), mine.end());
  return true;
}

void dist_context_t::close_peers(void** peers) const
{
  for (int r = 0; r < world; ++r) {
    if (r == rank || peers[r] == nullptr) continue;
    cudaIpcCloseMemHandle(peers[r]);
    for (auto it = opened.begin(); it != opened.end(); ++it)
      if (*it == peers[r]) { opened.erase(it); break; }
    peers[r] = nullptr;
  }
}

}  // namespace cuopt_b200
