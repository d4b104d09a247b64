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
    a.get_error_string = reinterpret_cast<decltype(a.get_error_string)>(dlsym(h, "ncclGetErrorString"));
    a.ok = a.get_unique_id && a.comm_init_rank && a.comm_destroy && a.all_reduce && a.get_error_string;
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
  if (ctx->comm) api().comm_destroy(ctx->comm);
  delete ctx;
}

void dist_context_t::allreduce(double* buf, size_t count, bool is_max, cudaStream_t stream) const
{
  if (world <= 1 || count == 0) return;
  check(api().all_reduce(buf, buf, count, /*ncclFloat64*/ 8, is_max ? /*ncclMax*/ 2 : /*ncclSum*/ 0, comm, stream),
        "ncclAllReduce");
}

}  // namespace cuopt_b200
