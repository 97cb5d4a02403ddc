// Multi-GPU plumbing of the C ABI (include/sdfb200.h, "multi-GPU" section; SURVEY.md section 8(b) row
// sdf_allgather_slabs and 8(e)).  No reference interface exists to mirror (/root/reference/README.md:1).
//
// Two transports between the ranks of one node (one process per GPU):
//   * NCCL, resolved at run time (dlopen of the libnccl.so.2 the process already has, or the system one): the in-place
//     slab all-gather the survey names, the rank barrier, and the bootstrap of the second transport;
//   * peer memory: one symmetric buffer per rank, every rank's copy mapped into all the others through CUDA IPC.  A rank
//     PUSHES each finished sub-slab into the same offset of every peer's copy with the copy engines (cudaMemcpyAsync
//     device-to-device over NVLink / NVSwitch) while its SMs keep decoding the next sub-slab - the persistent decoder
//     kernel occupies every SM, so a collective that needs SMs of its own (an NCCL kernel) cannot overlap with it.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <dlfcn.h>

#include "../../include/sdfb200.h"
#include "comm.h"

namespace sdfb {
int set_error(int code, const char* fmt, ...);
}
using sdfb::set_error;

namespace {

// the handful of NCCL entry points used, by their stable C signatures (nccl.h 2.x)
struct NcclId { char b[128]; };     // ncclUniqueId: passed by value
typedef int (*FnGetUniqueId)(void* id128);
typedef int (*FnCommInitRank)(void** comm, int nranks, NcclId id, int rank);
typedef int (*FnCommDestroy)(void* comm);
typedef int (*FnAllGather)(const void* send, void* recv, size_t count, int dtype, void* comm, cudaStream_t st);
typedef int (*FnAllReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t st);
typedef const char* (*FnErrStr)(int);
constexpr int kNcclUint8 = 1, kNcclInt32 = 2, kNcclSum = 0;

struct NcclApi {
  void* handle = nullptr;
  FnGetUniqueId get_unique_id = nullptr;
  FnCommInitRank comm_init_rank = nullptr;
  FnCommDestroy comm_destroy = nullptr;
  FnAllGather all_gather = nullptr;
  FnAllReduce all_reduce = nullptr;
  FnErrStr err_str = nullptr;
  bool ok = false;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  // a library of that soname already in the process (torch's bundled NCCL) is what dlopen returns; otherwise the system's
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return nullptr;
  api.get_unique_id = reinterpret_cast<FnGetUniqueId>(dlsym(api.handle, "ncclGetUniqueId"));
  api.comm_init_rank = reinterpret_cast<FnCommInitRank>(dlsym(api.handle, "ncclCommInitRank"));
  api.comm_destroy = reinterpret_cast<FnCommDestroy>(dlsym(api.handle, "ncclCommDestroy"));
  api.all_gather = reinterpret_cast<FnAllGather>(dlsym(api.handle, "ncclAllGather"));
  api.all_reduce = reinterpret_cast<FnAllReduce>(dlsym(api.handle, "ncclAllReduce"));
  api.err_str = reinterpret_cast<FnErrStr>(dlsym(api.handle, "ncclGetErrorString"));
  api.ok = api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_gather && api.all_reduce;
  return api.ok ? &api : nullptr;
}

int nccl_fail(const char* what, int rc) {
  NcclApi* a = nccl_api();
  return set_error(SDFB_E_CUDA, "%s failed: NCCL error %d (%s)", what, rc, (a && a->err_str) ? a->err_str(rc) : "?");
}

#define CU_TRY(expr)                                                                         \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return set_error(SDFB_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

struct DevGuard {
  int prev = -1;
  explicit DevGuard(int dev) { if (cudaGetDevice(&prev) == cudaSuccess) cudaSetDevice(dev); }
  ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int comm_finish_init(sdfb_comm* c) {
  for (cudaStream_t& s : c->st_push) CU_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  CU_TRY(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
  for (cudaEvent_t& e : c->ev_done) CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CU_TRY(cudaMalloc(&c->token, 2 * sizeof(int)));
  CU_TRY(cudaMemset(c->token, 0, 2 * sizeof(int)));
  CU_TRY(cudaMalloc(&c->hbuf, static_cast<size_t>(c->world) * sizeof(cudaIpcMemHandle_t)));
  CU_TRY(cudaDeviceSynchronize());
  return SDFB_OK;
}

void release_shared(sdfb_comm* c) {
  for (int r = 0; r < static_cast<int>(c->peer.size()); ++r)
    if (r != c->rank && c->peer[r] != nullptr) cudaIpcCloseMemHandle(c->peer[r]);
  c->peer.clear();
  if (c->local) cudaFree(c->local);
  c->local = nullptr;
  c->bytes = 0;
}

}  // namespace

namespace sdfb {

int comm_barrier(sdfb_comm* c, cudaStream_t st) {
  if (c->world == 1) return SDFB_OK;
  if (const char* e = std::getenv("SDFB_PUSH_OFF")) if (e[0] == '2') return SDFB_OK;   // diagnostics only
  NcclApi* a = nccl_api();
  if (!a) return set_error(SDFB_E_CUDA, "NCCL is not available (libnccl.so.2 could not be loaded)");
  const int rc = a->all_reduce(c->token, c->token + 1, 1, kNcclInt32, kNcclSum, c->nccl, st);
  return rc == 0 ? SDFB_OK : nccl_fail("ncclAllReduce (barrier)", rc);
}

// Collective: every rank allocates `bytes`, the IPC handles travel through an NCCL all-gather, every rank maps the others.
int comm_shared_alloc(sdfb_comm* c, size_t bytes) {
  if (c->bytes >= bytes && c->local != nullptr) return SDFB_OK;
  NcclApi* a = nccl_api();
  if (c->world > 1 && !a) return set_error(SDFB_E_CUDA, "NCCL is not available (libnccl.so.2 could not be loaded)");
  CU_TRY(cudaDeviceSynchronize());
  if (c->world > 1 && c->local != nullptr) {   // nobody may still be pushing into the buffer that is about to go
    int brc = comm_barrier(c, nullptr);
    if (brc) return brc;
    CU_TRY(cudaStreamSynchronize(nullptr));
  }
  release_shared(c);
  CU_TRY(cudaMalloc(&c->local, bytes));
  c->bytes = bytes;
  c->peer.assign(c->world, nullptr);
  c->peer[c->rank] = c->local;
  if (c->world == 1) return SDFB_OK;
  cudaIpcMemHandle_t mine;
  CU_TRY(cudaIpcGetMemHandle(&mine, c->local));
  char* hb = static_cast<char*>(c->hbuf);
  CU_TRY(cudaMemcpy(hb + static_cast<size_t>(c->rank) * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice));
  const int rc = a->all_gather(hb + static_cast<size_t>(c->rank) * sizeof(mine), hb, sizeof(mine), kNcclUint8, c->nccl, nullptr);
  if (rc != 0) return nccl_fail("ncclAllGather (IPC handles)", rc);
  CU_TRY(cudaStreamSynchronize(nullptr));
  std::vector<cudaIpcMemHandle_t> all(c->world);
  CU_TRY(cudaMemcpy(all.data(), hb, all.size() * sizeof(mine), cudaMemcpyDeviceToHost));
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaError_t e = cudaIpcOpenMemHandle(&c->peer[r], all[r], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      c->peer[r] = nullptr;
      return set_error(SDFB_E_CUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s - the ranks must be processes of one node whose GPUs "
                       "have peer access", r, cudaGetErrorString(e));
    }
  }
  // nobody may push before every rank has mapped every buffer
  int brc = comm_barrier(c, nullptr);
  if (brc) return brc;
  CU_TRY(cudaStreamSynchronize(nullptr));
  return SDFB_OK;
}

int comm_push(sdfb_comm* c, size_t offset, size_t bytes, cudaStream_t after) {
  if (c->world == 1 || bytes == 0) return SDFB_OK;
  if (std::getenv("SDFB_PUSH_OFF") != nullptr) return SDFB_OK;      // diagnostics (tools/prof_sharded.py): peers get nothing
  if (offset + bytes > c->bytes) return set_error(SDFB_E_INVALID, "push outside the symmetric buffer");
  CU_TRY(cudaEventRecord(c->ev_ready, after));
  for (cudaStream_t s : c->st_push) CU_TRY(cudaStreamWaitEvent(s, c->ev_ready, 0));
  const char* src = static_cast<const char*>(c->local) + offset;
  int k = 0;
  for (int d = 1; d < c->world; ++d) {     // rank r starts with r + 1: at any moment the ranks write to different peers
    const int r = (c->rank + d) % c->world;
    CU_TRY(cudaMemcpyAsync(static_cast<char*>(c->peer[r]) + offset, src, bytes, cudaMemcpyDeviceToDevice,
                           c->st_push[k++ % sdfb_comm::kPushStreams]));
  }
  return SDFB_OK;
}

int comm_join_pushes(sdfb_comm* c, cudaStream_t st) {
  if (c->world == 1) return SDFB_OK;
  for (int k = 0; k < sdfb_comm::kPushStreams; ++k) {
    CU_TRY(cudaEventRecord(c->ev_done[k], c->st_push[k]));
    CU_TRY(cudaStreamWaitEvent(st, c->ev_done[k], 0));
  }
  return SDFB_OK;
}

}  // namespace sdfb

extern "C" {

int sdfb_comm_unique_id(void* id_out) {
  if (!id_out) return set_error(SDFB_E_INVALID, "null argument");
  NcclApi* a = nccl_api();
  if (!a) return set_error(SDFB_E_CUDA, "NCCL is not available (libnccl.so.2 could not be loaded)");
  const int rc = a->get_unique_id(id_out);
  return rc == 0 ? SDFB_OK : nccl_fail("ncclGetUniqueId", rc);
}

static int comm_new(int world, int rank, int device, sdfb_comm** out, sdfb_comm** c_out) {
  if (!out) return set_error(SDFB_E_INVALID, "null argument");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return set_error(SDFB_E_INVALID, "bad rank %d / world %d", rank, world);
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return set_error(SDFB_E_DEVICE, "no CUDA device visible");
  if (device < 0 || device >= count) return set_error(SDFB_E_INVALID, "device %d out of range (%d devices)", device, count);
  sdfb_comm* c = new (std::nothrow) sdfb_comm();
  if (!c) return set_error(SDFB_E_NOMEM, "out of host memory");
  c->world = world; c->rank = rank; c->device = device;
  *c_out = c;
  return SDFB_OK;
}

int sdfb_comm_create(const void* id, int world, int rank, int device, sdfb_comm** out) {
  if (!id && world > 1) return set_error(SDFB_E_INVALID, "null argument");
  sdfb_comm* c = nullptr;
  int rc = comm_new(world, rank, device, out, &c);
  if (rc) return rc;
  DevGuard g(device);
  if (world > 1) {
    NcclApi* a = nccl_api();
    if (!a) { delete c; return set_error(SDFB_E_CUDA, "NCCL is not available (libnccl.so.2 could not be loaded)"); }
    NcclId nid;
    std::memcpy(nid.b, id, sizeof(nid.b));
    const int nrc = a->comm_init_rank(&c->nccl, world, nid, rank);
    if (nrc != 0) { delete c; return nccl_fail("ncclCommInitRank", nrc); }
    c->own_nccl = true;
  }
  rc = comm_finish_init(c);
  if (rc) { sdfb_comm_destroy(c); return rc; }
  *out = c;
  return SDFB_OK;
}

int sdfb_comm_wrap(void* nccl_comm, int world, int rank, int device, sdfb_comm** out) {
  if (!nccl_comm && world > 1) return set_error(SDFB_E_INVALID, "null argument");
  sdfb_comm* c = nullptr;
  int rc = comm_new(world, rank, device, out, &c);
  if (rc) return rc;
  DevGuard g(device);
  c->nccl = nccl_comm;
  rc = comm_finish_init(c);
  if (rc) { sdfb_comm_destroy(c); return rc; }
  *out = c;
  return SDFB_OK;
}

int sdfb_comm_destroy(sdfb_comm* c) {
  if (!c) return SDFB_OK;
  DevGuard g(c->device);
  cudaDeviceSynchronize();
  release_shared(c);
  for (cudaStream_t s : c->st_push) if (s) cudaStreamDestroy(s);
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  for (cudaEvent_t e : c->ev_done) if (e) cudaEventDestroy(e);
  cudaFree(c->token);
  cudaFree(c->hbuf);
  if (c->own_nccl && c->nccl) {
    NcclApi* a = nccl_api();
    if (a) a->comm_destroy(c->nccl);
  }
  delete c;
  return SDFB_OK;
}

int sdfb_comm_barrier(sdfb_comm* c, void* stream) {
  if (!c) return set_error(SDFB_E_INVALID, "null argument");
  DevGuard g(c->device);
  return sdfb::comm_barrier(c, static_cast<cudaStream_t>(stream));
}

int sdfb_allgather_slabs(sdfb_comm* c, void* full_dev, size_t bytes_per_rank, void* stream) {
  if (!c || (!full_dev && bytes_per_rank > 0)) return set_error(SDFB_E_INVALID, "null argument");
  if (c->world == 1 || bytes_per_rank == 0) return SDFB_OK;
  NcclApi* a = nccl_api();
  if (!a) return set_error(SDFB_E_CUDA, "NCCL is not available (libnccl.so.2 could not be loaded)");
  DevGuard g(c->device);
  char* base = static_cast<char*>(full_dev);
  const int rc = a->all_gather(base + static_cast<size_t>(c->rank) * bytes_per_rank, base, bytes_per_rank, kNcclUint8, c->nccl,
                               static_cast<cudaStream_t>(stream));
  return rc == 0 ? SDFB_OK : nccl_fail("ncclAllGather", rc);
}

}  // extern "C"
