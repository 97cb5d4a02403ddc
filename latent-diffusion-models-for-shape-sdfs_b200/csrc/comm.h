// Internal: the multi-GPU context behind sdfb_comm_* (comm.cu).  Not part of the public C ABI.
#pragma once
#include <cstddef>
#include <vector>

#include <cuda_runtime.h>

struct sdfb_comm {
  int world = 1, rank = 0, device = 0;
  void* nccl = nullptr;          // ncclComm_t
  bool own_nccl = false;         // created by sdfb_comm_create (destroyed with the context) or wrapped
  // symmetric buffer: the same allocation size on every rank, every rank's copy mapped into all others (CUDA IPC)
  void* local = nullptr;
  size_t bytes = 0;
  std::vector<void*> peer;       // [world]; peer[rank] == local
  // pushes run on their own streams (copy engines over NVLink) behind an event of the caller's stream
  static constexpr int kPushStreams = 4;
  cudaStream_t st_push[kPushStreams] = {};
  cudaEvent_t ev_ready = nullptr, ev_done[kPushStreams] = {};
  int* token = nullptr;          // device word for the barrier's all-reduce
  void* hbuf = nullptr;          // device staging for the handle exchange
};

namespace sdfb {
// [offset, offset + bytes) of the local symmetric buffer -> the same range of every peer's copy, after everything queued
// on `after` so far; asynchronous.  join_pushes() makes `st` wait for all pushes issued so far.
int comm_push(sdfb_comm* c, size_t offset, size_t bytes, cudaStream_t after);
int comm_join_pushes(sdfb_comm* c, cudaStream_t st);
int comm_barrier(sdfb_comm* c, cudaStream_t st);
int comm_shared_alloc(sdfb_comm* c, size_t bytes);
}  // namespace sdfb
