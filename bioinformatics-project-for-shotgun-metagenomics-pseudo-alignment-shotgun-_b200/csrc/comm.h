// comm.h -- the communicator behind pa_comm (include/pa_b200.h): one process per GPU on one node.
//
// Small collectives (counts, handles, summaries) go over NCCL -- resolved with dlopen when pa_comm_init is called, so the
// library itself links only the CUDA runtime -- or over a caller-supplied host all-gather (pa_comm_init_callbacks:
// tests over gloo, MPI bootstraps).  The data plane of the index build is peer memory: receive buffers and table
// slices are shared through CUDA IPC and written / read by this library's own kernels and copies over NVLink.
#pragma once
#include "../../include/pa_b200.h"
#include "common.cuh"
#include <vector>

namespace pa {

struct Comm {
  int n_ranks = 1, rank = 0, device = 0;
  void* nccl = nullptr;              // ncclComm_t when the transport is NCCL
  pa_comm_callbacks cb{nullptr, nullptr};
  cudaStream_t stream = nullptr;     // the communicator's own stream (small collectives)
  void* d_stage = nullptr;           // device staging of the host collectives over NCCL
  size_t stage_bytes = 0;

  // receive buffers of the fused exchange and their peer mappings, kept between builds: cudaIpcOpenMemHandle of a
  // multi-GB buffer costs tens of milliseconds
  struct Exchange {
    void* my_k = nullptr;            // uint64 keys   [cap]
    void* my_v = nullptr;            // uint32 values [cap]
    uint64_t cap = 0;                // records
    std::vector<void*> peer_k, peer_v;   // [n_ranks]; own entry = my_k / my_v
    bool ipc = false;                // peers mapped (false: exchange through NCCL send / recv)
  } ex;

  ~Comm();
  bool single() const { return n_ranks <= 1; }
  // out = the `bytes` of every rank, in rank order (host memory)
  int32_t allgather_host(const void* in, void* out, size_t bytes);
  // variable sizes: out_sizes[n_ranks], out = concatenation in rank order
  int32_t allgatherv_host(const void* in, size_t bytes, std::vector<uint64_t>* out_sizes, std::vector<uint8_t>* out);
  int32_t barrier();
  int32_t allreduce_u64_host(uint64_t* v, size_t n, bool take_min);
  // in place on device memory (SUM or MIN over uint64), asynchronous on s for NCCL; returns after completion otherwise
  int32_t allreduce_u64_device(uint64_t* d, size_t n, bool take_min, cudaStream_t s);
  // every rank holds segment [off[rank], off[rank+1]) of `base` (the start of a cudaMalloc allocation of the same size
  // on every rank); afterwards every rank holds all segments.  NCCL broadcasts, or IPC pulls over NVLink.
  int32_t allgatherv_device_inplace(void* base, const uint64_t* off_bytes, cudaStream_t s);
  // the same for equal segments of `stride` bytes (rank r's at r * stride): one ncclAllGather -- every channel and the
  // switch's multicast -- where the grouped broadcasts above reached 236 GB/s per rank between two GPUs
  int32_t allgather_device_strided(void* base, uint64_t stride_bytes, cudaStream_t s);
  // (re)allocates the receive buffers so that every rank can take need[r] records and maps the peers' buffers
  int32_t ensure_exchange(const uint64_t* need /*[n_ranks]*/);
  void release_exchange();
  // exchange through NCCL send / recv (no peer mapping): send_off[n_ranks + 1] into the local send buffers
  int32_t alltoallv_records(const uint64_t* send_k, const uint32_t* send_v, const uint64_t* send_off, uint64_t* recv_k,
                            uint32_t* recv_v, const uint64_t* recv_off, cudaStream_t s);
  bool has_nccl() const { return nccl != nullptr; }
};

int32_t comm_unique_id(uint8_t id[128]);
int32_t comm_create_nccl(int n_ranks, int rank, const uint8_t id[128], int device, Comm** out);
int32_t comm_create_callbacks(int n_ranks, int rank, int device, const pa_comm_callbacks* cb, Comm** out);

}  // namespace pa
