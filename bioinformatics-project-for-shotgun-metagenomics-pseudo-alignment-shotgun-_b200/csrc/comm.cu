// comm.cu -- see comm.h.  Host code only (CUDA runtime calls, NCCL through dlopen).
#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace pa {

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
};

// NCCL is taken from the process when it is already there (torch loads its own libnccl.so.2), else from PA_NCCL_LIB or
// the loader's search path.  One NCCL per process: two copies would each grab their own transports.
NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  static bool ok = false;
  std::call_once(once, [] {
    void* h = nullptr;
    if (const char* e = getenv("PA_NCCL_LIB")) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.lib = h;
    bool all = true;
    auto sym = [&](const char* name) { void* p = dlsym(h, name); if (!p) all = false; return p; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    ok = all;
  });
  return ok ? &api : nullptr;
}

#define PA_NCCL(expr)                                                                                    \
  do {                                                                                                   \
    ncclResult_t r__ = (expr);                                                                           \
    if (r__ != ncclSuccess) {                                                                            \
      pa::set_error("%s failed: %s (%s:%d)", #expr, nccl_api()->GetErrorString(r__), __FILE__, __LINE__); \
      return pa::ST_CUDA;                                                                                \
    }                                                                                                    \
  } while (0)

int32_t new_comm(int n_ranks, int rank, int device, Comm** out) {
  if (!out) { set_error("null output handle"); return ST_INVALID_ARG; }
  *out = nullptr;
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) { set_error("comm: rank %d of %d", rank, n_ranks); return ST_INVALID_ARG; }
  if (n_ranks > 254) { set_error("comm: at most 254 ranks"); return ST_UNSUPPORTED; }
  // device < 0: a host-only communicator (callback transport; host collectives only -- the CPU tests of the summary exchange)
  if (device >= 0) PA_CUDA(cudaSetDevice(device));
  Comm* c = new (std::nothrow) Comm();
  if (!c) { set_error("out of host memory"); return ST_NOMEM; }
  c->n_ranks = n_ranks; c->rank = rank; c->device = device;
  if (device >= 0) {
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e)); delete c; return ST_CUDA; }
  }
  *out = c;
  return ST_OK;
}

int32_t ensure_stage(Comm& c, size_t bytes) {
  if (c.stage_bytes >= bytes) return ST_OK;
  if (c.d_stage) cudaFree(c.d_stage);
  c.d_stage = nullptr; c.stage_bytes = 0;
  const size_t want = std::max<size_t>(bytes + bytes / 2, 1u << 20);
  cudaError_t e = cudaMalloc(&c.d_stage, want);
  if (e != cudaSuccess) { (void)cudaGetLastError(); set_error("comm: cudaMalloc(%zu) failed", want); return ST_NOMEM; }
  c.stage_bytes = want;
  return ST_OK;
}

}  // namespace

Comm::~Comm() {
  if (device >= 0) cudaSetDevice(device);
  release_exchange();
  if (nccl && nccl_api()) nccl_api()->CommDestroy(reinterpret_cast<ncclComm_t>(nccl));
  if (d_stage) cudaFree(d_stage);
  if (stream) cudaStreamDestroy(stream);
}

int32_t comm_unique_id(uint8_t id[128]) {
  NcclApi* api = nccl_api();
  if (!api) { set_error("NCCL is not available (libnccl.so.2 not found; set PA_NCCL_LIB)"); return ST_UNSUPPORTED; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId u;
  PA_NCCL(api->GetUniqueId(&u));
  memcpy(id, &u, 128);
  return ST_OK;
}

int32_t comm_create_nccl(int n_ranks, int rank, const uint8_t id[128], int device, Comm** out) {
  if (device < 0) { set_error("comm: NCCL needs a device"); return ST_INVALID_ARG; }
  NcclApi* api = nccl_api();
  if (!api) { set_error("NCCL is not available (libnccl.so.2 not found; set PA_NCCL_LIB)"); return ST_UNSUPPORTED; }
  Comm* c = nullptr;
  PA_TRY(new_comm(n_ranks, rank, device, &c));
  ncclUniqueId u;
  memcpy(&u, id, 128);
  ncclComm_t nc = nullptr;
  ncclResult_t r = api->CommInitRank(&nc, n_ranks, u, rank);
  if (r != ncclSuccess) { set_error("ncclCommInitRank failed: %s", api->GetErrorString(r)); delete c; return ST_CUDA; }
  c->nccl = nc;
  *out = c;
  return ST_OK;
}

int32_t comm_create_callbacks(int n_ranks, int rank, int device, const pa_comm_callbacks* cb, Comm** out) {
  if (n_ranks > 1 && (!cb || !cb->allgather)) { set_error("comm: an allgather callback is required"); return ST_INVALID_ARG; }
  Comm* c = nullptr;
  PA_TRY(new_comm(n_ranks, rank, device, &c));
  if (cb) c->cb = *cb;
  *out = c;
  return ST_OK;
}

int32_t Comm::allgather_host(const void* in, void* out, size_t bytes) {
  if (bytes == 0) return ST_OK;
  if (single()) { if (out != in) memcpy(out, in, bytes); return ST_OK; }
  if (nccl) {
    NcclApi* api = nccl_api();
    PA_CUDA(cudaSetDevice(device));
    PA_TRY(ensure_stage(*this, bytes * (size_t)(n_ranks + 1)));
    char* d_in = static_cast<char*>(d_stage);
    char* d_out = d_in + bytes;
    PA_CUDA(cudaMemcpyAsync(d_in, in, bytes, cudaMemcpyHostToDevice, stream));
    PA_NCCL(api->AllGather(d_in, d_out, bytes, ncclUint8, reinterpret_cast<ncclComm_t>(nccl), stream));
    PA_CUDA(cudaMemcpyAsync(out, d_out, bytes * (size_t)n_ranks, cudaMemcpyDeviceToHost, stream));
    PA_CUDA(cudaStreamSynchronize(stream));
    return ST_OK;
  }
  const int32_t st = cb.allgather(cb.user, in, out, (uint64_t)bytes);
  if (st != 0) { set_error("comm: the allgather callback failed (%d)", st); return ST_CUDA; }
  return ST_OK;
}

int32_t Comm::allgatherv_host(const void* in, size_t bytes, std::vector<uint64_t>* out_sizes, std::vector<uint8_t>* out) {
  out_sizes->assign((size_t)n_ranks, 0);
  const uint64_t mine = bytes;
  PA_TRY(allgather_host(&mine, out_sizes->data(), 8));
  uint64_t mx = 0, total = 0;
  for (uint64_t v : *out_sizes) { mx = std::max(mx, v); total += v; }
  out->assign(total, 0);
  if (mx == 0) return ST_OK;
  std::vector<uint8_t> pad_in(mx, 0), pad_out(mx * (size_t)n_ranks);
  if (bytes) memcpy(pad_in.data(), in, bytes);
  PA_TRY(allgather_host(pad_in.data(), pad_out.data(), mx));
  uint64_t at = 0;
  for (int r = 0; r < n_ranks; ++r) {
    if ((*out_sizes)[r]) memcpy(out->data() + at, pad_out.data() + (size_t)r * mx, (*out_sizes)[r]);
    at += (*out_sizes)[r];
  }
  return ST_OK;
}

int32_t Comm::barrier() {
  if (single()) return ST_OK;
  std::vector<uint8_t> buf((size_t)n_ranks, 0);
  const uint8_t one = 1;
  return allgather_host(&one, buf.data(), 1);
}

int32_t Comm::allreduce_u64_host(uint64_t* v, size_t n, bool take_min) {
  if (single() || n == 0) return ST_OK;
  std::vector<uint64_t> all(n * (size_t)n_ranks);
  PA_TRY(allgather_host(v, all.data(), n * 8));
  for (size_t i = 0; i < n; ++i) {
    uint64_t acc = all[i];
    for (int r = 1; r < n_ranks; ++r) {
      const uint64_t x = all[(size_t)r * n + i];
      acc = take_min ? std::min(acc, x) : acc + x;
    }
    v[i] = acc;
  }
  return ST_OK;
}

int32_t Comm::allreduce_u64_device(uint64_t* d, size_t n, bool take_min, cudaStream_t s) {
  if (single() || n == 0) return ST_OK;
  if (nccl) {
    PA_NCCL(nccl_api()->AllReduce(d, d, n, ncclUint64, take_min ? ncclMin : ncclSum, reinterpret_cast<ncclComm_t>(nccl), s));
    return ST_OK;
  }
  std::vector<uint64_t> h(n);
  PA_CUDA(cudaMemcpyAsync(h.data(), d, n * 8, cudaMemcpyDeviceToHost, s));
  PA_CUDA(cudaStreamSynchronize(s));
  PA_TRY(allreduce_u64_host(h.data(), n, take_min));
  PA_CUDA(cudaMemcpyAsync(d, h.data(), n * 8, cudaMemcpyHostToDevice, s));
  PA_CUDA(cudaStreamSynchronize(s));
  return ST_OK;
}

int32_t Comm::allgatherv_device_inplace(void* base, const uint64_t* off, cudaStream_t s) {
  if (single()) return ST_OK;
  PA_CUDA(cudaSetDevice(device));
  const char* force = getenv("PA_TABLE_GATHER");   // "ipc" / "nccl": which way the table slices travel
  const bool want_ipc = force ? strcmp(force, "ipc") == 0 : !nccl;
  if (nccl && !want_ipc) {
    NcclApi* api = nccl_api();
    PA_NCCL(api->GroupStart());
    for (int r = 0; r < n_ranks; ++r) {
      const uint64_t n = off[r + 1] - off[r];
      if (!n) continue;
      char* p = static_cast<char*>(base) + off[r];
      PA_NCCL(api->Broadcast(p, p, n, ncclUint8, r, reinterpret_cast<ncclComm_t>(nccl), s));
    }
    PA_NCCL(api->GroupEnd());
    PA_CUDA(cudaStreamSynchronize(s));
    return ST_OK;
  }
  // IPC pull: every rank maps the peers' allocations and copies their segments over NVLink with the copy engines
  PA_CUDA(cudaStreamSynchronize(s));
  cudaIpcMemHandle_t mine;
  int32_t problem = 0;
  if (cudaIpcGetMemHandle(&mine, base) != cudaSuccess) { (void)cudaGetLastError(); memset(&mine, 0, sizeof(mine)); problem = 1; }
  std::vector<uint8_t> blob(72, 0), all((size_t)72 * n_ranks);
  memcpy(blob.data(), &mine, 64);
  blob[64] = (uint8_t)problem;
  PA_TRY(allgather_host(blob.data(), all.data(), 72));   // doubles as the barrier "every slice is complete"
  for (int r = 0; r < n_ranks; ++r) problem |= all[(size_t)r * 72 + 64];
  std::vector<void*> opened;
  if (!problem) {
    std::vector<cudaStream_t> streams;
    for (int r = 0; r < n_ranks && !problem; ++r) {
      if (r == rank || off[r + 1] == off[r]) continue;
      cudaIpcMemHandle_t h;
      memcpy(&h, all.data() + (size_t)r * 72, 64);
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); problem = 1; break; }
      opened.push_back(p);
      cudaStream_t cs = nullptr;
      if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) { problem = 1; break; }
      streams.push_back(cs);
      if (cudaMemcpyAsync(static_cast<char*>(base) + off[r], static_cast<char*>(p) + off[r], off[r + 1] - off[r],
                          cudaMemcpyDeviceToDevice, cs) != cudaSuccess) { (void)cudaGetLastError(); problem = 1; }
    }
    for (cudaStream_t cs : streams) { if (cudaStreamSynchronize(cs) != cudaSuccess) problem = 1; cudaStreamDestroy(cs); }
  }
  // nobody may free or overwrite its slice before every peer has pulled it; the verdict is shared
  std::vector<uint8_t> verdicts((size_t)n_ranks, 0);
  const uint8_t v = (uint8_t)problem;
  PA_TRY(allgather_host(&v, verdicts.data(), 1));
  for (void* p : opened) cudaIpcCloseMemHandle(p);
  for (uint8_t x : verdicts) problem |= x;
  if (problem) { set_error("comm: peer mapping of the table slices failed (CUDA IPC unavailable?)"); return ST_CUDA; }
  return ST_OK;
}

int32_t Comm::allgather_device_strided(void* base, uint64_t stride, cudaStream_t s) {
  if (single() || stride == 0) return ST_OK;
  const char* force = getenv("PA_TABLE_GATHER");
  if (nccl && !(force && strcmp(force, "ipc") == 0)) {
    PA_CUDA(cudaSetDevice(device));
    NcclApi* api = nccl_api();
    PA_NCCL(api->AllGather(static_cast<char*>(base) + (size_t)rank * stride, base, stride, ncclUint8, reinterpret_cast<ncclComm_t>(nccl), s));
    PA_CUDA(cudaStreamSynchronize(s));
    return ST_OK;
  }
  std::vector<uint64_t> off((size_t)n_ranks + 1);
  for (int r = 0; r <= n_ranks; ++r) off[r] = (uint64_t)r * stride;
  return allgatherv_device_inplace(base, off.data(), s);
}

void Comm::release_exchange() {
  for (int r = 0; r < (int)ex.peer_k.size(); ++r) {
    if (r == rank) continue;
    if (ex.peer_k[r]) cudaIpcCloseMemHandle(ex.peer_k[r]);
    if (ex.peer_v[r]) cudaIpcCloseMemHandle(ex.peer_v[r]);
  }
  ex.peer_k.clear(); ex.peer_v.clear();
  if (ex.my_k) cudaFree(ex.my_k);
  if (ex.my_v) cudaFree(ex.my_v);
  ex.my_k = ex.my_v = nullptr;
  ex.cap = 0; ex.ipc = false;
}

int32_t Comm::ensure_exchange(const uint64_t* need) {
  PA_CUDA(cudaSetDevice(device));
  // every rank knows every rank's need and capacity rule, so all ranks take the same decision without a collective
  std::vector<uint64_t> caps((size_t)n_ranks, 0);
  uint64_t my_cap = ex.cap;
  PA_TRY(allgather_host(&my_cap, caps.data(), 8));   // also: every rank has consumed its previous receive buffer
  bool fits = true;
  for (int r = 0; r < n_ranks; ++r) fits = fits && caps[r] > 0 && caps[r] >= need[r];
  if (fits) return ST_OK;
  release_exchange();
  const uint64_t cap = need[rank] + need[rank] / 16 + 4096;
  int32_t problem = 0;
  if (cudaMalloc(&ex.my_k, cap * 8) != cudaSuccess || cudaMalloc(&ex.my_v, cap * 4) != cudaSuccess) {
    (void)cudaGetLastError();
    problem = 2;
  }
  ex.cap = problem ? 0 : cap;
  ex.peer_k.assign((size_t)n_ranks, nullptr);
  ex.peer_v.assign((size_t)n_ranks, nullptr);
  ex.peer_k[rank] = ex.my_k; ex.peer_v[rank] = ex.my_v;
  if (single()) {
    if (problem) { release_exchange(); set_error("exchange: out of device memory (%llu records)", (unsigned long long)cap); return ST_NOMEM; }
    return ST_OK;
  }
  const char* no_ipc = getenv("PA_TEST_NO_IPC");
  bool try_ipc = !(no_ipc && *no_ipc == '1');
  std::vector<uint8_t> blob(136, 0), all((size_t)136 * n_ranks);
  if (!problem && try_ipc) {
    cudaIpcMemHandle_t hk, hv;
    if (cudaIpcGetMemHandle(&hk, ex.my_k) != cudaSuccess || cudaIpcGetMemHandle(&hv, ex.my_v) != cudaSuccess) { (void)cudaGetLastError(); problem = 1; }
    else { memcpy(blob.data(), &hk, 64); memcpy(blob.data() + 64, &hv, 64); }
  } else if (!problem) {
    problem = 1;
  }
  blob[128] = (uint8_t)problem;
  PA_TRY(allgather_host(blob.data(), all.data(), 136));
  int worst = 0;
  for (int r = 0; r < n_ranks; ++r) worst = std::max<int>(worst, all[(size_t)r * 136 + 128]);
  if (worst == 0) {
    for (int r = 0; r < n_ranks && !problem; ++r) {
      if (r == rank) continue;
      cudaIpcMemHandle_t hk, hv;
      memcpy(&hk, all.data() + (size_t)r * 136, 64); memcpy(&hv, all.data() + (size_t)r * 136 + 64, 64);
      if (cudaIpcOpenMemHandle(&ex.peer_k[r], hk, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&ex.peer_v[r], hv, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); problem = 1; }
    }
  }
  std::vector<uint8_t> verdicts((size_t)n_ranks, 0);
  const uint8_t v = (uint8_t)std::max(problem, worst);
  PA_TRY(allgather_host(&v, verdicts.data(), 1));
  for (uint8_t x : verdicts) worst = std::max<int>(worst, x);
  if (worst == 2) { release_exchange(); set_error("exchange: a rank is out of device memory"); return ST_NOMEM; }
  if (worst == 1) {
    // no peer mapping on this system (containers without IPC, no P2P path): all ranks go through NCCL send / recv
    for (int r = 0; r < n_ranks; ++r) {
      if (r == rank) continue;
      if (ex.peer_k[r]) cudaIpcCloseMemHandle(ex.peer_k[r]);
      if (ex.peer_v[r]) cudaIpcCloseMemHandle(ex.peer_v[r]);
      ex.peer_k[r] = ex.peer_v[r] = nullptr;
    }
    if (!nccl) { release_exchange(); set_error("exchange: peer memory cannot be mapped (CUDA IPC) and no NCCL transport is attached"); return ST_UNSUPPORTED; }
    ex.ipc = false;
    return ST_OK;
  }
  ex.ipc = true;
  return ST_OK;
}

int32_t Comm::alltoallv_records(const uint64_t* send_k, const uint32_t* send_v, const uint64_t* send_off, uint64_t* recv_k,
                                uint32_t* recv_v, const uint64_t* recv_off, cudaStream_t s) {
  if (!nccl) { set_error("exchange: no NCCL transport"); return ST_UNSUPPORTED; }
  NcclApi* api = nccl_api();
  ncclComm_t c = reinterpret_cast<ncclComm_t>(nccl);
  PA_NCCL(api->GroupStart());
  for (int r = 0; r < n_ranks; ++r) {
    const uint64_t ns = send_off[r + 1] - send_off[r], nr = recv_off[r + 1] - recv_off[r];
    if (ns) { PA_NCCL(api->Send(send_k + send_off[r], ns, ncclUint64, r, c, s)); PA_NCCL(api->Send(send_v + send_off[r], ns, ncclUint32, r, c, s)); }
    if (nr) { PA_NCCL(api->Recv(recv_k + recv_off[r], nr, ncclUint64, r, c, s)); PA_NCCL(api->Recv(recv_v + recv_off[r], nr, ncclUint32, r, c, s)); }
  }
  PA_NCCL(api->GroupEnd());
  return ST_OK;
}

}  // namespace pa
