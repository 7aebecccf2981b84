// comm.cu — NCCL plumbing for the robot-sharded multi-GPU query
// (databases shard by robot, queries are replicated, per-shard
// result records are merged with ONE ncclAllGather over NVLink).  NCCL is
// dlopen()ed so that single-GPU users carry no dependency on it; when the
// host process already loaded a libnccl.so.2 (e.g. torch's) that copy is used.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>

#include "handle.h"

namespace kml {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
};

static NcclApi* nccl_api(std::string* err) {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) {
#define KML_SYM(name) *(void**)(&api.name) = dlsym(api.lib, "nccl" #name)
      KML_SYM(GetUniqueId);
      KML_SYM(CommInitRank);
      KML_SYM(CommDestroy);
      KML_SYM(AllGather);
      KML_SYM(Broadcast);
      KML_SYM(GetErrorString);
      KML_SYM(CommGetAsyncError);
      KML_SYM(CommAbort);
#undef KML_SYM
    }
  }
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.AllGather) {
    if (err) *err = "libnccl.so.2 not loadable";
    return nullptr;
  }
  return &api;
}

struct Comm {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
};

int comm_nranks(const kml_handle* h) { return h->comm ? h->comm->nranks : 1; }
int comm_rank(const kml_handle* h) { return h->comm ? h->comm->rank : 0; }

// all ranks contribute `bytes` from d_send; d_recv receives nranks*bytes
int comm_allgather(kml_handle* h, const void* d_send, void* d_recv, size_t bytes) {
  if (!h->comm) {
    if (d_send != d_recv)
      KML_CUDA(cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, h->stream));
    return KML_OK;
  }
  NcclApi* api = nccl_api(&h->err);
  if (!api) return KML_ERR_NCCL;
  if (!h->comm->comm) {
    h->err = "communicator was aborted after an error";
    return KML_ERR_NCCL;
  }
  ncclResult_t r = api->AllGather(d_send, d_recv, bytes, ncclChar, h->comm->comm, h->stream);
  if (r != ncclSuccess) {
    h->err = std::string("ncclAllGather: ") + (api->GetErrorString ? api->GetErrorString(r) : "?");
    return KML_ERR_NCCL;
  }
  return KML_OK;
}

// Host wait for a stream that holds a collective: polls the stream like kml_handle::wait_stream,
// and every millisecond asks NCCL for an asynchronous error (a peer that died, a network fault)
// and checks a deadline (KML_NCCL_TIMEOUT_S, default 120 s) — either one aborts the communicator
// and returns KML_ERR_NCCL instead of waiting for ever on a rank that never arrives.
int comm_wait(kml_handle* h) {
  NcclApi* api = h->comm ? nccl_api(nullptr) : nullptr;
  static const double timeout_s = [] {
    const char* e = getenv("KML_NCCL_TIMEOUT_S");
    const double v = e ? atof(e) : 0.0;
    return v > 0.0 ? v : 120.0;
  }();
  struct timespec t0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (unsigned it = 0;; ++it) {
    const cudaError_t e = cudaStreamQuery(h->stream);
    if (e == cudaSuccess) return KML_OK;
    if (e != cudaErrorNotReady) KML_CUDA(e);
    if (api && (it & 63u) == 63u) {
      ncclResult_t ar = ncclSuccess;
      bool bad = api->CommGetAsyncError && api->CommGetAsyncError(h->comm->comm, &ar) == ncclSuccess && ar != ncclSuccess &&
                 ar != ncclInProgress;
      struct timespec t1;
      clock_gettime(CLOCK_MONOTONIC, &t1);
      const double el = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
      if (bad || el > timeout_s) {
        h->err = bad ? std::string("NCCL asynchronous error: ") + (api->GetErrorString ? api->GetErrorString(ar) : "?")
                     : std::string("collective did not complete within KML_NCCL_TIMEOUT_S");
        if (api->CommAbort) api->CommAbort(h->comm->comm);
        h->comm->comm = nullptr;
        return KML_ERR_NCCL;
      }
    }
    struct timespec ts = {0, 20000};
    nanosleep(&ts, nullptr);
  }
}

}  // namespace kml

using namespace kml;

extern "C" {

void kml_comm_destroy_internal(kml_handle* h) {
  if (h->comm) {
    NcclApi* api = nccl_api(nullptr);
    if (api && api->CommDestroy && h->comm->comm) api->CommDestroy(h->comm->comm);
    delete h->comm;
    h->comm = nullptr;
  }
}

int kml_comm_unique_id(void* id_out) {
  if (!id_out) return KML_ERR_ARG;
  static_assert(sizeof(ncclUniqueId) <= KML_UNIQUE_ID_BYTES, "unique id size");
  NcclApi* api = nccl_api(nullptr);
  if (!api) return KML_ERR_NCCL;
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return KML_ERR_NCCL;
  memset(id_out, 0, KML_UNIQUE_ID_BYTES);
  memcpy(id_out, &id, sizeof(id));
  return KML_OK;
}

int kml_comm_init(kml_handle* h, int nranks, int rank, const void* unique_id) {
  if (!h || !unique_id || nranks < 1 || rank < 0 || rank >= nranks) return KML_ERR_ARG;
  try {
    KML_CUDA(cudaSetDevice(h->device));
    NcclApi* api = nccl_api(&h->err);
    if (!api) return KML_ERR_NCCL;
    kml_comm_destroy_internal(h);
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    Comm* c = new Comm();
    c->nranks = nranks;
    c->rank = rank;
    ncclResult_t r = api->CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) {
      h->err = std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(r) : "?");
      delete c;
      return KML_ERR_NCCL;
    }
    h->comm = c;
    return KML_OK;
  } catch (const std::exception& e) {
    h->err = e.what();
    return KML_ERR_CUDA;
  }
}

}  // extern "C"
