// nccl_bcast.cu -- the one collective of the scan-matching path: the submap is broadcast once per
// map version from the rank that owns the map, TOGETHER WITH ITS CELL INDEX, so every other rank adopts the index
// instead of re-sorting the points (replaces nothing in the reference, which is single-process; the payload is what
// mapping_scan_matcher.cc:66-72 rebuilds per frame).  ncclBroadcast runs on the engine stream, so it is ordered with
// the kernels that read the submap and needs no host synchronisation beyond the 64-byte header.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a process that already carries a NCCL (torch bundles its own
// 2.28 next to the system's 2.27) keeps using exactly that one, and hosts that never call these entry points -- the
// single-GPU ROS node -- do not need NCCL installed at all.  Only the five stable C entry points below are used.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "msfl_internal.h"

namespace msfl {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*CommUserRank)(const ncclComm_t, int *) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy this process already uses, if any
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { set_error("NCCL not found (dlopen libnccl.so.2: %s)", dlerror()); return nullptr; }
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
  api.Broadcast = (decltype(api.Broadcast))dlsym(h, "ncclBroadcast");
  api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
  api.CommUserRank = (decltype(api.CommUserRank))dlsym(h, "ncclCommUserRank");
  if (!api.CommUserRank || !api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Broadcast || !api.GroupStart || !api.GroupEnd ||
      !api.GetErrorString) {
    set_error("libnccl.so.2 lacks a required entry point");
    return nullptr;
  }
  api.handle = h;
  return &api;
}

#define MSFL_NCCL_OK(api, expr)                                                                  \
  do {                                                                                           \
    ncclResult_t _r = (expr);                                                                    \
    if (_r != ncclSuccess) {                                                                     \
      msfl::set_error("%s failed: %s (%s:%d)", #expr, (api)->GetErrorString(_r), __FILE__, __LINE__); \
      return MSFL_ERR_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

// what a non-root rank needs to adopt the index: sizes and the two grid headers
struct BcastHeader {
  uint64_t magic;
  uint64_t n[2];
  int32_t dims[2][6];  // nx ny nz ox oy oz
  float inv_edge[2];
};
static_assert(sizeof(BcastHeader) <= 96, "header travels in one small message");
constexpr uint64_t kBcastMagic = 0x6d73666c5f626331ull;  // "msfl_bc1"

}  // namespace msfl

using namespace msfl;

extern "C" {

int msfl_nccl_get_unique_id(unsigned char id[MSFL_NCCL_UNIQUE_ID_BYTES]) {
  static_assert(MSFL_NCCL_UNIQUE_ID_BYTES == sizeof(ncclUniqueId), "msfl.h mirrors NCCL_UNIQUE_ID_BYTES");
  if (!id) { set_error("msfl_nccl_get_unique_id: null id"); return MSFL_ERR_ARG; }
  NcclApi *api = nccl_api();
  if (!api) return MSFL_ERR_CUDA;
  ncclUniqueId u;
  MSFL_NCCL_OK(api, api->GetUniqueId(&u));
  memcpy(id, &u, sizeof u);
  return MSFL_OK;
}

int msfl_nccl_comm_init(msfl_engine *e, const unsigned char id[MSFL_NCCL_UNIQUE_ID_BYTES], int nranks, int rank, void **comm) {
  if (!e || !id || !comm || nranks < 1 || rank < 0 || rank >= nranks) { set_error("msfl_nccl_comm_init: bad argument"); return MSFL_ERR_ARG; }
  *comm = nullptr;
  NcclApi *api = nccl_api();
  if (!api) return MSFL_ERR_CUDA;
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  ncclComm_t c = nullptr;
  MSFL_NCCL_OK(api, api->CommInitRank(&c, nranks, u, rank));
  *comm = (void *)c;
  return MSFL_OK;
}

int msfl_nccl_comm_destroy(void *comm) {
  if (!comm) return MSFL_OK;
  NcclApi *api = nccl_api();
  if (!api) return MSFL_ERR_CUDA;
  MSFL_NCCL_OK(api, api->CommDestroy((ncclComm_t)comm));
  return MSFL_OK;
}

int msfl_bcast_submap(msfl_engine *e, void *nccl_comm, int root) {
  if (!e || !nccl_comm || root < 0) { set_error("msfl_bcast_submap: bad argument"); return MSFL_ERR_ARG; }
  NcclApi *api = nccl_api();
  if (!api) return MSFL_ERR_CUDA;
  ncclComm_t comm = (ncclComm_t)nccl_comm;
  int rank = -1;
  MSFL_NCCL_OK(api, api->CommUserRank(comm, &rank));
  const bool is_root = rank == root;
  if (is_root && !e->has_submap) { set_error("msfl_bcast_submap: the root rank has no submap"); return MSFL_ERR_NOSUBMAP; }
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  int rc;
  if ((rc = e->d_misc.reserve(512))) return rc;
  if ((rc = e->h_misc.reserve(512))) return rc;
  // root: a pageable header (the async copy stages it before returning, so back-to-back calls cannot race on it);
  // other ranks: the pinned landing buffer of the header download
  BcastHeader root_hdr;
  BcastHeader *h = is_root ? &root_hdr : e->h_misc.as<BcastHeader>();
  Submap *cls[2] = {&e->map_corner, &e->map_surf};
  if (is_root) {
    memset(h, 0, sizeof *h);
    h->magic = kBcastMagic;
    for (int c = 0; c < 2; ++c) {
      const GridView &v = cls[c]->view;
      h->n[c] = cls[c]->n;
      const int32_t d[6] = {v.nx, v.ny, v.nz, v.ox, v.oy, v.oz};
      memcpy(h->dims[c], d, sizeof d);
      h->inv_edge[c] = v.inv_edge;
    }
    MSFL_CUDA_OK(cudaMemcpyAsync(e->d_misc.p, h, sizeof *h, cudaMemcpyHostToDevice, st));
  }
  MSFL_NCCL_OK(api, api->Broadcast(e->d_misc.p, e->d_misc.p, sizeof(BcastHeader), ncclUint8, root, comm, st));
  if (!is_root) {
    // the one host synchronisation of the exchange: buffer sizes have to be known before the payload is received
    MSFL_CUDA_OK(cudaMemcpyAsync(h, e->d_misc.p, sizeof *h, cudaMemcpyDeviceToHost, st));
    MSFL_CUDA_OK(cudaStreamSynchronize(st));
    if (h->magic != kBcastMagic) { set_error("msfl_bcast_submap: header mismatch (root rank has no submap or library versions differ)"); return MSFL_ERR_ARG; }
    e->has_submap = false;
    for (int c = 0; c < 2; ++c) {
      const size_t n = (size_t)h->n[c];
      const long long ncell = (long long)h->dims[c][0] * h->dims[c][1] * h->dims[c][2];
      if (n == 0 || n > 0x7fffffffull || ncell <= 0 || ncell > (1ll << 26)) { set_error("msfl_bcast_submap: bad header"); return MSFL_ERR_ARG; }
      if ((rc = cls[c]->orig.reserve(n * 16))) return rc;
      if ((rc = cls[c]->sorted.reserve(n * 16))) return rc;
      if ((rc = cls[c]->cell_start.reserve((size_t)(ncell + 1) * 4))) return rc;
    }
  }
  // payload: per class the points in caller order, the cell-sorted copy and the cell table -- the index is adopted as is
  MSFL_NCCL_OK(api, api->GroupStart());
  for (int c = 0; c < 2; ++c) {
    const size_t n = (size_t)h->n[c];
    const size_t ncell = (size_t)h->dims[c][0] * h->dims[c][1] * h->dims[c][2];
    MSFL_NCCL_OK(api, api->Broadcast(cls[c]->orig.p, cls[c]->orig.p, n * 16, ncclUint8, root, comm, st));
    MSFL_NCCL_OK(api, api->Broadcast(cls[c]->sorted.p, cls[c]->sorted.p, n * 16, ncclUint8, root, comm, st));
    MSFL_NCCL_OK(api, api->Broadcast(cls[c]->cell_start.p, cls[c]->cell_start.p, (ncell + 1) * 4, ncclUint8, root, comm, st));
  }
  MSFL_NCCL_OK(api, api->GroupEnd());
  e->launches += 7;  // NCCL kernels (library)
  if (!is_root) {
    for (int c = 0; c < 2; ++c) {
      Submap &m = *cls[c];
      m.n = (size_t)h->n[c];
      m.view.pts_sorted = m.sorted.as<float4>();
      m.view.pts_orig = m.orig.as<float4>();
      m.view.cell_start = m.cell_start.as<uint32_t>();
      m.view.nx = h->dims[c][0]; m.view.ny = h->dims[c][1]; m.view.nz = h->dims[c][2];
      m.view.ox = h->dims[c][3]; m.view.oy = h->dims[c][4]; m.view.oz = h->dims[c][5];
      m.view.inv_edge = h->inv_edge[c];
      m.view.n = (uint32_t)m.n;
      if ((rc = submap_row_mask(e, m))) return rc;  // derived from the adopted cell table (one small kernel)
    }
    e->has_submap = true;
  }
  return MSFL_OK;
}

}  // extern "C"
