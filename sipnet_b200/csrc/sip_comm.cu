// sip_comm.cu -- multi-GPU layer of the C ABI (include/sipnet_gpu.h "multi-GPU").
//
// Members are independent, so `run` needs no communication: every rank (one rank = one GPU) integrates its share.
// NCCL appears only in the final gather:
//   * log-likelihoods: one all-gather of a double per member, rank-major;
//   * ensemble summaries of a site whose members are spread over ranks: the lockstep select of sip_gsum.cu --
//     per level an all-reduce of 8 KB of histogram per row, plus small all-gathers -- instead of moving the
//     members' values (1 MB per row and rank at 131072 members).
// Two ways to form the team, same code below:
//   * one process per GPU (torchrun, MPI, ...): sipnet_gpu_comm_init_rank() with an id made by rank 0 and
//     broadcast by the caller's launcher;
//   * one process, one host thread, all GPUs: sipnet_gpu_multi_* (ncclCommInitAll, grouped calls).
// NCCL is bound at run time (dlopen of libnccl.so.2): single-GPU use never loads it, and a process that already
// carries an NCCL (e.g. PyTorch's) shares that copy.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sip_gsum.cuh"
#include "sip_handle.h"

using namespace sip;

namespace {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

const NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  const char *names[] = {getenv("SIPNET_GPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void *lib = nullptr;
  for (const char *n : names) {
    if (!n || !*n) continue;
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) return nullptr;
#define SIP_BIND(field, sym)                                              \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(lib, sym));      \
  if (!api.field) return nullptr;
  SIP_BIND(GetUniqueId, "ncclGetUniqueId")
  SIP_BIND(CommInitRank, "ncclCommInitRank")
  SIP_BIND(CommInitAll, "ncclCommInitAll")
  SIP_BIND(CommDestroy, "ncclCommDestroy")
  SIP_BIND(AllReduce, "ncclAllReduce")
  SIP_BIND(AllGather, "ncclAllGather")
  SIP_BIND(GroupStart, "ncclGroupStart")
  SIP_BIND(GroupEnd, "ncclGroupEnd")
  SIP_BIND(GetErrorString, "ncclGetErrorString")
#undef SIP_BIND
  api.lib = lib;
  return &api;
}

#define CUDA_OK(expr)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(SIPNET_GPU_ERR_NO_DEVICE, "%s failed: %s", #expr, cudaGetErrorString(e__));      \
  } while (0)
#define NCCL_OK(expr)                                                                              \
  do {                                                                                             \
    ncclResult_t r__ = (expr);                                                                     \
    if (r__ != ncclSuccess)                                                                        \
      return fail(SIPNET_GPU_ERR_NO_DEVICE, "%s failed: %s", #expr, nccl_api()->GetErrorString(r__)); \
  } while (0)

// one local rank of the team: a handle, its communicator and its scratch for the cross-rank summaries
struct Local {
  sipnet_gpu_handle *h = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0;
  // scratch, sized for `rowsCap` rows
  int64_t rowsCap = 0;
  gs::GsStat *statAll = nullptr;
  double *q2All = nullptr;
  uint32_t *hist = nullptr;
  gs::GsRow *state = nullptr;
  gs::GsTie *tieAll = nullptr;
  uint64_t *emitAll = nullptr;
  int32_t *flags = nullptr;
  double *llAll = nullptr;  // [nranks][maxMembers] log-likelihood gather
  int64_t llCap = 0;
};

}  // namespace

struct sipnet_gpu_comm {
  int nranks = 1;
  std::vector<Local> local;              // the ranks this process drives (1 with one process per GPU)
  std::vector<int64_t> memberCounts;     // [nranks]
  float lastExchangeMs = 0.f, lastSummaryMs = 0.f;
  int lastLevels = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace {

void free_local(Local &l) {
  if (l.h) cudaSetDevice(l.h->device);
  void *ptrs[] = {l.statAll, l.q2All, l.hist, l.state, l.tieAll, l.emitAll, l.flags, l.llAll};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  l.statAll = nullptr;
  l.q2All = nullptr;
  l.hist = nullptr;
  l.state = nullptr;
  l.tieAll = nullptr;
  l.emitAll = nullptr;
  l.flags = nullptr;
  l.llAll = nullptr;
  l.rowsCap = 0;
  l.llCap = 0;
}

// collectives over the team; with one rank they are no-ops (local data already sits at index `rank` = 0)
struct Group {
  const sipnet_gpu_comm *c;
  bool grouped;
  explicit Group(const sipnet_gpu_comm *cc) : c(cc), grouped(cc->nranks > 1 && cc->local.size() > 1) {}
  int begin() const {
    if (grouped) NCCL_OK(nccl_api()->GroupStart());
    return 0;
  }
  int end() const {
    if (grouped) NCCL_OK(nccl_api()->GroupEnd());
    return 0;
  }
};

int all_gather_bytes(sipnet_gpu_comm *c, size_t bytesPerRank, void *(*buf)(Local &)) {
  if (c->nranks == 1) return 0;
  const Group g(c);
  if (int rc = g.begin()) return rc;
  for (Local &l : c->local) {
    CUDA_OK(cudaSetDevice(l.h->device));
    char *base = static_cast<char *>(buf(l));
    NCCL_OK(nccl_api()->AllGather(base + (size_t)l.rank * bytesPerRank, base, bytesPerRank, ncclUint8, l.comm, l.h->stream));
  }
  return g.end();
}

int all_reduce_hist(sipnet_gpu_comm *c, size_t words) {
  if (c->nranks == 1) return 0;
  const Group g(c);
  if (int rc = g.begin()) return rc;
  for (Local &l : c->local) {
    CUDA_OK(cudaSetDevice(l.h->device));
    NCCL_OK(nccl_api()->AllReduce(l.hist, l.hist, words, ncclUint32, ncclSum, l.comm, l.h->stream));
  }
  return g.end();
}

int ensure_scratch(sipnet_gpu_comm *c, Local &l, int64_t rows) {
  if (rows <= l.rowsCap) return 0;
  const int R = c->nranks;
  sipnet_gpu_handle *h = l.h;
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  void *keepLl = l.llAll;
  const int64_t keepCap = l.llCap;
  l.llAll = nullptr;
  free_local(l);
  l.llAll = static_cast<double *>(keepLl);
  l.llCap = keepCap;
  CUDA_OK(cudaMalloc((void **)&l.statAll, (size_t)R * rows * sizeof(gs::GsStat)));
  CUDA_OK(cudaMalloc((void **)&l.q2All, (size_t)R * rows * sizeof(double)));
  CUDA_OK(cudaMalloc((void **)&l.hist, (size_t)rows * gs::kGsHistWords * sizeof(uint32_t)));
  CUDA_OK(cudaMalloc((void **)&l.state, (size_t)rows * sizeof(gs::GsRow)));
  CUDA_OK(cudaMalloc((void **)&l.tieAll, (size_t)R * rows * gs::kGsMaxStat * sizeof(gs::GsTie)));
  CUDA_OK(cudaMalloc((void **)&l.emitAll, (size_t)R * rows * gs::kGsMaxStat * gs::kGsEmit * sizeof(uint64_t)));
  CUDA_OK(cudaMalloc((void **)&l.flags, 16));
  l.rowsCap = rows;
  return 0;
}

// The summaries of the last run range over the WHOLE team, into every local handle's mean / var / quant buffers.
int team_summaries(sipnet_gpu_comm *c) {
  Local &l0 = c->local[0];
  sipnet_gpu_handle *h0 = l0.h;
  const int64_t n = h0->lastEnd - h0->lastBegin;
  if (n <= 0) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "no run range to summarise");
  const int ncols = (int)h0->summaryCols.size();
  if (ncols <= 0 || ncols > gs::kGsMaxCols)
    return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "cross-rank summaries take 1..%d summary columns", gs::kGsMaxCols);
  if (!h0->mean && !h0->quant) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "no summary output was requested at init");
  const int nqTotal = h0->quant ? (int)h0->quantiles.size() : 0;
  const int64_t rows = (int64_t)h0->nsites * ncols * n;
  for (Local &l : c->local) {
    sipnet_gpu_handle *h = l.h;
    if (h->lastEnd - h->lastBegin != n || h->nsites != h0->nsites || (int)h->summaryCols.size() != ncols)
      return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "the ranks of a team must hold the same sites, columns and run range");
    if (int rc = ensure_scratch(c, l, rows)) return rc;
  }
  std::vector<gs::GsArgs> args(c->local.size());
  for (size_t i = 0; i < c->local.size(); ++i) {
    Local &l = c->local[i];
    sipnet_gpu_handle *h = l.h;
    gs::GsArgs &a = args[i];
    memset(&a, 0, sizeof a);
    a.out = h->out;
    a.ld = h->ld;
    a.nsteps = n;
    a.sites = h->sites;
    a.nsites = (int32_t)h->nsites;
    a.ncols = ncols;
    for (int k = 0; k < ncols; ++k) a.colSlot[k] = h->colSlot[h->summaryCols[(size_t)k]];
    a.nranks = c->nranks;
    a.rank = l.rank;
    a.wantMoments = h->mean != nullptr;
    a.statAll = l.statAll;
    a.q2All = l.q2All;
    a.hist = l.hist;
    a.state = l.state;
    a.tieAll = l.tieAll;
    a.emitAll = l.emitAll;
    a.flags = l.flags;
    a.mean = h->mean;
    a.var = h->var;
    a.quant = h->quant;
    a.nqTotal = nqTotal;
  }
  c->lastLevels = 0;
  // quantiles in groups of kGsMaxQ; the moments ride with the first group
  for (int q0 = 0; q0 < std::max(nqTotal, 1); q0 += gs::kGsMaxQ) {
    const int nq = std::min(gs::kGsMaxQ, nqTotal - q0);
    for (size_t i = 0; i < c->local.size(); ++i) {
      Local &l = c->local[i];
      gs::GsArgs &a = args[i];
      a.nq = std::max(nq, 0);
      a.q0 = q0;
      for (int k = 0; k < a.nq; ++k) a.probs[k] = l.h->quantiles[(size_t)(q0 + k)];
      a.wantMoments = (l.h->mean != nullptr && q0 == 0) ? 1 : 0;
      CUDA_OK(cudaSetDevice(l.h->device));
      cudaError_t e = gs::launch_pass0(a, l.h->stream);
      l.h->launches++;
      if (e != cudaSuccess) return fail(SIPNET_GPU_ERR_NO_DEVICE, "summary launch failed: %s", cudaGetErrorString(e));
    }
    if (int rc = all_gather_bytes(c, (size_t)rows * sizeof(gs::GsStat), [](Local &l) -> void * { return l.statAll; })) return rc;
    if (args[0].nq > 0)
      if (int rc = all_reduce_hist(c, (size_t)rows * gs::kGsHistWords)) return rc;
    for (int level = 1; level <= gs::kGsLevels; ++level) {
      for (size_t i = 0; i < c->local.size(); ++i) {
        Local &l = c->local[i];
        CUDA_OK(cudaSetDevice(l.h->device));
        CUDA_OK(cudaMemsetAsync(l.flags, 0, 16, l.h->stream));
        cudaError_t e = gs::launch_level(args[i], level, l.h->stream);
        l.h->launches++;
        if (e != cudaSuccess) return fail(SIPNET_GPU_ERR_NO_DEVICE, "summary launch failed: %s", cudaGetErrorString(e));
      }
      // the flag is a function of exchanged data only: every rank reads the same value
      int32_t flag = 0;
      {
        Local &l = c->local[0];
        CUDA_OK(cudaSetDevice(l.h->device));
        CUDA_OK(cudaMemcpyAsync(&flag, l.flags, sizeof flag, cudaMemcpyDeviceToHost, l.h->stream));
        CUDA_OK(cudaStreamSynchronize(l.h->stream));
      }
      if (level == 1 && args[0].wantMoments)
        if (int rc = all_gather_bytes(c, (size_t)rows * sizeof(double), [](Local &l) -> void * { return l.q2All; })) return rc;
      if (!flag) break;
      c->lastLevels = std::max(c->lastLevels, level);
      if (int rc = all_reduce_hist(c, (size_t)rows * gs::kGsHistWords)) return rc;
      if (int rc = all_gather_bytes(c, (size_t)rows * gs::kGsMaxStat * sizeof(gs::GsTie), [](Local &l) -> void * { return l.tieAll; }))
        return rc;
    }
    if (args[0].nq > 0) {
      for (size_t i = 0; i < c->local.size(); ++i) {
        Local &l = c->local[i];
        CUDA_OK(cudaSetDevice(l.h->device));
        cudaError_t e = gs::launch_emit(args[i], l.h->stream);
        l.h->launches++;
        if (e != cudaSuccess) return fail(SIPNET_GPU_ERR_NO_DEVICE, "summary launch failed: %s", cudaGetErrorString(e));
      }
      if (int rc = all_gather_bytes(c, (size_t)rows * gs::kGsMaxStat * gs::kGsEmit * sizeof(uint64_t),
                                    [](Local &l) -> void * { return l.emitAll; }))
        return rc;
    }
    for (size_t i = 0; i < c->local.size(); ++i) {
      Local &l = c->local[i];
      CUDA_OK(cudaSetDevice(l.h->device));
      cudaError_t e = gs::launch_finish(args[i], l.h->stream);
      l.h->launches++;
      if (e != cudaSuccess) return fail(SIPNET_GPU_ERR_NO_DEVICE, "summary launch failed: %s", cudaGetErrorString(e));
    }
  }
  for (Local &l : c->local) l.h->summariesValid = true;
  return 0;
}

}  // namespace

// ---- C ABI: one process per GPU ---------------------------------------------------------------------------------
extern "C" int sipnet_gpu_comm_unique_id(void *id) {
  if (!id) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "NULL id");
  const NcclApi *api = nccl_api();
  if (!api) return fail(SIPNET_GPU_ERR_NO_DEVICE, "NCCL (libnccl.so.2) could not be loaded: %s", dlerror());
  ncclUniqueId uid;
  NCCL_OK(api->GetUniqueId(&uid));
  static_assert(sizeof(uid) == SIPNET_GPU_COMM_ID_BYTES, "id size");
  memcpy(id, &uid, sizeof uid);
  return 0;
}

static int finish_comm(sipnet_gpu_comm *c) {
  // member counts of all ranks (for the log-likelihood gather's layout)
  c->memberCounts.assign((size_t)c->nranks, 0);
  if (c->nranks == 1) {
    c->memberCounts[0] = c->local[0].h->nmembers;
    return 0;
  }
  std::vector<int64_t *> dev(c->local.size(), nullptr);
  const Group g(c);
  for (size_t i = 0; i < c->local.size(); ++i) {
    Local &l = c->local[i];
    CUDA_OK(cudaSetDevice(l.h->device));
    CUDA_OK(cudaMalloc((void **)&dev[i], (size_t)c->nranks * sizeof(int64_t)));
    const int64_t mine = l.h->nmembers;
    CUDA_OK(cudaMemcpyAsync(dev[i] + l.rank, &mine, sizeof mine, cudaMemcpyHostToDevice, l.h->stream));
    CUDA_OK(cudaStreamSynchronize(l.h->stream));
  }
  if (int rc = g.begin()) return rc;
  for (size_t i = 0; i < c->local.size(); ++i) {
    Local &l = c->local[i];
    CUDA_OK(cudaSetDevice(l.h->device));
    NCCL_OK(nccl_api()->AllGather(dev[i] + l.rank, dev[i], sizeof(int64_t), ncclUint8, l.comm, l.h->stream));
  }
  if (int rc = g.end()) return rc;
  for (size_t i = 0; i < c->local.size(); ++i) {
    Local &l = c->local[i];
    CUDA_OK(cudaSetDevice(l.h->device));
    CUDA_OK(cudaMemcpyAsync(c->memberCounts.data(), dev[i], (size_t)c->nranks * sizeof(int64_t), cudaMemcpyDeviceToHost,
                            l.h->stream));
    CUDA_OK(cudaStreamSynchronize(l.h->stream));
    cudaFree(dev[i]);
  }
  return 0;
}

extern "C" int sipnet_gpu_comm_init_rank(sipnet_gpu_handle *h, int32_t nranks, int32_t rank, const void *id,
                                         sipnet_gpu_comm **out) {
  if (out) *out = nullptr;
  if (!h || !out) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "NULL handle or comm pointer");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "bad rank %d of %d", rank, nranks);
  sipnet_gpu_comm *c = new sipnet_gpu_comm();
  c->nranks = nranks;
  c->local.resize(1);
  c->local[0].h = h;
  c->local[0].rank = rank;
  if (nranks > 1) {
    if (!id) {
      delete c;
      return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "NULL id");
    }
    const NcclApi *api = nccl_api();
    if (!api) {
      delete c;
      return fail(SIPNET_GPU_ERR_NO_DEVICE, "NCCL (libnccl.so.2) could not be loaded: %s", dlerror());
    }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    cudaSetDevice(h->device);
    ncclResult_t r = api->CommInitRank(&c->local[0].comm, nranks, uid, rank);
    if (r != ncclSuccess) {
      delete c;
      return fail(SIPNET_GPU_ERR_NO_DEVICE, "ncclCommInitRank failed: %s", api->GetErrorString(r));
    }
  }
  if (int rc = finish_comm(c)) {
    sipnet_gpu_comm_destroy(c);
    return rc;
  }
  *out = c;
  return 0;
}

extern "C" void sipnet_gpu_comm_destroy(sipnet_gpu_comm *c) {
  if (!c) return;
  for (Local &l : c->local) {
    if (l.h) {
      cudaSetDevice(l.h->device);
      cudaStreamSynchronize(l.h->stream);
    }
    free_local(l);
    if (l.comm) nccl_api()->CommDestroy(l.comm);
  }
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  delete c;
}

extern "C" int32_t sipnet_gpu_comm_nranks(const sipnet_gpu_comm *c) { return c ? c->nranks : 0; }

extern "C" int sipnet_gpu_comm_summaries(sipnet_gpu_comm *c) {
  if (!c) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "NULL comm");
  return team_summaries(c);
}

extern "C" int32_t sipnet_gpu_comm_last_levels(const sipnet_gpu_comm *c) { return c ? c->lastLevels : 0; }

// log-likelihoods of every member of the team, rank-major, into each local rank's gather buffer; returns the
// device pointer of local rank `i`'s copy through `dev` (or copies to `dst` on the host when given)
static int team_loglik(sipnet_gpu_comm *c, double *dst, size_t bytes, int what) {
  int64_t total = 0, width = 0;
  for (int64_t m : c->memberCounts) {
    total += m;
    width = std::max(width, m);
  }
  if (dst && bytes != (size_t)total * sizeof(double))
    return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "log-likelihood gather: buffer is %zu bytes, expected %zu", bytes,
                (size_t)total * sizeof(double));
  for (Local &l : c->local) {
    sipnet_gpu_handle *h = l.h;
    const double *src = what == SIPNET_GPU_GATHER_LOGLIK ? h->loglik : h->loglikN;
    if (!src) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "log-likelihood output was not requested at init");
    CUDA_OK(cudaSetDevice(h->device));
    if (l.llCap < width) {
      if (l.llAll) cudaFree(l.llAll);
      l.llAll = nullptr;
      CUDA_OK(cudaMalloc((void **)&l.llAll, (size_t)c->nranks * width * sizeof(double)));
      l.llCap = width;
    }
    CUDA_OK(cudaMemcpyAsync(l.llAll + (size_t)l.rank * width, src, (size_t)h->nmembers * sizeof(double),
                            cudaMemcpyDeviceToDevice, h->stream));
  }
  if (int rc = all_gather_bytes(c, (size_t)width * sizeof(double), [](Local &l) -> void * { return l.llAll; })) return rc;
  if (dst) {
    Local &l = c->local[0];
    CUDA_OK(cudaSetDevice(l.h->device));
    int64_t off = 0;
    for (int q = 0; q < c->nranks; ++q) {
      CUDA_OK(cudaMemcpyAsync(dst + off, l.llAll + (size_t)q * width, (size_t)c->memberCounts[(size_t)q] * sizeof(double),
                              cudaMemcpyDeviceToHost, l.h->stream));
      off += c->memberCounts[(size_t)q];
    }
    CUDA_OK(cudaStreamSynchronize(l.h->stream));
  }
  // every local stream has finished its part before the caller reuses the handles
  for (Local &l : c->local) {
    CUDA_OK(cudaSetDevice(l.h->device));
    CUDA_OK(cudaStreamSynchronize(l.h->stream));
  }
  return 0;
}

extern "C" int sipnet_gpu_comm_gather_loglik(sipnet_gpu_comm *c, double *dst, size_t bytes) {
  if (!c || !dst) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "NULL comm or destination");
  return team_loglik(c, dst, bytes, SIPNET_GPU_GATHER_LOGLIK);
}

extern "C" int sipnet_gpu_comm_member_counts(const sipnet_gpu_comm *c, int64_t *counts) {
  if (!c || !counts) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "NULL argument");
  for (int q = 0; q < c->nranks; ++q) counts[q] = c->memberCounts[(size_t)q];
  return 0;
}

// ---- C ABI: one process, all GPUs ---------------------------------------------------------------------------------
// The configuration is the single-GPU one; members are partitioned like this:
//   * at least as many sites as devices: whole sites per device (forcing is not replicated), contiguous site ranges;
//   * fewer sites than devices: every device gets every site and an even, contiguous share of each site's members;
//     the summaries of such split sites go through the team select above.
struct sipnet_gpu_multi {
  sipnet_gpu_comm *comm = nullptr;
  std::vector<sipnet_gpu_handle *> handles;
  bool split = false;  // sites are split over devices (else whole sites per device)
  int64_t nmembers = 0, nsites = 0, nsummary = 0, nquant = 0;
  uint32_t outputs = 0;
  // per device: for whole-site partitions the global member / site range; for split sites per site the global
  // offset of the device's share
  struct Part {
    int64_t site0 = 0, site1 = 0, member0 = 0, member1 = 0;
    std::vector<int64_t> siteGlobal0, siteLocal0, siteCount;  // split: per site
  };
  std::vector<Part> parts;
};

extern "C" void sipnet_gpu_multi_destroy(sipnet_gpu_multi *m) {
  if (!m) return;
  if (m->comm) {
    for (Local &l : m->comm->local) l.h = l.h;  // handles are destroyed below, after the communicators
    sipnet_gpu_comm_destroy(m->comm);
  }
  for (sipnet_gpu_handle *h : m->handles) sipnet_gpu_destroy(h);
  delete m;
}

extern "C" int sipnet_gpu_multi_init(const sipnet_gpu_config *cfg, int32_t ndevices, const int32_t *devices,
                                     sipnet_gpu_multi **out) {
  if (out) *out = nullptr;
  if (!cfg || !out) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "NULL config or handle pointer");
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible <= 0)
    return fail(SIPNET_GPU_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
  std::vector<int> devs;
  if (ndevices <= 0) {
    for (int d = 0; d < visible; ++d) devs.push_back(d);
  } else {
    for (int i = 0; i < ndevices; ++i) devs.push_back(devices ? devices[i] : i);
  }
  for (int d : devs)
    if (d < 0 || d >= visible) return fail(SIPNET_GPU_ERR_NO_DEVICE, "device %d not present", d);
  if (cfg->nsites <= 0 || !cfg->sites || cfg->nmembers <= 0 || !cfg->params)
    return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "no sites / members");
  if (cfg->stream) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "a caller stream cannot be shared by several devices");
  // never more devices than members
  while ((int64_t)devs.size() > cfg->nmembers) devs.pop_back();
  const int D = (int)devs.size();

  sipnet_gpu_multi *m = new sipnet_gpu_multi();
  m->nmembers = cfg->nmembers;
  m->nsites = cfg->nsites;
  m->nsummary = cfg->n_summary_cols;
  m->nquant = cfg->n_quantiles;
  m->outputs = cfg->outputs;
  m->split = cfg->nsites < D;
  m->parts.resize((size_t)D);
  std::vector<int32_t> memberSite((size_t)cfg->nmembers, 0);
  if (cfg->member_site)
    for (int64_t i = 0; i < cfg->nmembers; ++i) memberSite[(size_t)i] = cfg->member_site[i];
  // first member and count of every site (validated again by sipnet_gpu_init)
  std::vector<int64_t> siteM0((size_t)cfg->nsites, 0), siteCnt((size_t)cfg->nsites, 0);
  for (int64_t i = 0; i < cfg->nmembers; ++i) {
    const int32_t s = memberSite[(size_t)i];
    if (s < 0 || s >= cfg->nsites || (i > 0 && s < memberSite[(size_t)i - 1])) {
      delete m;
      return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "member_site must be non-decreasing and in range");
    }
    if (siteCnt[(size_t)s]++ == 0) siteM0[(size_t)s] = i;
  }

  // the same mean-NPP ring capacity on every device (sipnet_gpu_init sizes it from the shortest step it sees)
  int32_t ringSlots = cfg->ring_slots;
  if (ringSlots == 0) {
    double minLen = 1e300;
    for (int64_t s = 0; s < cfg->nsites; ++s)
      for (int64_t t = 0; t < cfg->sites[s].nsteps && cfg->sites[s].length; ++t)
        if (cfg->sites[s].length[t] > 0) minLen = std::min(minLen, cfg->sites[s].length[t]);
    const double bound = std::floor(kMeanNppDays / minLen) + 3.0;
    ringSlots = (int32_t)std::min<double>(kRingMax, std::max(2.0, bound));
  }

  for (int d = 0; d < D; ++d) {
    sipnet_gpu_multi::Part &p = m->parts[(size_t)d];
    sipnet_gpu_config sub = *cfg;
    sub.device = devs[(size_t)d];
    sub.ring_slots = ringSlots;
    std::vector<int32_t> subSite;
    std::vector<double> subParams;
    std::vector<int64_t> pick;  // global member indices of this device, ascending
    if (!m->split) {
      p.site0 = (cfg->nsites * d) / D;
      p.site1 = (cfg->nsites * (d + 1)) / D;
      for (int64_t i = 0; i < cfg->nmembers; ++i)
        if (memberSite[(size_t)i] >= p.site0 && memberSite[(size_t)i] < p.site1) pick.push_back(i);
      p.member0 = pick.empty() ? 0 : pick.front();
      p.member1 = pick.empty() ? 0 : pick.back() + 1;
      sub.nsites = p.site1 - p.site0;
      sub.sites = cfg->sites + p.site0;
      for (int64_t i : pick) subSite.push_back(memberSite[(size_t)i] - (int32_t)p.site0);
    } else {
      p.site0 = 0;
      p.site1 = cfg->nsites;
      int64_t local = 0;
      for (int64_t s = 0; s < cfg->nsites; ++s) {
        const int64_t lo = (siteCnt[(size_t)s] * d) / D, hi = (siteCnt[(size_t)s] * (d + 1)) / D;
        p.siteGlobal0.push_back(siteM0[(size_t)s] + lo);
        p.siteLocal0.push_back(local);
        p.siteCount.push_back(hi - lo);
        for (int64_t i = lo; i < hi; ++i) {
          pick.push_back(siteM0[(size_t)s] + i);
          subSite.push_back((int32_t)s);
        }
        local += hi - lo;
      }
    }
    if (pick.empty()) {
      sipnet_gpu_multi_destroy(m);
      return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "device %d would get no members", d);
    }
    const int64_t nm = (int64_t)pick.size();
    subParams.resize((size_t)SIPNET_GPU_NPARAMS * (size_t)nm);
    for (int k = 0; k < SIPNET_GPU_NPARAMS; ++k)
      for (int64_t j = 0; j < nm; ++j)
        subParams[(size_t)k * (size_t)nm + (size_t)j] = cfg->params[(size_t)k * (size_t)cfg->params_ld + (size_t)pick[(size_t)j]];
    sub.nmembers = nm;
    sub.member_site = subSite.data();
    sub.params = subParams.data();
    sub.params_ld = nm;
    sipnet_gpu_handle *h = nullptr;
    const int rc = sipnet_gpu_init(&sub, &h);
    if (rc) {
      sipnet_gpu_multi_destroy(m);
      return rc;
    }
    m->handles.push_back(h);
  }

  // the team: needed when sites are split (summaries) or log-likelihoods are gathered through NCCL; one device
  // needs none
  sipnet_gpu_comm *c = new sipnet_gpu_comm();
  c->nranks = D;
  c->local.resize((size_t)D);
  for (int d = 0; d < D; ++d) {
    c->local[(size_t)d].h = m->handles[(size_t)d];
    c->local[(size_t)d].rank = d;
  }
  m->comm = c;
  if (D > 1) {
    const NcclApi *api = nccl_api();
    if (!api) {
      sipnet_gpu_multi_destroy(m);
      return fail(SIPNET_GPU_ERR_NO_DEVICE, "NCCL (libnccl.so.2) could not be loaded: %s", dlerror());
    }
    std::vector<ncclComm_t> comms((size_t)D, nullptr);
    ncclResult_t r = api->CommInitAll(comms.data(), D, devs.data());
    if (r != ncclSuccess) {
      sipnet_gpu_multi_destroy(m);
      return fail(SIPNET_GPU_ERR_NO_DEVICE, "ncclCommInitAll failed: %s", api->GetErrorString(r));
    }
    for (int d = 0; d < D; ++d) c->local[(size_t)d].comm = comms[(size_t)d];
  }
  if (int rc = finish_comm(c)) {
    sipnet_gpu_multi_destroy(m);
    return rc;
  }
  *out = m;
  return 0;
}

extern "C" int32_t sipnet_gpu_multi_ndevices(const sipnet_gpu_multi *m) { return m ? (int32_t)m->handles.size() : 0; }
extern "C" sipnet_gpu_handle *sipnet_gpu_multi_handle(sipnet_gpu_multi *m, int32_t i) {
  return (m && i >= 0 && i < (int32_t)m->handles.size()) ? m->handles[(size_t)i] : nullptr;
}

extern "C" int sipnet_gpu_multi_run(sipnet_gpu_multi *m, int64_t step_begin, int64_t step_end) {
  if (!m) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "NULL handle");
  for (sipnet_gpu_handle *h : m->handles) {  // asynchronous launches: all devices run side by side
    // a device whose sites are all shorter than the range runs what it has
    const int64_t e = std::min<int64_t>(step_end, h->maxSteps);
    const int64_t b = std::min<int64_t>(step_begin, e);
    if (int rc = sipnet_gpu_run(h, b, e)) return rc;
  }
  return 0;
}

extern "C" int sipnet_gpu_multi_reset(sipnet_gpu_multi *m) {
  if (!m) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "NULL handle");
  for (sipnet_gpu_handle *h : m->handles)
    if (int rc = sipnet_gpu_reset(h)) return rc;
  return 0;
}

extern "C" int sipnet_gpu_multi_sync(sipnet_gpu_multi *m) {
  if (!m) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "NULL handle");
  for (sipnet_gpu_handle *h : m->handles)
    if (int rc = sipnet_gpu_sync(h)) return rc;
  return 0;
}

// steps of the last run range over all devices (a device whose sites are shorter ran fewer)
static int64_t multi_range_steps(const sipnet_gpu_multi *m) {
  int64_t n = 0;
  for (const sipnet_gpu_handle *h : m->handles) n = std::max(n, h->lastEnd - h->lastBegin);
  return n;
}

// per-member outputs: `rows` rows of M elements of `elem` bytes; per-step outputs have rows = columns x steps
static bool member_output_shape(const sipnet_gpu_multi *m, int what, size_t *cols, bool *perStep, size_t *elem) {
  const sipnet_gpu_handle *h0 = m->handles[0];
  *perStep = false;
  *elem = 8;
  switch (what) {
    case SIPNET_GPU_GATHER_FULL: *cols = SIPNET_GPU_NOUT; *perStep = true; return (h0->outputs & SIPNET_GPU_OUT_FULL) != 0;
    case SIPNET_GPU_GATHER_DEBUG: *cols = SIPNET_GPU_NDEBUG; *perStep = true; return h0->dbg != nullptr;
    case SIPNET_GPU_GATHER_BALANCE: *cols = SIPNET_GPU_NBALANCE; *perStep = true; return h0->dbg != nullptr;
    case SIPNET_GPU_GATHER_LOGLIK:
    case SIPNET_GPU_GATHER_LOGLIK_N: *cols = 1; return h0->loglik != nullptr;
    case SIPNET_GPU_GATHER_STATUS: *cols = 1; *elem = 4; return true;
    case SIPNET_GPU_GATHER_STATE: *cols = SIPNET_GPU_NSTATE; return true;
    case SIPNET_GPU_GATHER_RING_VALUES:
    case SIPNET_GPU_GATHER_RING_WEIGHTS: *cols = (size_t)h0->ringCap; return true;
    case SIPNET_GPU_GATHER_EVENT_COUNTS: *cols = 1; *elem = 4; return h0->recCount != nullptr;
    case SIPNET_GPU_GATHER_EVENT_RECORDS:
      *cols = 1;
      *elem = (size_t)h0->maxRecs * sizeof(sipnet_gpu_event_record);
      return h0->recs != nullptr;
    default: return false;
  }
}

extern "C" size_t sipnet_gpu_multi_gather_bytes(const sipnet_gpu_multi *m, int what) {
  if (!m || m->handles.empty()) return 0;
  const sipnet_gpu_handle *h0 = m->handles[0];
  const size_t n = (size_t)multi_range_steps(m);
  const size_t ns = h0->summaryCols.size();
  switch (what) {
    case SIPNET_GPU_GATHER_MEAN:
    case SIPNET_GPU_GATHER_VARIANCE: return h0->mean ? (size_t)m->nsites * ns * n * 8 : 0;
    case SIPNET_GPU_GATHER_QUANTILES: return h0->quant ? (size_t)m->nsites * ns * h0->quantiles.size() * n * 8 : 0;
    default: break;
  }
  size_t cols = 0, elem = 8;
  bool perStep = false;
  if (!member_output_shape(m, what, &cols, &perStep, &elem)) return 0;
  return cols * (perStep ? n : 1) * (size_t)m->nmembers * elem;
}

extern "C" int sipnet_gpu_multi_gather(sipnet_gpu_multi *m, int what, void *dst, size_t bytes) {
  if (!m || !dst) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "NULL handle or destination");
  const size_t need = sipnet_gpu_multi_gather_bytes(m, what);
  if (need == 0) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "gather(%d): output was not requested at init or nothing has run", what);
  if (bytes != need) return fail(SIPNET_GPU_ERR_BAD_ARGUMENT, "gather(%d): buffer is %zu bytes, expected %zu", what, bytes, need);
  const size_t D = m->handles.size();
  const size_t n = (size_t)multi_range_steps(m);
  const bool summary = what == SIPNET_GPU_GATHER_MEAN || what == SIPNET_GPU_GATHER_VARIANCE || what == SIPNET_GPU_GATHER_QUANTILES;
  if (summary) {
    if (m->split) {  // every device ends with the team's result: take device 0's
      if (!m->handles[0]->summariesValid)
        if (int rc = team_summaries(m->comm)) return rc;
      return sipnet_gpu_gather(m->handles[0], what, dst, bytes);
    }
    // whole sites per device: local summaries; rows [site][col]([q])[n], a shorter device's rows are NaN-padded
    const sipnet_gpu_handle *h0 = m->handles[0];
    const size_t per = h0->summaryCols.size() * (what == SIPNET_GPU_GATHER_QUANTILES ? h0->quantiles.size() : 1);
    double *out = static_cast<double *>(dst);
    std::vector<double> tmp;
    for (size_t d = 0; d < D; ++d) {
      sipnet_gpu_handle *h = m->handles[d];
      const size_t nd = (size_t)(h->lastEnd - h->lastBegin);
      const size_t rowsD = (size_t)h->nsites * per;
      if (nd == n) {
        if (int rc = sipnet_gpu_gather(h, what, out, rowsD * n * 8)) return rc;
      } else {
        for (size_t i = 0; i < rowsD * n; ++i) out[i] = __builtin_nan("");
        if (nd > 0) {
          tmp.resize(rowsD * nd);
          if (int rc = sipnet_gpu_gather(h, what, tmp.data(), rowsD * nd * 8)) return rc;
          for (size_t r = 0; r < rowsD; ++r) memcpy(out + r * n, tmp.data() + r * nd, nd * 8);
        }
      }
      out += rowsD * n;
    }
    return 0;
  }
  size_t cols = 0, elem = 8;
  bool perStep = false;
  member_output_shape(m, what, &cols, &perStep, &elem);
  const size_t M = (size_t)m->nmembers;
  const size_t steps = perStep ? n : 1;
  std::vector<char> tmp;
  for (size_t d = 0; d < D; ++d) {
    sipnet_gpu_handle *h = m->handles[d];
    const size_t nm = (size_t)h->nmembers;
    const size_t nd = perStep ? (size_t)(h->lastEnd - h->lastBegin) : 1;
    const sipnet_gpu_multi::Part &p = m->parts[d];
    if (nd > 0) {
      tmp.resize(cols * nd * nm * elem);
      if (int rc = sipnet_gpu_gather(h, what, tmp.data(), tmp.size())) return rc;
    }
    auto scatter = [&](size_t local0, size_t global0, size_t count) {
      for (size_t c = 0; c < cols; ++c)
        for (size_t t = 0; t < steps; ++t) {
          char *to = static_cast<char *>(dst) + ((c * steps + t) * M + global0) * elem;
          if (t < nd) {
            memcpy(to, tmp.data() + ((c * nd + t) * nm + local0) * elem, count * elem);
          } else {  // steps this device's (shorter) sites never ran: NaN, as on one GPU
            memset(to, 0xFF, count * elem);
          }
        }
    };
    if (!m->split) {
      scatter(0, (size_t)p.member0, nm);
    } else {
      for (size_t s = 0; s < p.siteCount.size(); ++s)
        scatter((size_t)p.siteLocal0[s], (size_t)p.siteGlobal0[s], (size_t)p.siteCount[s]);
    }
  }
  return 0;
}
