// pcaone_b200 — collectives between the ranks of a sharded job.
//
// One context per GPU; the exchange steps of the path are sums of small or tall partial products
// (N x l partial H, range x l int64 partial G accumulators, l x l Gram matrices, per-SNP genotype
// counts). They run on the context's stream through an NCCL communicator owned by the library:
//   pcaone_comm_unique_id + pcaone_comm_init   one process per GPU (torchrun): rank 0 makes the id,
//                                             the host broadcasts the 128 bytes, every rank joins
//   pcaone_comm_attach                         one process, several GPUs: the host made the
//                                             communicators itself (ncclCommInitAll)
// libnccl is resolved at run time (dlopen), so the library loads on a box without it and a
// Python host that already imported torch shares torch's copy. Without a communicator the host
// hooks are used: pcaone_set_allreduce2 (typed; any transport, e.g. gloo between two ranks that
// time-share one GPU in the tests) or the older pcaone_set_allreduce (double sums only).
#include <dlfcn.h>
#include <nccl.h>

#include "ctx.hpp"

struct pcaone_comm {
  ncclComm_t comm = nullptr;
  bool owned = false;
};

namespace pcaone {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
};

NcclApi& api() {
  static NcclApi a = [] {
    NcclApi n;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      n.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (n.handle) break;
    }
    if (!n.handle) {
      n.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
      return n;
    }
    auto sym = [&](const char* s) {
      void* p = dlsym(n.handle, s);
      if (!p) n.why = std::string("libnccl is missing ") + s;
      return p;
    };
    n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
    n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
    n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
    n.AllReduce = (decltype(n.AllReduce))sym("ncclAllReduce");
    n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
    n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
    n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
    return n;
  }();
  return a;
}

void need_api() {
  if (!api().why.empty() || !api().handle) throw std::runtime_error("NCCL unavailable: " + api().why);
}

void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess)
    throw std::runtime_error(std::string(what) + " failed: " + (api().GetErrorString ? api().GetErrorString(r) : "?"));
}

void reduce(pcaone_ctx* c, void* buf, uint64_t count, ncclDataType_t dt, ncclRedOp_t op) {
  if (c->cfg.world <= 1 || count == 0) return;
  Timed t(c, 5);
  if (c->comm && c->comm->comm) {
    nccl_check(api().AllReduce(buf, buf, (size_t)count, dt, op, c->comm->comm, c->stream), "ncclAllReduce");
    return;
  }
  if (c->allreduce2) {
    const int kind = dt == ncclFloat64 ? PCAONE_RED_F64_SUM : dt == ncclInt64 ? PCAONE_RED_I64_SUM
                     : dt == ncclUint64 ? PCAONE_RED_U64_MAX : PCAONE_RED_U32_SUM;
    if (c->allreduce2(c->allreduce2_user, buf, count, kind, c->stream)) throw std::runtime_error("allreduce hook failed");
    return;
  }
  if (dt == ncclFloat64 && op == ncclSum && c->allreduce) {
    if (c->allreduce(c->allreduce_user, buf, count, c->stream)) throw std::runtime_error("allreduce hook failed");
    return;
  }
  throw std::runtime_error("world > 1 but no communicator attached (pcaone_comm_init / pcaone_comm_attach)");
}

}  // namespace

void comm_allreduce_f64(pcaone_ctx* c, double* buf, uint64_t count) { reduce(c, buf, count, ncclFloat64, ncclSum); }
void comm_allreduce_i64(pcaone_ctx* c, long long* buf, uint64_t count) { reduce(c, buf, count, ncclInt64, ncclSum); }
void comm_allreduce_u64_max(pcaone_ctx* c, unsigned long long* buf, uint64_t count) {
  reduce(c, buf, count, ncclUint64, ncclMax);
}
void comm_allreduce_u32(pcaone_ctx* c, uint32_t* buf, uint64_t count) { reduce(c, buf, count, ncclUint32, ncclSum); }

// several reductions as ONE launch (only meaningful with the in-library communicator)
void comm_group_begin(pcaone_ctx* c) {
  if (c->cfg.world > 1 && c->comm && c->comm->comm) nccl_check(api().GroupStart(), "ncclGroupStart");
}
void comm_group_end(pcaone_ctx* c) {
  if (c->cfg.world > 1 && c->comm && c->comm->comm) nccl_check(api().GroupEnd(), "ncclGroupEnd");
}

void comm_destroy(pcaone_ctx* c) {
  for (void* p : c->peer_opened) cudaIpcCloseMemHandle(p);
  c->peer_opened.clear();
  if (c->d_mbox) cudaFree(c->d_mbox);
  if (c->d_peer_mbox) cudaFree(c->d_peer_mbox);
  if (c->d_peer_flag) cudaFree(c->d_peer_flag);
  c->d_mbox = nullptr;
  c->d_peer_mbox = nullptr;
  c->d_peer_flag = nullptr;
  c->peer_ready = false;
  if (!c->comm) return;
  if (c->comm->owned && c->comm->comm && api().CommDestroy) api().CommDestroy(c->comm->comm);
  delete c->comm;
  c->comm = nullptr;
}

}  // namespace pcaone

using namespace pcaone;

extern "C" {

int pcaone_comm_unique_id(uint8_t* out128) {
  try {
    need_api();
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    nccl_check(api().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(out128, &id, sizeof(id));
    return 0;
  } catch (const std::exception&) {
    return 1;
  }
}

int pcaone_comm_init(pcaone_ctx* c, const uint8_t* id128, int rank, int world) {
  CTX_GUARD(c, {
    need_api();
    if (rank != c->cfg.rank || world != c->cfg.world) throw std::runtime_error("comm_init: rank / world differ from the context's");
    comm_destroy(c);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    c->comm = new pcaone_comm();
    nccl_check(api().CommInitRank(&c->comm->comm, world, id, rank), "ncclCommInitRank");
    c->comm->owned = true;
  });
}

int pcaone_set_allreduce2(pcaone_ctx* c, pcaone_allreduce2_fn fn, void* user) {
  CTX_GUARD(c, {
    c->allreduce2 = fn;
    c->allreduce2_user = user;
  });
}

// Peer-memory mailboxes (one process per GPU): every rank exports its mailbox as a CUDA IPC handle, the host
// gathers the world's handles, every rank imports them. From then on the row-sharded Omega update is ONE
// cooperative launch whose three small exchanges run inside the kernel over NVLink (orth_fused.cuh).
static constexpr int kSlots = 4;   // == kPeerSlots of orth_fused.cuh
int pcaone_comm_peer_export(pcaone_ctx* c, uint8_t* out64) {
  CTX_GUARD(c, {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (c->cfg.world < 2 || c->cfg.world > 16) throw std::runtime_error("peer_export: world must be 2..16");
    if (!c->d_mbox) {
      const size_t data = (size_t)kSlots * c->cfg.world * c->l * c->lp * sizeof(double);
      c->mbox_flag_off = round_up(data, 256);
      c->mbox_bytes = c->mbox_flag_off + (size_t)kSlots * c->cfg.world * sizeof(unsigned long long);
      PCA_CUDA(cudaMalloc((void**)&c->d_mbox, c->mbox_bytes));
      PCA_CUDA(cudaMemset(c->d_mbox, 0, c->mbox_bytes));
    }
    cudaIpcMemHandle_t hnd;
    PCA_CUDA(cudaIpcGetMemHandle(&hnd, c->d_mbox));
    memcpy(out64, &hnd, sizeof(hnd));
  });
}

int pcaone_comm_peer_import(pcaone_ctx* c, const uint8_t* handles, int nranks) {
  CTX_GUARD(c, {
    if (nranks != c->cfg.world || !c->d_mbox) throw std::runtime_error("peer_import: call pcaone_comm_peer_export first, pass world handles");
    if (c->d_peer_mbox) throw std::runtime_error("peer_import: mailboxes are already mapped");
    std::vector<double*> mb(nranks);
    std::vector<unsigned long long*> fl(nranks);
    for (int r = 0; r < nranks; ++r) {
      void* base = nullptr;
      if (r == c->cfg.rank) {
        base = c->d_mbox;
      } else {
        cudaIpcMemHandle_t hnd;
        memcpy(&hnd, handles + (size_t)r * 64, sizeof(hnd));
        PCA_CUDA(cudaIpcOpenMemHandle(&base, hnd, cudaIpcMemLazyEnablePeerAccess));
        c->peer_opened.push_back(base);
      }
      mb[r] = reinterpret_cast<double*>(base);
      fl[r] = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(base) + c->mbox_flag_off);
    }
    PCA_CUDA(cudaMalloc((void**)&c->d_peer_mbox, nranks * sizeof(double*)));
    PCA_CUDA(cudaMalloc((void**)&c->d_peer_flag, nranks * sizeof(unsigned long long*)));
    PCA_CUDA(cudaMemcpy(c->d_peer_mbox, mb.data(), nranks * sizeof(double*), cudaMemcpyHostToDevice));
    PCA_CUDA(cudaMemcpy(c->d_peer_flag, fl.data(), nranks * sizeof(unsigned long long*), cudaMemcpyHostToDevice));
    c->peer_seq = 1;
    c->peer_ready = getenv("PCAONE_PEER_EXCHANGE") ? atoi(getenv("PCAONE_PEER_EXCHANGE")) != 0 : true;
  });
}

int pcaone_comm_attach(pcaone_ctx* c, void* nccl_comm) {
  CTX_GUARD(c, {
    need_api();
    comm_destroy(c);
    c->comm = new pcaone_comm();
    c->comm->comm = (ncclComm_t)nccl_comm;
    c->comm->owned = false;
  });
}

}  // extern "C"
