// Multi-GPU plumbing of libltxcuda: one context per process / GPU, NCCL communicators owned by the context.
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch already loaded when there is one), so the library
// builds and loads on machines without NCCL; multi-GPU entry points fail loudly there.
//
// Rank layout: world = pass_groups x sp_size, rank = group * sp_size + sp_rank.
//   * pass groups   -- the conditional / unconditional / STG forwards of one denoise step are independent given (x, sigma)
//                      (P/LTXPipeline.swift:829-848, 897-914): pass p runs on group p % pass_groups, velocities are
//                      broadcast, the guided Euler update is replicated.
//   * sp (Ulysses)  -- tokens are sharded over the sp ranks of a group for every row-wise op; q/k/v are exchanged to a
//                      head-sharded layout around self-attention (all-to-all), the output back (dit.cu).
//   * VAE           -- latent frames are sharded in contiguous temporal slabs over all ranks; every conv exchanges one
//                      boundary frame with each temporal neighbour (vae.cu).
#include <dlfcn.h>
#include <nccl.h>

#include "ctx.h"

namespace ltx {

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& nccl() {
  static NcclApi api;
  if (api.handle) return api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  LTX_CHECK(api.handle != nullptr, LTX_ERR_UNSUPPORTED, "libnccl.so.2 not found: multi-GPU entry points are unavailable");
  auto sym = [&](const char* s) {
    void* p = dlsym(api.handle, s);
    LTX_CHECK(p != nullptr, LTX_ERR_UNSUPPORTED, std::string("NCCL symbol missing: ") + s);
    return p;
  };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommSplit = reinterpret_cast<decltype(api.CommSplit)>(sym("ncclCommSplit"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
  api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
  api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  return api;
}

#define LTX_NCCL(expr)                                                                                          \
  do {                                                                                                          \
    ncclResult_t _r = (expr);                                                                                   \
    if (_r != ncclSuccess)                                                                                      \
      throw ::ltx::LtxError(LTX_ERR_CUDA, std::string("NCCL error: ") + nccl().GetErrorString(_r) + " in " #expr); \
  } while (0)

ncclComm_t world(ltx_ctx* c) { return reinterpret_cast<ncclComm_t>(c->dist.comm_world); }
ncclComm_t spc(ltx_ctx* c) { return reinterpret_cast<ncclComm_t>(c->dist.comm_sp); }

}  // namespace

void dist_get_unique_id(void* out128) {
  ncclUniqueId id;
  LTX_NCCL(nccl().GetUniqueId(&id));
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, sizeof(id));
}

void dist_init(ltx_ctx* c, const void* unique_id, int rank, int world_size, int sp_size, int pass_groups) {
  LTX_CHECK(world_size >= 1 && rank >= 0 && rank < world_size, LTX_ERR_INVALID_ARGUMENT, "bad rank / world size");
  LTX_CHECK(sp_size >= 1 && pass_groups >= 1 && sp_size * pass_groups == world_size, LTX_ERR_INVALID_CONFIGURATION,
            "sp_size * pass_groups must equal world_size");
  LTX_CHECK(c->cfg.num_heads % sp_size == 0, LTX_ERR_INVALID_CONFIGURATION, "sp_size must divide num_heads");
  LTX_CHECK(c->dist.comm_world == nullptr, LTX_ERR_INVALID_ARGUMENT, "distributed state already initialised");
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclComm_t w = nullptr, s = nullptr;
  LTX_NCCL(nccl().CommInitRank(&w, world_size, id, rank));
  c->dist.comm_world = w;
  c->dist.rank = rank; c->dist.world = world_size; c->dist.sp = sp_size; c->dist.groups = pass_groups;
  c->dist.group = rank / sp_size; c->dist.sp_rank = rank % sp_size;
  if (sp_size > 1 && pass_groups > 1) {
    LTX_NCCL(nccl().CommSplit(w, c->dist.group, c->dist.sp_rank, &s, nullptr));
    c->dist.comm_sp = s;
    c->dist.sp_is_world = false;
  } else {
    c->dist.comm_sp = w;   // sp group == world (or sp == 1: unused)
    c->dist.sp_is_world = true;
  }
}

void dist_destroy(ltx_ctx* c) {
  if (!c->dist.comm_world) return;
  if (c->dist.comm_sp && !c->dist.sp_is_world) nccl().CommDestroy(spc(c));
  nccl().CommDestroy(world(c));
  c->dist = DistState();
}

void dist_broadcast(ltx_ctx* c, void* buf, size_t bytes, int root_world_rank) {
  LTX_NCCL(nccl().Broadcast(buf, buf, bytes, ncclChar, root_world_rank, world(c), c->stream));
}

void dist_allgather_sp(ltx_ctx* c, const void* send, void* recv, size_t bytes_per_rank) {
  LTX_NCCL(nccl().AllGather(send, recv, bytes_per_rank, ncclChar, spc(c), c->stream));
}

// all-to-all inside the sp group: block s of `send` (bytes_per_peer each) goes to sp rank s, block s of `recv` comes from it
void dist_all_to_all_sp(ltx_ctx* c, const void* const* send, void* const* recv, int n_tensors, size_t bytes_per_peer) {
  const int P = c->dist.sp;
  LTX_NCCL(nccl().GroupStart());
  for (int t = 0; t < n_tensors; ++t) {
    const char* sb = reinterpret_cast<const char*>(send[t]);
    char* rb = reinterpret_cast<char*>(recv[t]);
    for (int s = 0; s < P; ++s) {
      LTX_NCCL(nccl().Send(sb + static_cast<size_t>(s) * bytes_per_peer, bytes_per_peer, ncclChar, s, spc(c), c->stream));
      LTX_NCCL(nccl().Recv(rb + static_cast<size_t>(s) * bytes_per_peer, bytes_per_peer, ncclChar, s, spc(c), c->stream));
    }
  }
  LTX_NCCL(nccl().GroupEnd());
}

// exchange with the temporal neighbours (world ranks rank-1 / rank+1 among the first n_active ranks)
void dist_halo_exchange(ltx_ctx* c, const void* send_prev, void* recv_prev, const void* send_next, void* recv_next,
                        size_t bytes, int n_active) {
  const int r = c->dist.rank;
  LTX_NCCL(nccl().GroupStart());
  if (r > 0 && r < n_active) {
    LTX_NCCL(nccl().Send(send_prev, bytes, ncclChar, r - 1, world(c), c->stream));
    LTX_NCCL(nccl().Recv(recv_prev, bytes, ncclChar, r - 1, world(c), c->stream));
  }
  if (r + 1 < n_active) {
    LTX_NCCL(nccl().Send(send_next, bytes, ncclChar, r + 1, world(c), c->stream));
    LTX_NCCL(nccl().Recv(recv_next, bytes, ncclChar, r + 1, world(c), c->stream));
  }
  LTX_NCCL(nccl().GroupEnd());
}

}  // namespace ltx
