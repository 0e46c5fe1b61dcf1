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
#include <unistd.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>

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

NcclApi& nccl_load(NcclApi& api) {
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
// loaded once per process, on first use, from any thread (function-local static initialisation is thread-safe; a failed
// load throws and is retried by the next caller)
NcclApi& nccl() {
  static NcclApi api;
  static NcclApi& ready = nccl_load(api);
  return ready;
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

// All ranks in ONE process (the drop-in surface is a single `actor LTXPipeline`, Pipeline/LTXPipeline.swift:117, so a Swift host
// has no launcher): contexts[i] becomes rank i.  One thread creates every communicator inside an NCCL group; afterwards the
// collective entry points (ltx_denoise_step, ltx_dit_forward, ltx_vae_decode, ...) must be called concurrently, one host thread
// per context -- the same contract torchrun's one-process-per-GPU gives, minus the processes.
void dist_init_local(ltx_ctx** cs, int n, int sp_size, int pass_groups) {
  LTX_CHECK(cs != nullptr && n >= 1 && n <= 64, LTX_ERR_INVALID_ARGUMENT, "bad context list");
  LTX_CHECK(sp_size >= 1 && pass_groups >= 1 && sp_size * pass_groups == n, LTX_ERR_INVALID_CONFIGURATION,
            "sp_size * pass_groups must equal the number of contexts");
  for (int i = 0; i < n; ++i) {
    LTX_CHECK(cs[i] != nullptr && cs[i]->dist.comm_world == nullptr, LTX_ERR_INVALID_ARGUMENT, "context missing or already distributed");
    LTX_CHECK(cs[i]->cfg.num_heads % sp_size == 0, LTX_ERR_INVALID_CONFIGURATION, "sp_size must divide num_heads");
    for (int j = 0; j < i; ++j) LTX_CHECK(cs[j]->device != cs[i]->device, LTX_ERR_INVALID_ARGUMENT, "one context per device");
  }
  ncclUniqueId id;
  LTX_NCCL(nccl().GetUniqueId(&id));
  std::vector<ncclComm_t> w(n, nullptr), s(n, nullptr);
  LTX_NCCL(nccl().GroupStart());
  for (int i = 0; i < n; ++i) {
    LTX_CUDA(cudaSetDevice(cs[i]->device));
    LTX_NCCL(nccl().CommInitRank(&w[i], n, id, i));
  }
  LTX_NCCL(nccl().GroupEnd());
  const bool split = sp_size > 1 && pass_groups > 1;
  if (split) {
    LTX_NCCL(nccl().GroupStart());
    for (int i = 0; i < n; ++i) {
      LTX_CUDA(cudaSetDevice(cs[i]->device));
      LTX_NCCL(nccl().CommSplit(w[i], i / sp_size, i % sp_size, &s[i], nullptr));
    }
    LTX_NCCL(nccl().GroupEnd());
  }
  for (int i = 0; i < n; ++i) {
    DistState& d = cs[i]->dist;
    d.comm_world = w[i];
    d.rank = i; d.world = n; d.sp = sp_size; d.groups = pass_groups;
    d.group = i / sp_size; d.sp_rank = i % sp_size;
    d.comm_sp = split ? s[i] : w[i];
    d.sp_is_world = !split;
  }
}

// ------------------------------------------------------------------------------------------------ peer-memory Ulysses
namespace {

constexpr size_t P2P_FLAG_BYTES = 256;   // [2 kinds][8 source ranks] uint32 flags + [2] local epoch counters, at the end of the exported allocation

// The epoch of barrier `kind` lives in this rank's own allocation (behind the flags) and is advanced by the kernel itself:
// every rank runs the same sequence of barriers, so the counters agree, and a CUDA-graph replay of a captured step gets
// fresh epochs without any host-side argument.
__global__ void p2p_barrier_kernel(PeerTable peers, size_t flag_off, int P, int me, int kind) {
  const int t = threadIdx.x;
  __shared__ uint32_t epoch_s;
  if (t == 0) {
    uint32_t* ctr = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(peers.p[me]) + flag_off) + 2 * LTX_MAX_PEERS + kind;
    epoch_s = *ctr + 1;
    *ctr = epoch_s;
  }
  __syncthreads();
  const uint32_t epoch = epoch_s;
  if (t >= P) return;
  // every store of the preceding kernels of this stream has completed (kernel boundary); publish that to rank t ...
  __threadfence_system();
  uint32_t* remote = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(peers.p[t]) + flag_off) + kind * LTX_MAX_PEERS + me;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  // ... and wait until rank t has published the same for its stores into our buffer
  const uint32_t* local = reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(peers.p[me]) + flag_off) + kind * LTX_MAX_PEERS + t;
  uint32_t v = 0;
  unsigned long long spins = 0;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(local) : "memory");
    if (++spins > (1ull << 25)) {   // ~ tens of seconds: a peer died or the protocol is broken -> abort instead of hanging
      printf("ltxcuda: peer barrier timeout (rank %d waiting for rank %d, kind %d, epoch %u, seen %u)\n", me, t, kind, epoch, v);
      __trap();
    }
  } while (static_cast<int32_t>(v - epoch) < 0);
}

void p2p_release(ltx_ctx* c) {
  DistState& d = c->dist;
  for (int r = 0; r < d.sp; ++r)
    if (d.p2p_peer[r] && r != d.sp_rank && d.p2p_ipc[r]) cudaIpcCloseMemHandle(d.p2p_peer[r]);
  for (auto& q : d.p2p_peer) q = nullptr;
  if (d.p2p_pool) {   // same-process mode: the buffer came from the context's peer-visible pool
    if (d.p2p_local) { cudaFreeAsync(d.p2p_local, c->stream); cudaStreamSynchronize(c->stream); }
    cudaMemPoolDestroy(reinterpret_cast<cudaMemPool_t>(d.p2p_pool));
    d.p2p_pool = nullptr;
  } else if (d.p2p_local) {
    cudaFree(d.p2p_local);
  }
  d.p2p_local = nullptr;
  d.p2p_bytes = 0;
  d.p2p = false;
}

}  // namespace

bool dist_p2p_ensure(ltx_ctx* c, size_t bytes) {
  DistState& d = c->dist;
  if (d.sp <= 1 || d.sp > LTX_MAX_PEERS) return false;
  static const bool enabled = [] { const char* e = getenv("LTX_P2P"); return e ? atoi(e) != 0 : true; }();
  if (!enabled) return false;
  if (d.p2p && d.p2p_bytes >= bytes) return true;
  if (d.p2p_tried && !d.p2p) return false;   // mapping failed before (all ranks agree, see below): stay on NCCL
  d.p2p_tried = true;
  // (re)registration is collective: every rank of the sp group runs the same forward and gets here with the same size
  LTX_CUDA(cudaStreamSynchronize(c->stream));
  const int P = d.sp, me = d.sp_rank;
  // Ranks of the same process (ltx_dist_init_local) cannot open each other's IPC handles -- and do not need to: they share an
  // address space.  Their buffers come from a per-context memory pool whose access list names the peer devices
  // (cudaMemPoolSetAccess), so a peer's pointer is usable as it is.  NOT cudaDeviceEnablePeerAccess: with blanket peer access
  // every later cudaMalloc on one device has to be mapped into the other and waits for it to go idle -- which it never does
  // while its barrier kernel spins for a flag this rank has yet to write (seen as a peer-barrier timeout).
  struct Msg { cudaIpcMemHandle_t h; int ok; int pid; int device; int pad0; void* ptr; int pad[10]; };
  static_assert(sizeof(Msg) == 128, "handle message is 128 bytes");
  Msg* dev = nullptr;
  LTX_CUDA(cudaMalloc(&dev, sizeof(Msg) * (P + 1)));
  auto exchange = [&](Msg mine, std::vector<Msg>& all) {
    LTX_CUDA(cudaMemcpyAsync(dev + P, &mine, sizeof(Msg), cudaMemcpyHostToDevice, c->stream));
    LTX_NCCL(nccl().AllGather(dev + P, dev, sizeof(Msg), ncclChar, spc(c), c->stream));
    all.resize(P);
    LTX_CUDA(cudaMemcpyAsync(all.data(), dev, sizeof(Msg) * P, cudaMemcpyDeviceToHost, c->stream));
    LTX_CUDA(cudaStreamSynchronize(c->stream));
  };
  std::vector<Msg> all;
  if (d.p2p_local) {   // growing: nobody may still be storing into the old buffers
    Msg m = {};
    exchange(m, all);
    p2p_release(c);
  }
  // round 0: who lives where
  Msg mine = {};
  mine.pid = static_cast<int>(getpid());
  mine.device = c->device;
  exchange(mine, all);
  bool inproc = true;
  for (int r = 0; r < P; ++r) inproc = inproc && all[r].pid == mine.pid;
  const size_t cap = (bytes + 1023) / 1024 * 1024;
  void* local = nullptr;
  if (inproc) {
    cudaMemPool_t pool = nullptr;
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = c->device;
    bool ok = cudaMemPoolCreate(&pool, &props) == cudaSuccess;
    std::vector<cudaMemAccessDesc> acc;
    for (int r = 0; r < P && ok; ++r) {
      if (r == me) continue;
      int can = 0;
      ok = cudaDeviceCanAccessPeer(&can, all[r].device, c->device) == cudaSuccess && can;
      cudaMemAccessDesc a = {};
      a.location.type = cudaMemLocationTypeDevice;
      a.location.id = all[r].device;
      a.flags = cudaMemAccessFlagsProtReadWrite;
      acc.push_back(a);
    }
    ok = ok && (acc.empty() || cudaMemPoolSetAccess(pool, acc.data(), acc.size()) == cudaSuccess);
    ok = ok && cudaMallocFromPoolAsync(&local, cap + P2P_FLAG_BYTES, pool, c->stream) == cudaSuccess &&
         cudaMemsetAsync(local, 0, cap + P2P_FLAG_BYTES, c->stream) == cudaSuccess && cudaStreamSynchronize(c->stream) == cudaSuccess;
    if (!ok) {
      cudaGetLastError();
      if (local) { cudaFreeAsync(local, c->stream); cudaStreamSynchronize(c->stream); local = nullptr; }
      if (pool) cudaMemPoolDestroy(pool);
      pool = nullptr;
    }
    d.p2p_pool = pool;
    mine.ok = ok ? 1 : 0;
  } else {
    mine.ok = cudaMalloc(&local, cap + P2P_FLAG_BYTES) == cudaSuccess && cudaMemset(local, 0, cap + P2P_FLAG_BYTES) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.h, local) == cudaSuccess;
    cudaGetLastError();
  }
  mine.ptr = local;
  exchange(mine, all);
  bool ok = true;
  for (int r = 0; r < P; ++r) ok = ok && all[r].ok;
  void* peer[LTX_MAX_PEERS] = {};
  bool via_ipc[LTX_MAX_PEERS] = {};
  if (ok) {
    for (int r = 0; r < P && ok; ++r) {
      if (r == me) { peer[r] = local; continue; }
      if (inproc) {
        peer[r] = all[r].ptr;   // same address space, access granted by the owner's pool
      } else {
        ok = cudaIpcOpenMemHandle(&peer[r], all[r].h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
        via_ipc[r] = ok;
      }
      if (!ok) cudaGetLastError();
    }
  }
  // second round: everyone must have mapped everyone, otherwise all ranks fall back to NCCL together
  Msg res = {};
  res.ok = ok ? 1 : 0;
  exchange(res, all);
  bool all_ok = true;
  for (int r = 0; r < P; ++r) all_ok = all_ok && all[r].ok;
  cudaFree(dev);
  if (!all_ok) {
    for (int r = 0; r < P; ++r)
      if (peer[r] && r != me && via_ipc[r]) cudaIpcCloseMemHandle(peer[r]);
    d.p2p_local = local;
    p2p_release(c);   // frees `local` the way it was allocated
    return false;
  }
  d.p2p_local = local;
  d.p2p_bytes = cap;
  for (int r = 0; r < P; ++r) { d.p2p_peer[r] = peer[r]; d.p2p_ipc[r] = via_ipc[r]; }
  d.p2p = true;
  return true;
}

void dist_p2p_barrier(ltx_ctx* c, int kind) {
  DistState& d = c->dist;
  LTX_CHECK(d.p2p && (kind == 0 || kind == 1), LTX_ERR_INVALID_ARGUMENT, "peer barrier without peer memory");
  PeerTable t = {};
  for (int r = 0; r < d.sp; ++r) t.p[r] = d.p2p_peer[r];
  p2p_barrier_kernel<<<1, 32, 0, c->stream>>>(t, d.p2p_bytes, d.sp, d.sp_rank, kind);
  LTX_CUDA(cudaGetLastError());
}

void dist_destroy(ltx_ctx* c) {
  if (!c->dist.comm_world) return;
  if (c->dist.p2p_local) {
    // peers may still be storing into our buffer: drain the stream, meet at a collective, then unmap and free
    cudaStreamSynchronize(c->stream);
    char* one = nullptr;
    if (cudaMalloc(&one, 2 * LTX_MAX_PEERS) == cudaSuccess) {
      nccl().AllGather(one + LTX_MAX_PEERS, one, 1, ncclChar, spc(c), c->stream);
      cudaStreamSynchronize(c->stream);
      cudaFree(one);
    }
    p2p_release(c);
  }
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
