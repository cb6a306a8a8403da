// One process per GPU: NCCL halo exchange (grouped send/recv to the <= few neighbouring row
// blocks) and scalar all-reduce for the PCG dot products, over NVLink 5 / NVSwitch.
// The reference has no distributed code at all; the partition is SURVEY §8e.
//
// NCCL is resolved with dlopen("libnccl.so.2") at fe_dist_init time: in a torch process
// that returns the copy torch already loaded, and libfe_b200.so itself stays loadable on
// a machine without NCCL (the CPU-side symbol test).
#include <dlfcn.h>

#include "dist.h"

namespace fe {

typedef struct ncclComm *ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };  // nccl.h: ncclDouble = 8
enum { ncclSum = 0 };

struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.lib) return FE_OK;
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return fail(FE_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define FE_SYM(field, name)                                                        \
  do {                                                                             \
    *(void **)(&g_nccl.field) = dlsym(lib, name);                                  \
    if (!g_nccl.field) return fail(FE_ERR_NCCL, "libnccl: missing symbol %s", name); \
  } while (0)
  FE_SYM(GetUniqueId, "ncclGetUniqueId");
  FE_SYM(CommInitRank, "ncclCommInitRank");
  FE_SYM(CommDestroy, "ncclCommDestroy");
  FE_SYM(AllReduce, "ncclAllReduce");
  FE_SYM(Send, "ncclSend");
  FE_SYM(Recv, "ncclRecv");
  FE_SYM(GroupStart, "ncclGroupStart");
  FE_SYM(GroupEnd, "ncclGroupEnd");
  FE_SYM(GetErrorString, "ncclGetErrorString");
#undef FE_SYM
  g_nccl.lib = lib;
  return FE_OK;
}

#define FE_NCCL(call)                                                                              \
  do {                                                                                             \
    int r__ = (call);                                                                              \
    if (r__ != ncclSuccess)                                                                        \
      return fail(FE_ERR_NCCL, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                \
                  g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?");                       \
  } while (0)

__global__ void k_halo_pack(int32_t n, const int32_t *__restrict__ idx, const double *__restrict__ vec,
                            double *__restrict__ buf) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = vec[idx[i]];
}

int halo_exchange(fe_ctx *ctx, cudaStream_t s, const HaloPlan *h, double *vec, int32_t n_rows) {
  if (!h || h->n_nbr == 0) return FE_OK;
  FE_REQUIRE(ctx->nccl_comm, "halo_exchange: fe_dist_init was not called");
  const int32_t n_send = h->send_ptr[h->n_nbr];
  int rc = ctx->halo_send.reserve((size_t)(n_send > 0 ? n_send : 1) * sizeof(double));
  if (rc) return rc;
  double *buf = (double *)ctx->halo_send.ptr;
  if (n_send > 0) {
    k_halo_pack<<<grid_for(n_send, 256), 256, 0, s>>>(n_send, h->send_idx, vec, buf);
    FE_LAUNCH_CHECK(ctx);
  }
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  FE_NCCL(g_nccl.GroupStart());
  for (int k = 0; k < h->n_nbr; ++k) {
    const int32_t ns = h->send_ptr[k + 1] - h->send_ptr[k];
    const int32_t nr = h->recv_ptr[k + 1] - h->recv_ptr[k];
    if (ns > 0) FE_NCCL(g_nccl.Send(buf + h->send_ptr[k], (size_t)ns, ncclFloat64, h->nbr_rank[k], comm, s));
    if (nr > 0)
      FE_NCCL(g_nccl.Recv(vec + n_rows + h->recv_ptr[k], (size_t)nr, ncclFloat64, h->nbr_rank[k], comm, s));
  }
  FE_NCCL(g_nccl.GroupEnd());
  return FE_OK;
}

// ---------------------------------------------------------------------------------------
// peer-memory transport (no NCCL inside the PCG loop)
// ---------------------------------------------------------------------------------------
// One launch per exchange.  Every thread first stores its share of the interface values straight
// into the neighbours' ghost cells (peer stores over NVLink, LL cells: no fence, no flag), then
// waits for its share of this rank's own ghost cells and moves them behind `vec`.  The wait is per
// cell, so no grid-wide synchronisation is needed; the last CTA to finish bumps halo_seq.
__global__ void __launch_bounds__(256) k_halo_ll(P2PDev *pp, HaloDev *hd, const int32_t *__restrict__ send_idx,
                                                const double *__restrict__ vec, double *__restrict__ vec_tail) {
  const unsigned seq = pp->halo_seq + 1, par = seq & 1;  // cell of ghost j and parity: 2 j + par
  const int n_ghost = hd->n_ghost;
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  halo_push(pp, hd, send_idx, vec, seq, t0, stride);
  const uint4 *mine = pp->ghost[pp->rank] + par;
  for (int i = t0; i < n_ghost; i += stride) vec_tail[i] = ll_wait(mine + 2 * (size_t)i, seq);
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(&hd->ticket, 1u) == gridDim.x - 1) {
    hd->ticket = 0;
    pp->halo_seq = seq;
  }
}

int halo_exchange_p2p(fe_ctx *ctx, cudaStream_t s, const HaloPlan *h, double *vec, int32_t n_rows) {
  if (!h || h->n_nbr == 0) return FE_OK;
  FE_REQUIRE(ctx->p2p_dev && ctx->p2p_halo.ptr, "halo_exchange_p2p: peer memory is not set up");
  HaloDev *hd = (HaloDev *)ctx->p2p_halo.ptr;
  const int32_t n_send = h->send_ptr[h->n_nbr], n_ghost = h->recv_ptr[h->n_nbr];
  const int32_t n_max = n_send > n_ghost ? n_send : n_ghost;
  if (n_max > 0) {
    // every CTA spins on cells a peer fills: the grid must be co-resident (<= 1 CTA per SM is plenty
    // for an interface; a 2-D stripe at S16M is 16 K values)
    int grid = grid_for(n_max, 256);
    if (grid > ctx->num_sms) grid = ctx->num_sms;
    k_halo_ll<<<grid, 256, 0, s>>>(ctx->p2p_dev, hd, h->send_idx, vec, vec + n_rows);
    FE_LAUNCH_CHECK(ctx);
  }
  return FE_OK;
}

// every neighbour's slice of send_idx ascending?  (the persistent PCG kernel finds a CTA's share of a
// send list by binary search)
__global__ void k_check_sorted(int32_t n, const int32_t *__restrict__ idx, int32_t n_seg, const int32_t *__restrict__ seg_ptr,
                               int *__restrict__ bad) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i + 1 >= n) return;
  for (int k = 1; k < n_seg; ++k)
    if (seg_ptr[k] == i + 1) return;  // segment boundary
  if (idx[i] >= idx[i + 1]) *bad = 1;
}

static int upload_p2p_halo(fe_ctx *ctx, cudaStream_t s, const HaloPlan *h) {
  FE_REQUIRE(h->n_nbr <= kMaxRanks, "p2p: more than %d neighbours", kMaxRanks);
  FE_REQUIRE(h->recv_ptr[h->n_nbr] <= ctx->p2p_n_ghost, "p2p: ghost block too small (%d > %d): call fe_dist_p2p_export again",
             h->recv_ptr[h->n_nbr], ctx->p2p_n_ghost);
  HaloDev hd;
  memset(&hd, 0, sizeof(hd));
  hd.n_nbr = h->n_nbr;
  hd.n_send = h->send_ptr[h->n_nbr];
  hd.n_ghost = h->recv_ptr[h->n_nbr];
  for (int k = 0; k < h->n_nbr; ++k) {
    hd.nbr_rank[k] = h->nbr_rank[k];
    hd.send_ptr[k] = h->send_ptr[k];
    hd.dst_off[k] = h->peer_dst_off[k];
  }
  hd.send_ptr[h->n_nbr] = h->send_ptr[h->n_nbr];
  int rc = ctx->p2p_halo.reserve(sizeof(HaloDev) + 16);
  if (rc) return rc;
  FE_CUDA(cudaMemcpyAsync(ctx->p2p_halo.ptr, &hd, sizeof(hd), cudaMemcpyHostToDevice, s));
  int bad = 0;
  if (hd.n_send > 1) {
    int *flag = (int *)((char *)ctx->p2p_halo.ptr + sizeof(HaloDev));
    FE_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));
    k_check_sorted<<<grid_for(hd.n_send, 256), 256, 0, s>>>(hd.n_send, h->send_idx, hd.n_nbr + 1,
                                                            ((HaloDev *)ctx->p2p_halo.ptr)->send_ptr, flag);
    FE_LAUNCH_CHECK(ctx);
    FE_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  }
  FE_CUDA(cudaStreamSynchronize(s));  // hd lives on this stack frame
  ctx->p2p_send_sorted = !bad;
  return FE_OK;
}

int allreduce_sum(fe_ctx *ctx, cudaStream_t s, double *dev, int count) {
  if (ctx->nranks <= 1) return FE_OK;
  FE_REQUIRE(ctx->nccl_comm, "allreduce_sum: fe_dist_init was not called");
  FE_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, s));
  return FE_OK;
}

}  // namespace fe

using namespace fe;

extern "C" {

// communication block: red cell[2][nranks][4] | ghost cell[n_ghost][2]
static size_t p2p_block_bytes(int nranks, int n_ghost, size_t *off_ghost) {
  size_t o = (size_t)2 * nranks * 4 * sizeof(uint4);
  o = (o + 255) / 256 * 256;
  *off_ghost = o;
  return o + (size_t)2 * (n_ghost > 0 ? n_ghost : 1) * sizeof(uint4);
}

int fe_dist_p2p_export(fe_ctx *ctx, int32_t n_ghost_dofs, void *handle64) {
  FE_REQUIRE(ctx && handle64 && n_ghost_dofs >= 0, "fe_dist_p2p_export: bad argument");
  FE_REQUIRE(ctx->nranks > 1 && ctx->nranks <= kMaxRanks, "fe_dist_p2p_export: needs fe_dist_init with 2..%d ranks",
             kMaxRanks);
  FE_CUDA(cudaSetDevice(ctx->device));
  FE_CUDA(cudaDeviceSynchronize());
  const bool had_block = ctx->p2p_buf != nullptr;
  for (int r = 0; r < kMaxRanks; ++r) {
    if (ctx->p2p_peer[r] && ctx->p2p_peer[r] != ctx->p2p_buf) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
    ctx->p2p_peer[r] = nullptr;
  }
  if (had_block) {
    // a second mesh in the same process group: nobody frees its old block before EVERY rank has
    // unmapped it (the call is collective, like fe_dist_init)
    int rc = ctx->scratch_b.reserve(256);
    if (rc) return rc;
    FE_CUDA(cudaMemset(ctx->scratch_b.ptr, 0, sizeof(double)));
    if ((rc = allreduce_sum(ctx, nullptr, (double *)ctx->scratch_b.ptr, 1))) return rc;
    FE_CUDA(cudaDeviceSynchronize());
    cudaFree(ctx->p2p_buf);
  }
  ctx->p2p_buf = nullptr;
  size_t off_ghost;
  ctx->p2p_bytes = p2p_block_bytes(ctx->nranks, n_ghost_dofs, &off_ghost);
  FE_CUDA(cudaMalloc(&ctx->p2p_buf, ctx->p2p_bytes));
  FE_CUDA(cudaMemset(ctx->p2p_buf, 0, ctx->p2p_bytes));
  ctx->p2p_n_ghost = n_ghost_dofs;
  cudaIpcMemHandle_t h;
  FE_CUDA(cudaIpcGetMemHandle(&h, ctx->p2p_buf));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return FE_OK;
}

int fe_dist_p2p_import(fe_ctx *ctx, const void *handles) {
  FE_REQUIRE(ctx && handles && ctx->p2p_buf, "fe_dist_p2p_import: call fe_dist_p2p_export first");
  FE_CUDA(cudaSetDevice(ctx->device));
  P2PDev host;
  memset(&host, 0, sizeof(host));
  host.nranks = ctx->nranks;
  host.rank = ctx->rank;
  host.n_ghost = ctx->p2p_n_ghost;
  size_t off_ghost;
  p2p_block_bytes(ctx->nranks, 0, &off_ghost);
  for (int r = 0; r < ctx->nranks; ++r) {
    void *base = ctx->p2p_buf;
    if (r != ctx->rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char *)handles + 64 * r, 64);
      FE_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    }
    ctx->p2p_peer[r] = base;
    host.red[r] = (uint4 *)base;
    host.ghost[r] = (uint4 *)((char *)base + off_ghost);
  }
  if (!ctx->p2p_dev) FE_CUDA(cudaMalloc((void **)&ctx->p2p_dev, sizeof(P2PDev)));
  FE_CUDA(cudaMemcpy(ctx->p2p_dev, &host, sizeof(host), cudaMemcpyHostToDevice));
  return FE_OK;
}

__attribute__((visibility("hidden"))) void fe_dist_teardown(fe_ctx *ctx) {  // library-internal (fe_ctx_destroy), not part of the ABI
  if (ctx) {
    for (int r = 0; r < kMaxRanks; ++r)
      if (ctx->p2p_peer[r] && ctx->p2p_peer[r] != ctx->p2p_buf) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
    if (ctx->p2p_buf) cudaFree(ctx->p2p_buf);
    if (ctx->p2p_dev) cudaFree(ctx->p2p_dev);
    ctx->p2p_halo.release();
  }
  if (ctx && ctx->nccl_comm && g_nccl.CommDestroy) {
    g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
}

int fe_dist_unique_id(void *out128) {
  FE_REQUIRE(out128, "fe_dist_unique_id: NULL");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  FE_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return FE_OK;
}

int fe_dist_init(fe_ctx *ctx, const void *nccl_unique_id, int32_t rank, int32_t nranks) {
  FE_REQUIRE(ctx && nccl_unique_id, "fe_dist_init: NULL argument");
  FE_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "fe_dist_init: bad rank %d / %d", rank, nranks);
  int rc = load_nccl();
  if (rc) return rc;
  FE_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, nccl_unique_id, sizeof(id));
  ncclComm_t comm = nullptr;
  FE_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->nranks = nranks;
  return FE_OK;
}

int fe_dist_pcg(fe_ctx *ctx, void *stream, int32_t n_rows, int32_t n_cols, const int32_t *rowptr,
                const int32_t *colidx, const double *vals, const double *b, double *x, double *work, int32_t n_nbr,
                const int32_t *nbr_rank, const int32_t *send_ptr, const int32_t *send_idx, const int32_t *recv_ptr,
                const int32_t *peer_dst_off, int32_t block_dim, double rtol, int32_t maxit, int32_t fixed_iters, int32_t *iters, double *relres) {
  FE_REQUIRE(ctx, "fe_dist_pcg: NULL ctx");
  FE_REQUIRE(n_nbr == 0 || (nbr_rank && send_ptr && recv_ptr), "fe_dist_pcg: NULL halo description");
  HaloPlan h;
  h.n_nbr = n_nbr;
  h.nbr_rank = nbr_rank;
  h.send_ptr = send_ptr;
  h.send_idx = send_idx;
  h.recv_ptr = recv_ptr;
  h.peer_dst_off = ctx->p2p_dev ? peer_dst_off : nullptr;  // non-NULL selects the peer-memory transport
  if (n_nbr > 0) {
    FE_REQUIRE(n_rows + recv_ptr[n_nbr] == n_cols, "fe_dist_pcg: ghost count %d does not match n_cols - n_rows = %d",
               recv_ptr[n_nbr], n_cols - n_rows);
    FE_REQUIRE(send_ptr[n_nbr] == 0 || send_idx, "fe_dist_pcg: NULL send_idx");
  }
  if (h.peer_dst_off) {
    int rc = upload_p2p_halo(ctx, as_stream(stream), &h);
    if (rc) return rc;
  }
  const bool fixed = fixed_iters > 0;
  return pcg_drive(ctx, as_stream(stream), n_rows, n_cols, rowptr, colidx, vals, b, x, work, &h, block_dim, rtol,
                   fixed ? fixed_iters : maxit, fixed, iters, relres);
}

}  // extern "C"
