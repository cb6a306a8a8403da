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
struct HaloDev {  // device copy of the halo description, in ctx->p2p_halo
  int n_nbr, n_send, n_ghost, ticket;
  int nbr_rank[kMaxRanks];
  int send_ptr[kMaxRanks + 1];
  int dst_off[kMaxRanks];
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Interface values -> the neighbours' ghost blocks (peer stores over NVLink), then their flags.
__global__ void __launch_bounds__(256) k_halo_push(P2PDev *pp, HaloDev *hd, const int32_t *__restrict__ send_idx,
                                                  const double *__restrict__ vec) {
  __shared__ bool is_last;
  const int n_send = hd->n_send;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_send; i += gridDim.x * blockDim.x) {
    int k = 0;
    while (i >= hd->send_ptr[k + 1]) ++k;
    pp->ghost[hd->nbr_rank[k]][hd->dst_off[k] + (i - hd->send_ptr[k])] = vec[send_idx[i]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(&hd->ticket, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence_system();
    const unsigned long long seq = pp->halo_seq + 1;
    for (int k = 0; k < hd->n_nbr; ++k) st_release_sys_u64(pp->hflags[hd->nbr_rank[k]] + pp->rank, seq);
    pp->halo_seq = seq;
    hd->ticket = 0;
  }
}

// Wait for every neighbour's delivery of the current exchange, then ghost block -> vec tail.
__global__ void __launch_bounds__(256) k_halo_wait_copy(const P2PDev *pp, const HaloDev *hd,
                                                       double *__restrict__ vec_tail) {
  const unsigned long long seq = pp->halo_seq;
  if (threadIdx.x < hd->n_nbr) {
    const unsigned long long *f = pp->hflags[pp->rank] + hd->nbr_rank[threadIdx.x];
    while (ld_acquire_sys_u64(f) < seq) {
    }
  }
  __syncthreads();
  const volatile double *g = pp->ghost[pp->rank];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hd->n_ghost; i += gridDim.x * blockDim.x) vec_tail[i] = g[i];
}

int halo_exchange_p2p(fe_ctx *ctx, cudaStream_t s, const HaloPlan *h, double *vec, int32_t n_rows) {
  if (!h || h->n_nbr == 0) return FE_OK;
  FE_REQUIRE(ctx->p2p_dev && ctx->p2p_halo.ptr, "halo_exchange_p2p: peer memory is not set up");
  HaloDev *hd = (HaloDev *)ctx->p2p_halo.ptr;
  const int32_t n_send = h->send_ptr[h->n_nbr], n_ghost = h->recv_ptr[h->n_nbr];
  if (n_send > 0) {
    int grid = grid_for(n_send, 256);
    if (grid > 64) grid = 64;
    k_halo_push<<<grid, 256, 0, s>>>(ctx->p2p_dev, hd, h->send_idx, vec);
    FE_LAUNCH_CHECK(ctx);
  }
  if (n_ghost > 0) {
    int grid = grid_for(n_ghost, 256);
    if (grid > 64) grid = 64;
    k_halo_wait_copy<<<grid, 256, 0, s>>>(ctx->p2p_dev, hd, vec + n_rows);
    FE_LAUNCH_CHECK(ctx);
  }
  return FE_OK;
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
  int rc = ctx->p2p_halo.reserve(sizeof(HaloDev));
  if (rc) return rc;
  FE_CUDA(cudaMemcpyAsync(ctx->p2p_halo.ptr, &hd, sizeof(hd), cudaMemcpyHostToDevice, s));
  FE_CUDA(cudaStreamSynchronize(s));  // hd lives on this stack frame
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

static size_t p2p_block_bytes(int nranks, int n_ghost, size_t *off_rflags, size_t *off_hflags, size_t *off_ghost) {
  size_t o = (size_t)2 * nranks * 4 * sizeof(double);
  *off_rflags = o;
  o += (size_t)2 * nranks * sizeof(unsigned long long);
  *off_hflags = o;
  o += (size_t)nranks * sizeof(unsigned long long);
  o = (o + 255) / 256 * 256;
  *off_ghost = o;
  return o + (size_t)(n_ghost > 0 ? n_ghost : 1) * sizeof(double);
}

int fe_dist_p2p_export(fe_ctx *ctx, int32_t n_ghost_dofs, void *handle64) {
  FE_REQUIRE(ctx && handle64 && n_ghost_dofs >= 0, "fe_dist_p2p_export: bad argument");
  FE_REQUIRE(ctx->nranks > 1 && ctx->nranks <= kMaxRanks, "fe_dist_p2p_export: needs fe_dist_init with 2..%d ranks",
             kMaxRanks);
  FE_CUDA(cudaSetDevice(ctx->device));
  FE_CUDA(cudaDeviceSynchronize());
  for (int r = 0; r < kMaxRanks; ++r) {
    if (ctx->p2p_peer[r] && ctx->p2p_peer[r] != ctx->p2p_buf) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
    ctx->p2p_peer[r] = nullptr;
  }
  if (ctx->p2p_buf) cudaFree(ctx->p2p_buf);
  ctx->p2p_buf = nullptr;
  size_t o1, o2, o3;
  ctx->p2p_bytes = p2p_block_bytes(ctx->nranks, n_ghost_dofs, &o1, &o2, &o3);
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
  size_t o1, o2, o3;
  p2p_block_bytes(ctx->nranks, 0, &o1, &o2, &o3);
  for (int r = 0; r < ctx->nranks; ++r) {
    void *base = ctx->p2p_buf;
    if (r != ctx->rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char *)handles + 64 * r, 64);
      FE_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    }
    ctx->p2p_peer[r] = base;
    host.slots[r] = (double *)base;
    host.rflags[r] = (unsigned long long *)((char *)base + o1);
    host.hflags[r] = (unsigned long long *)((char *)base + o2);
    host.ghost[r] = (double *)((char *)base + o3);
  }
  if (!ctx->p2p_dev) FE_CUDA(cudaMalloc((void **)&ctx->p2p_dev, sizeof(P2PDev)));
  FE_CUDA(cudaMemcpy(ctx->p2p_dev, &host, sizeof(host), cudaMemcpyHostToDevice));
  return FE_OK;
}

void fe_dist_teardown(fe_ctx *ctx) {
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
