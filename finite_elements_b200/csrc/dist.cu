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

int allreduce_sum(fe_ctx *ctx, cudaStream_t s, double *dev, int count) {
  if (ctx->nranks <= 1) return FE_OK;
  FE_REQUIRE(ctx->nccl_comm, "allreduce_sum: fe_dist_init was not called");
  FE_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, s));
  return FE_OK;
}

}  // namespace fe

using namespace fe;

extern "C" {

void fe_dist_teardown(fe_ctx *ctx) {
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
                int32_t block_dim, double rtol, int32_t maxit, int32_t fixed_iters, int32_t *iters, double *relres) {
  FE_REQUIRE(ctx, "fe_dist_pcg: NULL ctx");
  FE_REQUIRE(n_nbr == 0 || (nbr_rank && send_ptr && recv_ptr), "fe_dist_pcg: NULL halo description");
  HaloPlan h;
  h.n_nbr = n_nbr;
  h.nbr_rank = nbr_rank;
  h.send_ptr = send_ptr;
  h.send_idx = send_idx;
  h.recv_ptr = recv_ptr;
  if (n_nbr > 0) {
    FE_REQUIRE(n_rows + recv_ptr[n_nbr] == n_cols, "fe_dist_pcg: ghost count %d does not match n_cols - n_rows = %d",
               recv_ptr[n_nbr], n_cols - n_rows);
    FE_REQUIRE(send_ptr[n_nbr] == 0 || send_idx, "fe_dist_pcg: NULL send_idx");
  }
  const bool fixed = fixed_iters > 0;
  return pcg_drive(ctx, as_stream(stream), n_rows, n_cols, rowptr, colidx, vals, b, x, work, &h, block_dim, rtol,
                   fixed ? fixed_iters : maxit, fixed, iters, relres);
}

}  // extern "C"
