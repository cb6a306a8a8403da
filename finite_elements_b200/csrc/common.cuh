// Internals shared by every translation unit of libfe_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/fe_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libfe_b200 is written for sm_100a (B200) only"
#endif

namespace fe {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- thread-local error text --------------------------------------------------------
char *last_error_buf();
int fail(int code, const char *fmt, ...);

#define FE_CUDA(call)                                                                     \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess)                                                               \
      return fe::fail(FE_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,   \
                      cudaGetErrorString(e__));                                           \
  } while (0)

#define FE_LAUNCH_CHECK(ctx)                                                              \
  do {                                                                                    \
    (ctx)->launches++;                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess)                                                               \
      return fe::fail(FE_ERR_CUDA, "kernel launch failed at %s:%d: %s", __FILE__,         \
                      __LINE__, cudaGetErrorString(e__));                                 \
  } while (0)

#define FE_REQUIRE(cond, ...)                                                             \
  do {                                                                                    \
    if (!(cond)) return fe::fail(FE_ERR_ARG, __VA_ARGS__);                                \
  } while (0)

// grow-only device scratch buffer
struct Scratch {
  void *ptr = nullptr;
  size_t bytes = 0;
  int reserve(size_t need) {
    if (need <= bytes) return FE_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    size_t want = need + need / 8 + 256;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) return fail(FE_ERR_CUDA, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    bytes = want;
    return FE_OK;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
};

}  // namespace fe

namespace fe {
// Peer-memory communication block (dist.cu).  Every rank owns one cudaMalloc'ed buffer, exported
// with CUDA IPC and mapped by all peers of the node (NVLink / NVSwitch).  Everything in it is made
// of 16-byte "LL" cells {lo32(value), seq, hi32(value), seq}: a double travels together with the
// sequence number of the exchange it belongs to, so the receiver needs no separate flag and the
// sender needs no fence -- each 8-byte half is self-validating (8-byte stores are atomic on
// NVLink; the protocol NCCL calls LL).  Cells are double-buffered by the parity of the sequence
// number, which is enough because two exchanges of the same parity are always separated by a
// full round trip between the two ranks involved.
//   red    cell[2][nranks][4]   per-source partial sums of an all-reduce
//   ghost  cell[2][n_ghost]     ghost values, written by the owning neighbours
constexpr int kMaxRanks = 16;
struct P2PDev {
  int nranks, rank, n_ghost, pad;
  unsigned red_seq;   // all-reduces pushed by this rank so far
  unsigned halo_seq;  // halo exchanges pushed by this rank so far
  uint4 *red[kMaxRanks];
  uint4 *ghost[kMaxRanks];
};
// Device copy of a rank's halo description (ctx->p2p_halo), read by the kernels that push
// interface values: ghost j of this rank (parity par) lives in cell ghost[rank][2 j + par].
struct HaloDev {
  int n_nbr, n_send, n_ghost;
  unsigned ticket;
  int nbr_rank[kMaxRanks];
  int send_ptr[kMaxRanks + 1];
  int dst_off[kMaxRanks];
};
}  // namespace fe

struct fe_ctx {
  int device = 0;
  int num_sms = fe::kNumSMs;
  int64_t launches = 0;
  fe::Scratch scratch_a;   // material tables, BC flags, scan block sums
  fe::Scratch scratch_b;   // reduction partials / PCG scalars
  fe::Scratch scratch_c;   // node-level block pattern of the streamed SpMV
  fe::Scratch scratch_g;   // fe_tet_assemble: per-element gradient table (128 B per element)
  fe::Scratch bc_map;      // fe_dirichlet_apply: column -> condition index + 1 (kept all-zero between calls)
  fe::Scratch scratch_p;   // persistent PCG kernel: barrier flags (zeroed on allocation) + per-CTA partials
  bool p2p_send_sorted = false;  // every neighbour's send list is ascending (checked when the halo is uploaded)
  // fe_pcg_cache_pattern: the caller vouches that the CSR pattern at (rowptr, colidx) is immutable
  // while `token` stays the same, so the block pattern in scratch_c is reused across solves
  const void *bp_rowptr = nullptr, *bp_colidx = nullptr;
  int64_t bp_token = 0, bp_built_token = 0;
  int32_t bp_built_rows = -1;
  int bp_max_deg = 0;
  void *pinned = nullptr;  // small pinned host buffer for scalar read-back
  void *pcg_graph = nullptr;  // cached cudaGraphExec_t of one PCG iteration chunk
  const void *pcg_graph_key[16] = {nullptr};  // every pointer / size the captured launches bake in
  void *work_stream = nullptr;  // cudaStream_t used when the caller passes a default stream
  void *work_event = nullptr;   // cudaEvent_t ordering work_stream after the caller's stream
  // multi-GPU
  void *nccl_comm = nullptr;
  int rank = 0, nranks = 1;
  fe::Scratch halo_send, halo_recv;
  // NCCL-free peer-memory path (fe_dist_p2p_export / _import)
  void *p2p_buf = nullptr;          // this rank's communication block
  size_t p2p_bytes = 0;
  void *p2p_peer[fe::kMaxRanks] = {nullptr};  // mapped peer blocks (self = p2p_buf)
  fe::P2PDev *p2p_dev = nullptr;    // device copy of the pointer table + sequence counters
  int p2p_n_ghost = 0;
  fe::Scratch p2p_halo;             // device copies of send_ptr / dst_off / neighbour ranks
};

namespace fe {

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int grid_for(int64_t work, int block) {
  int64_t g = (work + block - 1) / block;
  return (int)(g < 1 ? 1 : g);
}

// Exclusive scan of int32 counts -> int32 offsets (out[n] = total), total also returned as
// int64 in *total_dev (device).  in/out may alias.  Uses ctx->scratch_a.
int exclusive_scan_i32(fe_ctx *ctx, cudaStream_t st, const int32_t *in, int32_t *out, int64_t n,
                       int64_t *total_dev);

// ---- device helpers -----------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum in a fixed order (warp shuffles, then warp 0 over the per-warp partials):
// the result depends only on blockDim, never on scheduling.  Valid in thread 0.
template <int BLOCK>
__device__ __forceinline__ double block_sum(double v, double *smem /* BLOCK/32 doubles */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) smem[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (lane < BLOCK / 32) ? smem[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

__device__ __forceinline__ int2 ldg_nc_int2(const int2 *p) {
  int2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

// ---- LL cells (see P2PDev) -------------------------------------------------------------------
__device__ __forceinline__ void ll_store(uint4 *cell, double v, unsigned seq) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(cell), "r"((unsigned)b), "r"(seq),
               "r"((unsigned)(b >> 32)), "r"(seq)
               : "memory");
}
__device__ __forceinline__ bool ll_try(const uint4 *cell, unsigned seq, double *out) {
  unsigned a, fa, b, fb;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(cell) : "memory");
  *out = __longlong_as_double((long long)(((unsigned long long)b << 32) | a));
  return fa == seq && fb == seq;
}
// Spins until the cell carries `seq`.  A peer that never delivers (crashed rank, mismatched call
// sequence) must not hang the GPU: after ~30 s of back-off polling the wait gives up and returns
// NaN, which the PCG kernels turn into FE_ERR_BREAKDOWN (p.Ap not finite).
__device__ __forceinline__ double ll_wait(const uint4 *cell, unsigned seq) {
  double v;
#pragma unroll 1
  for (int fast = 0; fast < 4096; ++fast)
    if (ll_try(cell, seq, &v)) return v;
#pragma unroll 1
  for (long long slow = 0; slow < (1ll << 27); ++slow) {
    if (ll_try(cell, seq, &v)) return v;
    __nanosleep(200);
  }
  return __longlong_as_double(0x7ff8000000000000ll);
}
// Stores this rank's interface values of `vec` into the neighbours' ghost cells (peer stores over
// NVLink); `t0`/`stride` spread the send list over the calling threads.
__device__ __forceinline__ void halo_push(const P2PDev *pp, const HaloDev *hd, const int32_t *__restrict__ send_idx,
                                          const double *__restrict__ vec, unsigned seq, int t0, int stride) {
  const int n_send = hd->n_send;
  for (int i = t0; i < n_send; i += stride) {
    int k = 0;
    while (i >= hd->send_ptr[k + 1]) ++k;
    ll_store(pp->ghost[hd->nbr_rank[k]] + 2 * (size_t)(hd->dst_off[k] + (i - hd->send_ptr[k])) + (seq & 1),
             vec[send_idx[i]], seq);
  }
}
#endif

}  // namespace fe
