// Multi-GPU plumbing used by the PCG driver (solve.cu); implemented in dist.cu.
#pragma once
#include "common.cuh"

namespace fe {

struct HaloPlan {
  int32_t n_nbr = 0;
  const int32_t *nbr_rank = nullptr;  // host [n_nbr]
  const int32_t *send_ptr = nullptr;  // host [n_nbr + 1]
  const int32_t *send_idx = nullptr;  // device [send_ptr[n_nbr]]  owned local DOFs to pack
  const int32_t *recv_ptr = nullptr;  // host [n_nbr + 1]  ghost DOFs of neighbour k land at
                                      // vec[n_rows + recv_ptr[k] .. n_rows + recv_ptr[k+1])
  const int32_t *peer_dst_off = nullptr;  // host [n_nbr]  offset of MY values in neighbour k's ghost block
                                          // (= its recv_ptr entry for me); NULL -> NCCL transport
};

// Pack the interface values of `vec`, exchange them with the neighbours (grouped
// ncclSend/ncclRecv on `s`) and receive straight into the ghost tail of `vec`.
int halo_exchange(fe_ctx *ctx, cudaStream_t s, const HaloPlan *h, double *vec, int32_t n_rows);

// Same exchange without NCCL: k_halo_push stores the interface values straight into the
// neighbours' ghost blocks over the NVLink peer mappings and raises their flags; k_halo_wait_copy
// waits for this rank's neighbours and moves its ghost block behind `vec`.
int halo_exchange_p2p(fe_ctx *ctx, cudaStream_t s, const HaloPlan *h, double *vec, int32_t n_rows);

// In-place sum over all ranks of `count` doubles in device memory (ncclAllReduce on `s`).
int allreduce_sum(fe_ctx *ctx, cudaStream_t s, double *dev, int count);

// host driver in solve.cu
int pcg_drive(fe_ctx *ctx, cudaStream_t s, int32_t n_rows, int32_t n_cols, const int32_t *rowptr,
              const int32_t *colidx, const double *vals, const double *b, double *x, double *work,
              const HaloPlan *halo, int block_dim, double rtol, int32_t maxit, bool fixed, int32_t *iters_out,
              double *relres_out);

}  // namespace fe
