// Streaming SpMV for node-blocked (2 DOF per node) matrices -- the PCG hot kernel.
//
// Included by solve.cu after PcgState / publish_and_reduce.
//
// The CSR of an fe_plan with dim == 2 is a block-CSR in disguise: node i owns rows 2i, 2i+1,
// whose values are ONE contiguous run of 4*valence doubles, and whose columns are the
// (2m, 2m+1) pairs of its neighbour nodes m.  k_block_pattern extracts the node-level pattern
// once per solve:  bptr[i] = rowptr[2i] / 4,  bidx[b] = colidx[rowptr[2i] + 2j] / 2  -- 4 bytes
// per 2x2 block instead of 16.
//
// k_spmv_stream then runs persistent CTAs (2 per SM).  A tile = T consecutive nodes = one
// contiguous slice of vals (32 B per block) and one of bidx (4 B per block).  Thread 0 hands both
// slices of tile t+1 to the TMA engine (cp.async.bulk global -> shared, completion counted on an
// mbarrier) while all threads compute tile t out of shared memory, so HBM streaming never waits
// for the per-row dependent work (pointer look-up -> column -> gather of x).  The tile's bptr
// slice travels with cp.async (LDGSTS).  8 lanes per node, one 2x2 block per lane:
// LDS.128 x2 (vals), LDS.32 (column), LDG.128 (x, L1/L2 resident), 4 FMAs; fused p.q partials.
//
// HBM bytes per call:  8 nnz (vals) + nnz (bidx) + 2 n (bptr) + 8 n (x) + 8 n (y).
#pragma once
// (included from inside namespace fe)

__global__ void __launch_bounds__(256) k_block_pattern(int32_t n_nodes, const int32_t *__restrict__ rowptr,
                                                      const int32_t *__restrict__ colidx,
                                                      int32_t *__restrict__ bptr, int32_t *__restrict__ bidx,
                                                      int *__restrict__ max_deg, unsigned char *__restrict__ diag_slot) {
  // 8 lanes per node
  const int64_t node = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 3;
  const int lane = threadIdx.x & 7;
  if (node > n_nodes) return;
  if (node == n_nodes) {
    if (lane == 0) bptr[n_nodes] = rowptr[2 * n_nodes] >> 2;
    return;
  }
  const int32_t s0 = rowptr[2 * node], s2 = rowptr[2 * node + 2];
  const int deg = (s2 - s0) >> 2;
  if (lane == 0) {
    bptr[node] = s0 >> 2;
    if (deg > *reinterpret_cast<volatile int *>(max_deg)) atomicMax(max_deg, deg);  // rarely taken
  }
  bool found = false;
  for (int j = lane; j < deg; j += 8) {
    const int32_t m = colidx[s0 + 2 * j] >> 1;
    bidx[(s0 >> 2) + j] = m;
    if (m == node && j < 255) {  // position of the diagonal block: k_dinv_b2 reads it instead of searching
      diag_slot[node] = (unsigned char)j;
      found = true;
    }
  }
  if (!__any_sync(0xffu << ((threadIdx.x & 31) & ~7), found) && lane == 0) diag_slot[node] = 255;  // not stored
}

// dinv of a 2-DOF-per-node matrix from the cached diagonal positions: two 8-byte loads per node instead
// of one thread scanning a whole row (k_extract_dinv: 0.43 ms at 16.8 M rows).
__global__ void __launch_bounds__(256) k_dinv_b2(int32_t n_nodes, const int32_t *__restrict__ bptr,
                                                const unsigned char *__restrict__ diag_slot,
                                                const double *__restrict__ vals, double *__restrict__ dinv,
                                                int *__restrict__ breakdown) {
  const int32_t node = blockIdx.x * 256 + threadIdx.x;
  if (node >= n_nodes) return;
  const int32_t b0 = bptr[node], deg = bptr[node + 1] - b0;
  const int k = diag_slot[node];
  double d0 = 0.0, d1 = 0.0;
  if (k < deg) {
    d0 = vals[4 * (int64_t)b0 + 2 * k];
    d1 = vals[4 * (int64_t)b0 + 2 * deg + 2 * k + 1];
  }
  if (!(d0 > 0.0) || !(d1 > 0.0)) *breakdown = 1;  // not SPD (e.g. no Dirichlet condition at all)
  *reinterpret_cast<double2 *>(dinv + 2 * (int64_t)node) = make_double2(1.0 / d0, 1.0 / d1);
}


constexpr int kStreamConsumerWarps = 15;                        // + 1 producer warp
constexpr int kStreamThreads = (kStreamConsumerWarps + 1) * 32;  // 512
constexpr int kStreamGroups = kStreamConsumerWarps * 4;          // 8-lane groups = nodes per pass
constexpr int kStreamTile = 2 * kStreamGroups;                   // 120 nodes per tile, 2 passes
constexpr int kStreamStages = 3;
constexpr int kStreamPtrInts = (kStreamTile + 1 + 3) & ~3;       // bptr slice copied per tile (16 B units)

// smem: [full[S], empty[S] mbarriers][S x kStreamPtrInts ints][S x cap x 32 B vals][S x (cap+8) x 4 B bidx]
// Value pair of ghost node g (columns n_nodes + g) straight from this rank's LL cells: the SpMV of
// the peer-memory transport never waits for a separate exchange kernel -- only the lanes whose
// column is a ghost wait, and only until that one neighbour value has landed.
__device__ __forceinline__ double2 ghost_pair(const uint4 *cells /* + parity */, int32_t g, unsigned seq) {
  return make_double2(ll_wait(cells + 4 * (size_t)g, seq), ll_wait(cells + 4 * (size_t)g + 2, seq));
}

// HALO: multi-GPU peer-memory transport.  Every CTA first pushes its share of this rank's interface
// values of x to the neighbours' ghost cells (halo_push), then streams its tiles; columns >= n_nodes
// are read from the ghost cells (ghost_pair).  The last CTA to finish bumps the exchange counter.
template <bool DOT, bool HALO>
__global__ void __launch_bounds__(kStreamThreads, 2) k_spmv_stream(
    int32_t n_nodes, int /*T == kStreamTile*/, int cap /* blocks per stage */, const int32_t *__restrict__ bptr,
    const int32_t *__restrict__ bidx, const double *__restrict__ vals, const double *__restrict__ x,
    double *__restrict__ y, double *__restrict__ partials, PcgState *__restrict__ st, P2PDev *pp, HaloDev *hd,
    const int32_t *__restrict__ send_idx) {
  constexpr int T = kStreamTile, S = kStreamStages, GROUPS = kStreamGroups, PASSES = 2;
  __shared__ double red[kStreamThreads / 32];
  if (DOT && (st->converged | st->breakdown)) return;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);
  uint64_t *empty = full + S;
  constexpr size_t kHdr = 128;
  int32_t *ptr_s = reinterpret_cast<int32_t *>(smem + kHdr);
  constexpr size_t ptr_bytes = (kHdr + (size_t)S * kStreamPtrInts * sizeof(int32_t) + 127) / 128 * 128;
  double *vals_s = reinterpret_cast<double *>(smem + ptr_bytes);
  int32_t *idx_s = reinterpret_cast<int32_t *>(smem + ptr_bytes + (size_t)S * cap * 32);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_tiles = (n_nodes + T - 1) / T;
  unsigned hseq = 0;
  const uint4 *gcells = nullptr;
  if (HALO) {
    hseq = pp->halo_seq + 1;
    gcells = pp->ghost[pp->rank] + (hseq & 1);
    // only the first CTAs push: they are dispatched first, so a CTA that waits for a neighbour's
    // values never waits for a push that is itself waiting for an SM
    const int pushers = min((int)gridDim.x, 32);
    if ((int)blockIdx.x < pushers)
      halo_push(pp, hd, send_idx, x, hseq, blockIdx.x * kStreamThreads + tid, pushers * kStreamThreads);
  }
  if (tid == 0) {
    for (int i = 0; i < S; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], kStreamConsumerWarps);
    }
    ptx::mbar_init_fence();
  }
  __syncthreads();

  double dot = 0.0;
  if (warp == kStreamConsumerWarps) {
    // ===== producer warp: one lane feeds the ring with TMA bulk loads =====
    if ((tid & 31) == 0) {
      int j = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
        const int stage = j % S, use = j / S;
        if (use > 0) ptx::mbar_wait(&empty[stage], (uint32_t)((use - 1) & 1));  // consumers released it
        const int32_t n0 = tile * T, n1 = min(n0 + T, n_nodes);
        const int32_t b0 = __ldg(bptr + n0), b1 = __ldg(bptr + n1);
        const int32_t a0 = b0 & ~3;  // bidx copy is 16-byte aligned on both sides
        const uint32_t nb = (uint32_t)(b1 - b0);
        const uint32_t ni = ((uint32_t)(b1 - a0) + 3u) & ~3u;
        ptx::mbar_expect_tx(&full[stage], kStreamPtrInts * 4u + nb * 32u + (nb ? ni * 4u : 0u));
        ptx::bulk_load(ptr_s + stage * kStreamPtrInts, bptr + n0, kStreamPtrInts * 4u, &full[stage]);
        if (nb) {
          ptx::bulk_load(vals_s + (size_t)stage * cap * 4, vals + 4 * (int64_t)b0, nb * 32u, &full[stage]);
          ptx::bulk_load(idx_s + (size_t)stage * (cap + 8), bidx + a0, ni * 4u, &full[stage]);
        }
      }
    }
  } else {
    // ===== consumer warps: 8 lanes per node, one 2x2 block per lane =====
    const int grp = tid >> 3, lane = tid & 7;
    const double2 *x2 = reinterpret_cast<const double2 *>(x);
    double2 *y2 = reinterpret_cast<double2 *>(y);
    int j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      const int stage = j % S, use = j / S;
      ptx::mbar_wait(&full[stage], (uint32_t)(use & 1));
      const int32_t n0 = tile * T;
      const int nn = min(T, n_nodes - n0);
      const int32_t *ps = ptr_s + stage * kStreamPtrInts;
      const int32_t b0 = ps[0];
      const double2 *vs = reinterpret_cast<const double2 *>(vals_s + (size_t)stage * cap * 4);
      const int32_t *is = idx_s + (size_t)stage * (cap + 8) + (b0 - (b0 & ~3));

      int32_t s[PASSES], deg[PASSES], c[PASSES];
      double2 v0[PASSES], v1[PASSES], xv[PASSES];
#pragma unroll
      for (int p = 0; p < PASSES; ++p) {  // all shared-memory reads of the tile first ...
        const int i = grp + p * GROUPS;
        s[p] = 0;
        deg[p] = 0;
        if (i < nn) {
          s[p] = ps[i] - b0;
          deg[p] = ps[i + 1] - ps[i];
        }
        const bool act = lane < deg[p];
        c[p] = act ? is[s[p] + lane] : 0;
        v0[p] = act ? vs[2 * s[p] + lane] : make_double2(0.0, 0.0);
        v1[p] = act ? vs[2 * s[p] + deg[p] + lane] : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int p = 0; p < PASSES; ++p) xv[p] = __ldg(x2 + c[p]);  // ... then every gather of x in flight
      if (HALO) {
#pragma unroll
        for (int p = 0; p < PASSES; ++p)
          if (c[p] >= n_nodes) xv[p] = ghost_pair(gcells, c[p] - n_nodes, hseq);
      }
      double a0[PASSES], a1[PASSES];
#pragma unroll
      for (int p = 0; p < PASSES; ++p) {
        a0[p] = v0[p].x * xv[p].x + v0[p].y * xv[p].y;
        a1[p] = v1[p].x * xv[p].x + v1[p].y * xv[p].y;
        for (int k = lane + 8; k < deg[p]; k += 8) {  // valence > 8
          const int32_t ck = is[s[p] + k];
          const double2 xx = (HALO && ck >= n_nodes) ? ghost_pair(gcells, ck - n_nodes, hseq) : __ldg(x2 + ck);
          const double2 w0 = vs[2 * s[p] + k], w1 = vs[2 * s[p] + deg[p] + k];
          a0[p] += w0.x * xx.x + w0.y * xx.y;
          a1[p] += w1.x * xx.x + w1.y * xx.y;
        }
      }
      // this warp is done reading the stage: hand it back to the producer
      __syncwarp();
      if ((tid & 31) == 0) ptx::mbar_arrive(&empty[stage]);
#pragma unroll
      for (int p = 0; p < PASSES; ++p) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          a0[p] += __shfl_xor_sync(0xffffffffu, a0[p], o);
          a1[p] += __shfl_xor_sync(0xffffffffu, a1[p], o);
        }
        const int i = grp + p * GROUPS;
        if (lane == 0 && i < nn) {
          y2[n0 + i] = make_double2(a0[p], a1[p]);
          if (DOT) {
            const double2 xs = x2[n0 + i];
            dot += a0[p] * xs.x + a1[p] * xs.y;
          }
        }
      }
    }
  }
  if (DOT) {
    const double loc[1] = {dot};
    publish_and_reduce<1, kStreamThreads>(loc, partials, st, 0, red);
  }
  if (HALO) {  // every CTA has read halo_seq once the last one gets here
    __syncthreads();
    if (tid == 0 && atomicAdd(&hd->ticket, 1u) == gridDim.x - 1) {
      hd->ticket = 0;
      pp->halo_seq = hseq;
    }
  }
}

// ---------------------------------------------------------------------------------------
// The same streaming structure for scalar CSR (1 DOF per node: magnetostatics).  rowptr / colidx
// ARE the node-level pattern, so nothing is pre-extracted.  A tile = T1 consecutive rows = one
// contiguous slice of vals (8 B per entry) and one of colidx (4 B per entry); rows are lighter than
// 2x2-block rows (7 entries = 84 B on a structured mesh), so a tile has 4 passes of 60 rows.
// 8 lanes per row, one entry per lane: LDS.64 (value), LDS.32 (column), LDG.64 (x), one FMA.
// HBM bytes per call: 12 nnz + 4 n (rowptr) + 8 n (x) + 8 n (y) -- SURVEY §8d's SpMV figure.
// ---------------------------------------------------------------------------------------
// PASSES rows per 8-lane group and tile: 4 (240 rows) for the 7-entry rows of a scalar triangle mesh; 2 or 1
// when the rows are long (45-entry rows of a tetrahedral mesh: a 240-row tile would not fit the ring).
__host__ __device__ constexpr int stream1_tile(int passes) { return passes * kStreamGroups; }
__host__ __device__ constexpr int stream1_ptr_ints(int passes) { return (stream1_tile(passes) + 1 + 3) & ~3; }

template <bool DOT, bool HALO, int PASSES>
__global__ void __launch_bounds__(kStreamThreads, 2) k_spmv_stream1(
    int32_t n_rows, int cap /* entries per stage */, const int32_t *__restrict__ rowptr,
    const int32_t *__restrict__ colidx, const double *__restrict__ vals, const double *__restrict__ x,
    double *__restrict__ y, double *__restrict__ partials, PcgState *__restrict__ st, P2PDev *pp, HaloDev *hd,
    const int32_t *__restrict__ send_idx) {
  constexpr int T = stream1_tile(PASSES), S = kStreamStages, GROUPS = kStreamGroups;
  constexpr int kStream1PtrInts = stream1_ptr_ints(PASSES);
  __shared__ double red[kStreamThreads / 32];
  if (DOT && (st->converged | st->breakdown)) return;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);
  uint64_t *empty = full + S;
  constexpr size_t kHdr = 128;
  int32_t *ptr_s = reinterpret_cast<int32_t *>(smem + kHdr);
  constexpr size_t ptr_bytes = (kHdr + (size_t)S * kStream1PtrInts * sizeof(int32_t) + 127) / 128 * 128;
  double *vals_s = reinterpret_cast<double *>(smem + ptr_bytes);                       // S x (cap + 2) doubles
  int32_t *idx_s = reinterpret_cast<int32_t *>(smem + ptr_bytes + (size_t)S * (cap + 2) * 8);  // S x (cap + 8) ints

  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_tiles = (n_rows + T - 1) / T;
  unsigned hseq = 0;
  const uint4 *gcells = nullptr;
  if (HALO) {
    hseq = pp->halo_seq + 1;
    gcells = pp->ghost[pp->rank] + (hseq & 1);
    const int pushers = min((int)gridDim.x, 32);
    if ((int)blockIdx.x < pushers)
      halo_push(pp, hd, send_idx, x, hseq, blockIdx.x * kStreamThreads + tid, pushers * kStreamThreads);
  }
  if (tid == 0) {
    for (int i = 0; i < S; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], kStreamConsumerWarps);
    }
    ptx::mbar_init_fence();
  }
  __syncthreads();

  double dot = 0.0;
  if (warp == kStreamConsumerWarps) {
    // producer warp: lane 0 hands the tile's pointer / value / index slices to the TMA engine.  The
    // 16-byte granularity of bulk copies would read a few elements past the END of the caller's
    // arrays on the last two tiles, so those are copied by the whole warp with exact bounds.
    const int pl = tid & 31;
    int j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      const int stage = j % S, use = j / S;
      if (use > 0) {
        if (pl == 0) ptx::mbar_wait(&empty[stage], (uint32_t)((use - 1) & 1));
        __syncwarp();
      }
      const int32_t r0 = tile * T, r1 = min(r0 + T, n_rows);
      const int32_t b0 = __ldg(rowptr + r0), b1 = __ldg(rowptr + r1);
      const int32_t av = b0 & ~1, ai = b0 & ~3;  // 16-byte aligned starts of the two slices
      if (tile < n_tiles - 2) {
        if (pl == 0) {
          const uint32_t nv = ((uint32_t)(b1 - av) + 1u) & ~1u, ni = ((uint32_t)(b1 - ai) + 3u) & ~3u;
          const bool any = b1 > b0;
          ptx::mbar_expect_tx(&full[stage], kStream1PtrInts * 4u + (any ? nv * 8u + ni * 4u : 0u));
          ptx::bulk_load(ptr_s + stage * kStream1PtrInts, rowptr + r0, kStream1PtrInts * 4u, &full[stage]);
          if (any) {
            ptx::bulk_load(vals_s + (size_t)stage * (cap + 2), vals + av, nv * 8u, &full[stage]);
            ptx::bulk_load(idx_s + (size_t)stage * (cap + 8), colidx + ai, ni * 4u, &full[stage]);
          }
        }
      } else {
        for (int i = pl; i <= r1 - r0; i += 32) ptr_s[stage * kStream1PtrInts + i] = __ldg(rowptr + r0 + i);
        for (int i = pl; i < b1 - b0; i += 32) {
          vals_s[(size_t)stage * (cap + 2) + (b0 - av) + i] = __ldg(vals + b0 + i);
          idx_s[(size_t)stage * (cap + 8) + (b0 - ai) + i] = __ldg(colidx + b0 + i);
        }
        __syncwarp();                                   // the lanes' stores ordered before lane 0's arrive
        if (pl == 0) ptx::mbar_arrive(&full[stage]);    // (release) -- no transaction bytes on this phase
      }
    }
  } else {
    const int grp = tid >> 3, lane = tid & 7;
    int j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      const int stage = j % S, use = j / S;
      ptx::mbar_wait(&full[stage], (uint32_t)(use & 1));
      const int32_t r0 = tile * T;
      const int nn = min(T, n_rows - r0);
      const int32_t *ps = ptr_s + stage * kStream1PtrInts;
      const int32_t b0 = ps[0];
      const double *vs = vals_s + (size_t)stage * (cap + 2) + (b0 - (b0 & ~1));
      const int32_t *is = idx_s + (size_t)stage * (cap + 8) + (b0 - (b0 & ~3));

      int32_t s[PASSES], deg[PASSES], c[PASSES];
      double v[PASSES], xv[PASSES];
#pragma unroll
      for (int p = 0; p < PASSES; ++p) {  // all shared-memory reads of the tile first ...
        const int i = grp + p * GROUPS;
        s[p] = 0;
        deg[p] = 0;
        if (i < nn) {
          s[p] = ps[i] - b0;
          deg[p] = ps[i + 1] - ps[i];
        }
        const bool act = lane < deg[p];
        c[p] = act ? is[s[p] + lane] : 0;
        v[p] = act ? vs[s[p] + lane] : 0.0;
      }
#pragma unroll
      for (int p = 0; p < PASSES; ++p) xv[p] = __ldg(x + c[p]);  // ... then every gather of x in flight
      if (HALO) {
#pragma unroll
        for (int p = 0; p < PASSES; ++p)
          if (c[p] >= n_rows) xv[p] = ll_wait(gcells + 2 * (size_t)(c[p] - n_rows), hseq);
      }
      double a[PASSES];
#pragma unroll
      for (int p = 0; p < PASSES; ++p) {
        a[p] = v[p] * xv[p];
        for (int k = lane + 8; k < deg[p]; k += 8) {  // more than 8 entries in the row
          const int32_t ck = is[s[p] + k];
          const double xx = (HALO && ck >= n_rows) ? ll_wait(gcells + 2 * (size_t)(ck - n_rows), hseq) : __ldg(x + ck);
          a[p] += vs[s[p] + k] * xx;
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) ptx::mbar_arrive(&empty[stage]);
#pragma unroll
      for (int p = 0; p < PASSES; ++p) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) a[p] += __shfl_xor_sync(0xffffffffu, a[p], o);
        const int i = grp + p * GROUPS;
        if (lane == 0 && i < nn) {
          y[r0 + i] = a[p];
          if (DOT) dot += a[p] * x[r0 + i];
        }
      }
    }
  }
  if (DOT) {
    const double loc[1] = {dot};
    publish_and_reduce<1, kStreamThreads>(loc, partials, st, 0, red);
  }
  if (HALO) {
    __syncthreads();
    if (tid == 0 && atomicAdd(&hd->ticket, 1u) == gridDim.x - 1) {
      hd->ticket = 0;
      pp->halo_seq = hseq;
    }
  }
}

// ---------------------------------------------------------------------------------------
// The streaming structure for 3 DOF per node (tetrahedra): node i owns rows 3i .. 3i+2, whose values are one
// contiguous run of 9 * valence doubles (row r at offset r * 3 * valence), and whose columns are the triples
// (3m, 3m+1, 3m+2) of its neighbour nodes m.  k_block_pattern3 extracts bptr[i] = rowptr[3i] / 9 and
// bidx[b] = colidx[rowptr[3i] + 3j] / 3: 4 index bytes per 72-byte block instead of 36 -- the scalar streamed kernel
// the tetrahedral PCG used before moves 12 nnz, this one 8.44 nnz.  A tile = 30 nodes (16 lanes per node, one
// 3x3 block per lane: 9 LDS.64, one index, three gathers of x, 9 FMAs); value slices are 8-byte aligned only, so
// the copies are aligned down / rounded up to 16 bytes and the last two tiles are copied with exact bounds, as in
// k_spmv_stream1.  HBM bytes per call: 8 nnz + 4 nnz / 9 + 4 n / 3 (bptr) + 8 n (x) + 8 n (y).
// ---------------------------------------------------------------------------------------
constexpr int kBlock3Tile = 2 * kStreamConsumerWarps;           // 30 nodes per tile: two 16-lane groups per warp
constexpr int kBlock3PtrInts = (kBlock3Tile + 1 + 3 + 3) & ~3;  // 36: the slice starts at the tile's first node rounded down to 4

__global__ void __launch_bounds__(256) k_block_pattern3(int32_t n_nodes, const int32_t *__restrict__ rowptr,
                                                       const int32_t *__restrict__ colidx, int32_t *__restrict__ bptr,
                                                       int32_t *__restrict__ bidx, int *__restrict__ bad) {
  const int64_t node = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 4;  // 16 lanes per node
  const int lane = threadIdx.x & 15;
  if (node > n_nodes) return;
  if (node == n_nodes) {
    if (lane == 0) bptr[n_nodes] = rowptr[3 * n_nodes] / 9;
    return;
  }
  const int32_t s0 = rowptr[3 * node], s1 = rowptr[3 * node + 1], s3 = rowptr[3 * node + 3];
  const int deg = (s3 - s0) / 9;
  // the three rows of a node must have the same length and start on a block boundary
  if (lane == 0) {
    if (s0 % 9 != 0 || (s3 - s0) != 9 * deg || (s1 - s0) != 3 * deg || rowptr[3 * node + 2] - s1 != 3 * deg) *bad = 1;
    bptr[node] = s0 / 9;
  }
  for (int j = lane; j < deg; j += 16) {
    const int32_t c = colidx[s0 + 3 * j];
    if (c % 3 != 0) *bad = 1;
    bidx[s0 / 9 + j] = c / 3;
  }
}

static size_t stream3_smem_bytes(int cap) {
  const size_t ptr_bytes = (128 + (size_t)kStreamStages * kBlock3PtrInts * sizeof(int32_t) + 127) / 128 * 128;
  return ptr_bytes + (size_t)kStreamStages * (9 * (size_t)cap + 2) * 8 + (size_t)kStreamStages * (cap + 8) * 4;
}

template <bool DOT, bool HALO>
__global__ void __launch_bounds__(kStreamThreads, 2) k_spmv_stream3(
    int32_t n_nodes, int cap /* blocks per stage */, const int32_t *__restrict__ bptr, const int32_t *__restrict__ bidx,
    const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y,
    double *__restrict__ partials, PcgState *__restrict__ st, P2PDev *pp, HaloDev *hd,
    const int32_t *__restrict__ send_idx) {
  constexpr int T = kBlock3Tile, S = kStreamStages;
  __shared__ double red[kStreamThreads / 32];
  if (DOT && (st->converged | st->breakdown)) return;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);
  uint64_t *empty = full + S;
  constexpr size_t kHdr = 128;
  int32_t *ptr_s = reinterpret_cast<int32_t *>(smem + kHdr);
  constexpr size_t ptr_bytes = (kHdr + (size_t)S * kBlock3PtrInts * sizeof(int32_t) + 127) / 128 * 128;
  const size_t vstride = 9 * (size_t)cap + 2;                                             // doubles per stage
  double *vals_s = reinterpret_cast<double *>(smem + ptr_bytes);
  int32_t *idx_s = reinterpret_cast<int32_t *>(smem + ptr_bytes + (size_t)S * vstride * 8);  // S x (cap + 8) ints

  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_tiles = (n_nodes + T - 1) / T;
  unsigned hseq = 0;
  const uint4 *gcells = nullptr;
  if (HALO) {
    hseq = pp->halo_seq + 1;
    gcells = pp->ghost[pp->rank] + (hseq & 1);
    const int pushers = min((int)gridDim.x, 32);
    if ((int)blockIdx.x < pushers)
      halo_push(pp, hd, send_idx, x, hseq, blockIdx.x * kStreamThreads + tid, pushers * kStreamThreads);
  }
  if (tid == 0) {
    for (int i = 0; i < S; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], kStreamConsumerWarps);
    }
    ptx::mbar_init_fence();
  }
  __syncthreads();

  double dot = 0.0;
  if (warp == kStreamConsumerWarps) {
    const int pl = tid & 31;
    int j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      const int stage = j % S, use = j / S;
      if (use > 0) {
        if (pl == 0) ptx::mbar_wait(&empty[stage], (uint32_t)((use - 1) & 1));
        __syncwarp();
      }
      const int32_t n0 = tile * T, n1 = min(n0 + T, n_nodes);
      const int32_t b0 = __ldg(bptr + n0), b1 = __ldg(bptr + n1);
      const int64_t e0 = 9 * (int64_t)b0, e1 = 9 * (int64_t)b1;  // value range of the tile
      const int64_t av = e0 & ~(int64_t)1;
      const int32_t ai = b0 & ~3;  // 16-byte aligned starts of the two slices
      if (tile < n_tiles - 2) {
        if (pl == 0) {
          const uint32_t nv = ((uint32_t)(e1 - av) + 1u) & ~1u, ni = ((uint32_t)(b1 - ai) + 3u) & ~3u;
          const bool any = b1 > b0;
          ptx::mbar_expect_tx(&full[stage], kBlock3PtrInts * 4u + (any ? nv * 8u + ni * 4u : 0u));
          ptx::bulk_load(ptr_s + stage * kBlock3PtrInts, bptr + (n0 & ~3), kBlock3PtrInts * 4u, &full[stage]);
          if (any) {
            ptx::bulk_load(vals_s + (size_t)stage * vstride, vals + av, nv * 8u, &full[stage]);
            ptx::bulk_load(idx_s + (size_t)stage * (cap + 8), bidx + ai, ni * 4u, &full[stage]);
          }
        }
      } else {
        for (int i = pl; i <= n1 - n0; i += 32) ptr_s[stage * kBlock3PtrInts + (n0 & 3) + i] = __ldg(bptr + n0 + i);
        for (int64_t i = pl; i < e1 - e0; i += 32) vals_s[(size_t)stage * vstride + (e0 - av) + i] = __ldg(vals + e0 + i);
        for (int i = pl; i < b1 - b0; i += 32) idx_s[(size_t)stage * (cap + 8) + (b0 - ai) + i] = __ldg(bidx + b0 + i);
        __syncwarp();
        if (pl == 0) ptx::mbar_arrive(&full[stage]);
      }
    }
  } else {
    const int grp = tid >> 4, lane = tid & 15;  // 30 groups of 16 lanes
    int j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      const int stage = j % S, use = j / S;
      ptx::mbar_wait(&full[stage], (uint32_t)(use & 1));
      const int32_t n0 = tile * T;
      const int nn = min(T, n_nodes - n0);
      const int32_t *ps = ptr_s + stage * kBlock3PtrInts + (n0 & 3);
      const int32_t b0 = ps[0];
      const double *vs = vals_s + (size_t)stage * vstride + ((9 * (int64_t)b0) & 1);
      const int32_t *is = idx_s + (size_t)stage * (cap + 8) + (b0 - (b0 & ~3));
      int32_t sb = 0, deg = 0;
      if (grp < nn) {
        sb = ps[grp] - b0;
        deg = ps[grp + 1] - ps[grp];
      }
      const double *vn = vs + 9 * (size_t)sb;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      for (int k = lane; k < deg; k += 16) {
        const int32_t c = is[sb + k];
        double x0, x1, x2;
        if (HALO && c >= n_nodes) {
          const uint4 *gc = gcells + 2 * (size_t)(3 * (c - n_nodes));
          x0 = ll_wait(gc, hseq);
          x1 = ll_wait(gc + 2, hseq);
          x2 = ll_wait(gc + 4, hseq);
        } else {
          const double *xp = x + 3 * (size_t)c;
          x0 = __ldg(xp);
          x1 = __ldg(xp + 1);
          x2 = __ldg(xp + 2);
        }
        const double *r0 = vn + 3 * k, *r1 = r0 + 3 * deg, *r2 = r1 + 3 * deg;
        a0 += r0[0] * x0 + r0[1] * x1 + r0[2] * x2;
        a1 += r1[0] * x0 + r1[1] * x1 + r1[2] * x2;
        a2 += r2[0] * x0 + r2[1] * x1 + r2[2] * x2;
      }
      __syncwarp();
      if ((tid & 31) == 0) ptx::mbar_arrive(&empty[stage]);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      }
      if (lane == 0 && grp < nn) {
        double *yp = y + 3 * (size_t)(n0 + grp);
        yp[0] = a0;
        yp[1] = a1;
        yp[2] = a2;
        if (DOT) {
          const double *xs = x + 3 * (size_t)(n0 + grp);
          dot += a0 * xs[0] + a1 * xs[1] + a2 * xs[2];
        }
      }
    }
  }
  if (DOT) {
    const double loc[1] = {dot};
    publish_and_reduce<1, kStreamThreads>(loc, partials, st, 0, red);
  }
  if (HALO) {
    __syncthreads();
    if (tid == 0 && atomicAdd(&hd->ticket, 1u) == gridDim.x - 1) {
      hd->ticket = 0;
      pp->halo_seq = hseq;
    }
  }
}

// max over the tiles of the streamed scalar kernel of (entries in the tile): sizes its ring stages
__global__ void __launch_bounds__(256) k_tile_max_entries(int32_t n_rows, int32_t tile, const int32_t *__restrict__ rowptr,
                                                         int *__restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t r0 = t * tile;
  if (r0 >= n_rows) return;
  const int64_t r1 = r0 + tile < n_rows ? r0 + tile : n_rows;
  const int cnt = rowptr[r1] - rowptr[r0];
  if (cnt > *reinterpret_cast<volatile int *>(out)) atomicMax(out, cnt);
}

static size_t stream1_smem_bytes(int cap, int passes) {
  const size_t ptr_bytes = (128 + (size_t)kStreamStages * stream1_ptr_ints(passes) * sizeof(int32_t) + 127) / 128 * 128;
  return ptr_bytes + (size_t)kStreamStages * (cap + 2) * 8 + (size_t)kStreamStages * (cap + 8) * 4;
}

struct StreamPlan {
  bool on = false;
  bool scalar = false;  // k_spmv_stream1 (1 DOF per node) instead of the 2x2-block kernel
  bool bs3 = false;     // k_spmv_stream3 (3 DOF per node)
  int passes = 4;       // k_spmv_stream1: rows per 8-lane group and tile (4, 2 or 1)
  int T = 0, cap = 0, grid = 0;
  size_t smem = 0;
  int32_t *bptr = nullptr, *bidx = nullptr;
};

static size_t stream_smem_bytes(int /*T*/, int cap) {
  const size_t ptr_bytes = (128 + (size_t)kStreamStages * kStreamPtrInts * sizeof(int32_t) + 127) / 128 * 128;
  return ptr_bytes + (size_t)kStreamStages * cap * 32 + (size_t)kStreamStages * (cap + 8) * 4;
}
