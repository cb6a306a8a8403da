// The whole Jacobi-PCG solve as ONE persistent cooperative kernel (2 DOF per node).
//
// Included by solve.cu after spmv_stream.cuh (inside namespace fe).
//
// Why: with three kernels per iteration (k_spmv_stream, k_pcg_update, k_pcg_pupdate) an 8-GPU
// iteration of 74 us of HBM time paid ~25 us for three kernel boundaries and two exposed all-reduce
// round trips.  Here an iteration has NO kernel boundary and ONE reduction:
//
//   * single-reduction CG (Chronopoulos & Gear 1989):  with u = D^-1 r and w = A u,
//       gamma = (r, u), delta = (w, u)   -- both known after the SpMV, reduced TOGETHER
//       beta = gamma / gamma_prev,  alpha = gamma / (delta - beta gamma / alpha_prev)
//       p = u + beta p;  s = w + beta s;  x += alpha p;  r -= alpha s;  u = D^-1 r
//     Same Krylov iterates as the textbook recurrence in exact arithmetic; same HBM bytes
//     (SpMV + 96 n of vector traffic against 88 n + 16 n).
//   * persistent CTAs (2 per SM, all co-resident: cooperative launch).  SpMV tiles and the 480-node
//     chunks of the vector phase are both dealt round-robin (CTA b: b, b + grid, ...), so at any
//     moment the grid works on one narrow window of the mesh and the gathers of u hit L2 (a first
//     version with one contiguous range per CTA moved 16 % more DRAM bytes: 296 windows of u do not
//     fit L2).  The tile order is rotated by half the mesh so that the rows that read ghost values
//     come up in the middle of a pass, long after the neighbours' stores have landed.
//   * per iteration:  phase A  w = A u over the CTA's tiles (the TMA ring of k_spmv_stream) + partial
//                              (delta, gamma, ||r||^2)
//                     barrier  (one arrival atomic per CTA, one polling thread per CTA, one acquire
//                              fence), then every CTA sums the per-CTA partials
//                              in the same fixed order: all CTAs (and all ranks) hold bit-identical
//                              scalars and take the same decisions without further communication
//                     phase C  the fused vector update (chunks of 480 nodes)
//                     barrier  (u complete before anybody gathers it), then every CTA stores its share
//                              of the rank's interface values of u straight into the neighbours' LL cells
//     The producer warp is released for the next SpMV as soon as the scalars are known, so its first
//     ring stages load while the consumers still run phase C and the barrier.
//   * multi-GPU: after the first barrier CTA 0 pushes the rank's three sums into every peer's LL
//     cells (common.cuh) and every CTA waits for the R cells of its own block: one cross-GPU round
//     trip per iteration.
//   * the kernel also does the initial / restart residual (r = b - A x from scratch) itself, so one
//     launch runs a whole solve; the host only verifies the true residual when the recurrence
//     reports convergence (pcg_drive) and relaunches if it has drifted.
// u is WRITTEN during the kernel by other CTAs, so it is never read through the non-coherent path
// (no __ldg / ld.global.nc): plain ld.global + the acquire fence of the barrier (MEMBAR + CCTL.IVALL).
#pragma once

constexpr int kPConsumers = kStreamConsumerWarps * 32;  // 480 threads run the numerics, 1 warp feeds TMA
constexpr int kPQ = 4;                                  // reduced quantities: delta, gamma, r.r, b.b

struct PersistArgs {
  int32_t n_nodes;
  int32_t cap;     // blocks per ring stage (StreamPlan)
  int32_t it_end;  // stop when this many iterations are done (== iters at entry: verify only)
  int32_t pad;
  const int32_t *bptr, *bidx;
  const double *vals, *b, *dinv;
  double *x, *r, *w, *u, *p, *s;
  double *partials;  // [2][kPQ] published totals, then [grid][kPQ] per-CTA partials
  unsigned *flags;   // barrier words (grid_barrier): [0] epoch carried across launches, [32] counter, [64] generation
  PcgState *st;
  P2PDev *pp;        // NULL on one GPU
  HaloDev *hd;
  const int32_t *send_idx;
};

struct PersistShared {
  double red[kStreamConsumerWarps][kPQ];
  double part[kMaxRanks][kPQ];
  double tot[kPQ];
  uint64_t go_bar;
  volatile int go_val;
  volatile int fail;
  volatile int is_last;
  unsigned long long t_mark, t_acc[5];  // phase timers (CTA 0, thread 0)
};

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned *p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// coherent (never .nc) loads of data other CTAs write during the kernel
__device__ __forceinline__ double2 ld_coherent2(const double *p) {
  double2 r;
  asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ double ld_coherent(const double *p) {
  double r;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

// Grid-wide barrier over the consumer threads of all CTAs: ONE arrival atomic per CTA on a counter, the
// last arriver resets it and publishes the epoch in a generation word that one thread per CTA polls
// (relaxed loads), followed by one acquire fence (MEMBAR + CCTL.IVALL: the SM's L1 forgets u).  A first
// version let 296 threads of every CTA poll 296 per-CTA flags: 87 K loads per round on ten cache
// lines of one L2 slice -- 15 us per barrier, measured.  Bounded: after ~10 s without release the
// poller raises st->breakdown, which every later wait honours at once.
//   flags[0] epoch carried from launch to launch | flags[32] arrival counter | flags[64] generation
__device__ __forceinline__ void grid_barrier(unsigned *flags, unsigned ep, int grid, PcgState *st, int ctid,
                                             PersistShared &sh) {
  ptx::named_barrier(1, kPConsumers);
  if (ctid == 0) {
    __threadfence();  // the CTA's writes (ordered before by the bar.sync) are visible before the arrival
    unsigned old;
    asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(flags + 32) : "memory");
    if (old == (unsigned)grid - 1u) {
      st_relaxed_u32(flags + 32, 0u);
      __threadfence();
      st_relaxed_u32(flags + 64, ep);
    } else {
      long long n = 0;
      while ((int)(ld_relaxed_u32(flags + 64) - ep) < 0) {
        if (sh.fail) break;
        if (++n > 2000) {
          __nanosleep(64);
          if ((n & 1023) == 0 && *reinterpret_cast<volatile int *>(&st->breakdown)) sh.fail = 1;
          if (n > (1ll << 27)) {
            sh.fail = 1;
            st->breakdown = 2;
          }
        }
      }
    }
    __threadfence();
  }
  ptx::named_barrier(1, kPConsumers);
}

// v[0..kPQ) summed over the 480 consumer threads in a fixed order -> sh.tot (valid for every thread
// after the call).
__device__ __forceinline__ void block_reduce_q(double (&v)[kPQ], PersistShared &sh, int ctid) {
#pragma unroll
  for (int q = 0; q < kPQ; ++q) v[q] = warp_sum(v[q]);
  if ((ctid & 31) == 0) {
#pragma unroll
    for (int q = 0; q < kPQ; ++q) sh.red[ctid >> 5][q] = v[q];
  }
  ptx::named_barrier(1, kPConsumers);
  if (ctid < kPQ) {
    double t = 0.0;
#pragma unroll
    for (int wv = 0; wv < kStreamConsumerWarps; ++wv) t += sh.red[wv][ctid];
    sh.tot[ctid] = t;
  }
  ptx::named_barrier(1, kPConsumers);
}

// MULTI: peer-memory transport (ghost columns read from LL cells, sums all-reduced through LL cells).
template <bool MULTI>
__global__ void __launch_bounds__(kStreamThreads, 2) k_pcg_persist(const PersistArgs a) {
  constexpr int T = kStreamTile, S = kStreamStages, GROUPS = kStreamGroups, PASSES = 2;
  __shared__ PersistShared sh;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);
  uint64_t *empty = full + S;
  constexpr size_t kHdr = 128;
  int32_t *ptr_s = reinterpret_cast<int32_t *>(smem + kHdr);
  constexpr size_t ptr_bytes = (kHdr + (size_t)S * kStreamPtrInts * sizeof(int32_t) + 127) / 128 * 128;
  const int cap = a.cap;
  double *vals_s = reinterpret_cast<double *>(smem + ptr_bytes);
  int32_t *idx_s = reinterpret_cast<int32_t *>(smem + ptr_bytes + (size_t)S * cap * 32);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int grid = gridDim.x;
  const int32_t n_nodes = a.n_nodes;
  const int n_tiles = (n_nodes + T - 1) / T;
  const int rot = MULTI ? n_tiles / 2 : 0;  // multi-GPU: ghost-reading tiles in the middle of the pass
  const int n_chunks = (n_nodes + kPConsumers - 1) / kPConsumers;
  if (tid == 0) {
    for (int i = 0; i < S; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], kStreamConsumerWarps);
    }
    ptx::mbar_init(&sh.go_bar, 1);
    sh.go_val = 0;
    sh.fail = 0;
    ptx::mbar_init_fence();
  }
  __syncthreads();

  if (warp == kStreamConsumerWarps) {
    // ===== producer warp: one lane feeds the ring; a pass (= one SpMV) starts when the consumers say so
    if ((tid & 31) == 0) {
      int j = 0;
      for (unsigned pass = 0;; ++pass) {
        ptx::mbar_wait(&sh.go_bar, pass & 1);
        if (!sh.go_val) break;
        for (int tt = blockIdx.x; tt < n_tiles; tt += grid, ++j) {
          const int tile = tt + rot < n_tiles ? tt + rot : tt + rot - n_tiles;
          const int stage = j % S, use = j / S;
          if (use > 0) ptx::mbar_wait(&empty[stage], (uint32_t)((use - 1) & 1));
          const int32_t n0 = tile * T, n1 = min(n0 + T, n_nodes);
          const int32_t b0 = __ldg(a.bptr + n0), b1 = __ldg(a.bptr + n1);
          const int32_t a0 = b0 & ~3;
          const uint32_t nb = (uint32_t)(b1 - b0);
          const uint32_t ni = ((uint32_t)(b1 - a0) + 3u) & ~3u;
          ptx::mbar_expect_tx(&full[stage], kStreamPtrInts * 4u + nb * 32u + (nb ? ni * 4u : 0u));
          ptx::bulk_load(ptr_s + stage * kStreamPtrInts, a.bptr + n0, kStreamPtrInts * 4u, &full[stage]);
          if (nb) {
            ptx::bulk_load(vals_s + (size_t)stage * cap * 4, a.vals + 4 * (int64_t)b0, nb * 32u, &full[stage]);
            ptx::bulk_load(idx_s + (size_t)stage * (cap + 8), a.bidx + a0, ni * 4u, &full[stage]);
          }
        }
      }
    }
    return;
  }

  // ===== consumer warps =====
  const int ctid = tid;  // consumers are warps 0 .. 14
  const int grp = tid >> 3, lane = tid & 7;
  PcgState *st = a.st;
  P2PDev *pp = MULTI ? a.pp : nullptr;
  const HaloDev *hd = a.hd;
  const int R = pp ? pp->nranks : 1, me = pp ? pp->rank : 0;
  unsigned ep = a.flags[0];
  unsigned rseq = pp ? pp->red_seq : 0u, hseq = pp ? pp->halo_seq : 0u;
  const double tol2 = st->tol2;
  int it = st->iters;
  const int breakdown_in = st->breakdown;
  int jc = 0;  // ring position of the consumers (same sequence as the producer's j)
  double2 *x2 = reinterpret_cast<double2 *>(a.x), *r2 = reinterpret_cast<double2 *>(a.r);
  double2 *w2 = reinterpret_cast<double2 *>(a.w), *u2 = reinterpret_cast<double2 *>(a.u);
  double2 *p2 = reinterpret_cast<double2 *>(a.p), *s2 = reinterpret_cast<double2 *>(a.s);
  const double2 *b2 = reinterpret_cast<const double2 *>(a.b), *d2 = reinterpret_cast<const double2 *>(a.dinv);

  // phase timers of CTA 0 (globaltimer, ns): where an iteration's time goes, reported through st->prof
  const bool prof_on = blockIdx.x == 0 && ctid == 0;
  auto now = [&]() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
  };
  auto lap = [&](int slot) {
    if (prof_on) {
      const unsigned long long t = now();
      sh.t_acc[slot] += t - sh.t_mark;
      sh.t_mark = t;
    }
  };
  if (prof_on) {
    for (int q = 0; q < 5; ++q) sh.t_acc[q] = 0;
    sh.t_mark = now();
  }
  auto give_go = [&](int v) {
    if (ctid == 0) {
      sh.go_val = v;
      ptx::mbar_arrive(&sh.go_bar);
    }
  };
  // u is complete (grid barrier): this CTA's share of the rank's interface values -> the neighbours'
  // ghost cells, exchange number `seq` (fire-and-forget peer stores; the readers wait per cell)
  auto barrier_and_push = [&](unsigned seq) {
    grid_barrier(a.flags, ++ep, grid, st, ctid, sh);
    if (!pp) return;
    const int n_send = hd->n_send;
    for (int i = blockIdx.x * kPConsumers + ctid; i < n_send; i += grid * kPConsumers) {
      int k = 0;
      while (i >= hd->send_ptr[k + 1]) ++k;
      ll_store(pp->ghost[hd->nbr_rank[k]] + 2 * (size_t)(hd->dst_off[k] + (i - hd->send_ptr[k])) + (seq & 1),
               ld_coherent(a.u + __ldg(a.send_idx + i)), seq);
    }
  };
  // w = A u over the CTA's tiles; returns this thread's share of (w, u)
  auto spmv_pass = [&]() -> double {
    double dot = 0.0;
    const uint4 *gcells = pp ? pp->ghost[me] + (hseq & 1) : nullptr;
    for (int tt = blockIdx.x; tt < n_tiles; tt += grid, ++jc) {
      const int tile = tt + rot < n_tiles ? tt + rot : tt + rot - n_tiles;
      const int stage = jc % S, use = jc / S;
      ptx::mbar_wait(&full[stage], (uint32_t)(use & 1));
      const int32_t n0 = tile * T;
      const int nn = min(T, n_nodes - n0);
      const int32_t *ps = ptr_s + stage * kStreamPtrInts;
      const int32_t b0 = ps[0];
      const double2 *vs = reinterpret_cast<const double2 *>(vals_s + (size_t)stage * cap * 4);
      const int32_t *is = idx_s + (size_t)stage * (cap + 8) + (b0 - (b0 & ~3));
      int32_t sft[PASSES], deg[PASSES], c[PASSES];
      double2 v0[PASSES], v1[PASSES], xv[PASSES];
#pragma unroll
      for (int q = 0; q < PASSES; ++q) {
        const int i = grp + q * GROUPS;
        sft[q] = 0;
        deg[q] = 0;
        if (i < nn) {
          sft[q] = ps[i] - b0;
          deg[q] = ps[i + 1] - ps[i];
        }
        const bool act = lane < deg[q];
        c[q] = act ? is[sft[q] + lane] : n0;  // (inactive lanes gather an own, always valid entry)
        v0[q] = act ? vs[2 * sft[q] + lane] : make_double2(0.0, 0.0);
        v1[q] = act ? vs[2 * sft[q] + deg[q] + lane] : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int q = 0; q < PASSES; ++q)  // every gather of u in flight (a ghost column reads an own entry here ...)
        xv[q] = ld_coherent2(a.u + 2 * (size_t)((!MULTI || c[q] < n_nodes) ? c[q] : n0));
      if (MULTI) {                       // ... and takes its value pair from the LL cells)
#pragma unroll
        for (int q = 0; q < PASSES; ++q)
          if (c[q] >= n_nodes) xv[q] = ghost_pair(gcells, c[q] - n_nodes, hseq);
      }
      double a0[PASSES], a1[PASSES];
#pragma unroll
      for (int q = 0; q < PASSES; ++q) {
        a0[q] = v0[q].x * xv[q].x + v0[q].y * xv[q].y;
        a1[q] = v1[q].x * xv[q].x + v1[q].y * xv[q].y;
        for (int k = lane + 8; k < deg[q]; k += 8) {  // valence > 8
          const int32_t ck = is[sft[q] + k];
          const double2 xx = (!MULTI || ck < n_nodes) ? ld_coherent2(a.u + 2 * (size_t)ck) : ghost_pair(gcells, ck - n_nodes, hseq);
          const double2 k0 = vs[2 * sft[q] + k], k1 = vs[2 * sft[q] + deg[q] + k];
          a0[q] += k0.x * xx.x + k0.y * xx.y;
          a1[q] += k1.x * xx.x + k1.y * xx.y;
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) ptx::mbar_arrive(&empty[stage]);
#pragma unroll
      for (int q = 0; q < PASSES; ++q) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          a0[q] += __shfl_xor_sync(0xffffffffu, a0[q], o);
          a1[q] += __shfl_xor_sync(0xffffffffu, a1[q], o);
        }
        const int i = grp + q * GROUPS;
        if (lane == 0 && i < nn) {
          w2[n0 + i] = make_double2(a0[q], a1[q]);
          const double2 us = ld_coherent2(a.u + 2 * (size_t)(n0 + i));
          dot += a0[q] * us.x + a1[q] * us.y;
        }
      }
    }
    return dot;
  };

  // ---- (re)start from x:  u <- x,  w = A u,  r = b - w,  u = D^-1 r;  partials of (r,u), (r,r), (b,b)
  // (a grid barrier on both sides of the SpMV: u is gathered across CTAs)
  for (int ch = blockIdx.x; ch < n_chunks; ch += grid) {
    const int32_t i = ch * kPConsumers + ctid;
    if (i < n_nodes) u2[i] = x2[i];
  }
  barrier_and_push(++hseq);
  give_go(1);
  (void)spmv_pass();
  grid_barrier(a.flags, ++ep, grid, st, ctid, sh);
  double acc_g = 0.0, acc_rr = 0.0, acc_bb = 0.0;
  for (int ch = blockIdx.x; ch < n_chunks; ch += grid) {
    const int32_t i = ch * kPConsumers + ctid;
    if (i < n_nodes) {
      const double2 bi = b2[i], wi = w2[i], di = __ldg(d2 + i);
      const double2 ri = make_double2(bi.x - wi.x, bi.y - wi.y);
      const double2 ui = make_double2(di.x * ri.x, di.y * ri.y);
      r2[i] = ri;
      u2[i] = ui;
      acc_g += ri.x * ui.x + ri.y * ui.y;
      acc_rr += ri.x * ri.x + ri.y * ri.y;
      acc_bb += bi.x * bi.x + bi.y * bi.y;
    }
  }
  // a diagonal that is not positive (k_extract_dinv) must stop EVERY rank: poison the reduction
  if (blockIdx.x == 0 && ctid == 0 && breakdown_in) acc_bb = __longlong_as_double(0x7ff8000000000000ll);
  barrier_and_push(++hseq);
  give_go(1);
  lap(4);

  bool first = true;
  double alpha_prev = 0.0, gamma_prev = 0.0, bb = 0.0;
  while (true) {
    // ---- phase A
    const double dot = spmv_pass();
    lap(0);
    double v[kPQ] = {dot, acc_g, acc_rr, acc_bb};
    block_reduce_q(v, sh, ctid);
    double *cta_part = a.partials + 2 * kPQ, *totals = a.partials + (rseq & 1) * kPQ;  // [2][kPQ] | [grid][kPQ]
    if (ctid < kPQ) cta_part[(size_t)blockIdx.x * kPQ + ctid] = sh.tot[ctid];
    // ---- reduction = barrier.  Every CTA arrives on the counter; the LAST one adds the per-CTA partials in
    //      a fixed order and publishes the rank's sums: into every peer's LL cells (MULTI), or into the
    //      totals slot of this launch parity followed by the generation word.
    ++ep;
    ++rseq;
    ptx::named_barrier(1, kPConsumers);
    if (ctid == 0) {
      __threadfence();
      unsigned old;
      asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(a.flags + 32) : "memory");
      const int last = old == (unsigned)grid - 1u;
      if (last) {
        st_relaxed_u32(a.flags + 32, 0u);
        __threadfence();  // acquire: the other CTAs' partials
      }
      sh.is_last = last;
    }
    ptx::named_barrier(1, kPConsumers);
    if (sh.is_last) {
      double acc[kPQ] = {0.0, 0.0, 0.0, 0.0};
      for (int cta = ctid; cta < grid; cta += kPConsumers) {
        const double2 lo = ld_coherent2(cta_part + (size_t)cta * kPQ), hi = ld_coherent2(cta_part + (size_t)cta * kPQ + 2);
        acc[0] += lo.x;
        acc[1] += lo.y;
        acc[2] += hi.x;
        acc[3] += hi.y;
      }
      block_reduce_q(acc, sh, ctid);
      if (MULTI) {
        if (ctid < R) {
          uint4 *dst = pp->red[ctid] + ((size_t)(rseq & 1) * R + me) * 4;
#pragma unroll
          for (int q = 0; q < kPQ; ++q) ll_store(dst + q, sh.tot[q], rseq);
        }
      } else {
        if (ctid < kPQ) totals[ctid] = sh.tot[ctid];
        ptx::named_barrier(1, kPConsumers);
        if (ctid == 0) {
          __threadfence();
          st_relaxed_u32(a.flags + 64, ep);
        }
      }
    }
    if (MULTI) {
      // the deferred  x += alpha_prev p  of the previous iteration runs while the sums cross NVLink
      if (!first) {
        for (int ch = blockIdx.x; ch < n_chunks; ch += grid) {
          const int32_t i = ch * kPConsumers + ctid;
          if (i < n_nodes) {
            const double2 xi = x2[i], pi = p2[i];
            x2[i] = make_double2(xi.x + alpha_prev * pi.x, xi.y + alpha_prev * pi.y);
          }
        }
      }
      if (ctid < R) {
        const uint4 *src = pp->red[me] + ((size_t)(rseq & 1) * R + ctid) * 4;
#pragma unroll
        for (int q = 0; q < kPQ; ++q) sh.part[ctid][q] = ll_wait(src + q, rseq);
        __threadfence();  // acquire (+ CCTL.IVALL): w of the other CTAs, ordered before the sums by their fences
      }
      ptx::named_barrier(1, kPConsumers);
      if (ctid < kPQ) {
        double t = 0.0;
        for (int rk = 0; rk < R; ++rk) t += sh.part[rk][ctid];
        sh.tot[ctid] = t;
      }
      ptx::named_barrier(1, kPConsumers);
    } else {
      if (ctid == 0) {
        long long n = 0;
        while ((int)(ld_relaxed_u32(a.flags + 64) - ep) < 0) {
          if (sh.fail) break;
          if (++n > 2000) {
            __nanosleep(64);
            if ((n & 1023) == 0 && *reinterpret_cast<volatile int *>(&st->breakdown)) sh.fail = 1;
            if (n > (1ll << 27)) {
              sh.fail = 1;
              st->breakdown = 2;
            }
          }
        }
        __threadfence();
      }
      ptx::named_barrier(1, kPConsumers);
      if (ctid < kPQ) sh.tot[ctid] = ld_coherent(totals + ctid);
      ptx::named_barrier(1, kPConsumers);
    }
    const double delta = sh.tot[0], gamma = sh.tot[1], rr = sh.tot[2];
    if (first) bb = sh.tot[3];
    lap(1);
    // ---- decisions, identical in every CTA of every rank
    const bool conv = !(bb > 0.0) ? (bb == 0.0) : (rr <= tol2 * bb);  // NaN b.b is a breakdown, not convergence
    const double beta = first ? 0.0 : gamma / gamma_prev;
    const double denom = first ? delta : delta - beta * gamma / alpha_prev;  // = (p, A p)
    const double alpha = gamma / denom;
    int bad = 0;
    if (!conv && (!(denom > 0.0) || !isfinite(denom) || !isfinite(gamma) || !(bb == bb))) bad = 1;
    if (sh.fail || *reinterpret_cast<volatile int *>(&st->breakdown) == 2) bad = 2;
    if (conv || bad || it >= a.it_end) {
      if (blockIdx.x == 0 && ctid == 0) {
        st->sums[0] = first ? delta : denom;
        st->sums[1] = gamma;
        st->sums[2] = rr;
        st->sums[3] = bb;
        st->iters = it;
        st->converged = (conv && !bad) ? (first ? 2 : 1) : 0;
        if (bad) st->breakdown = bad;
        a.flags[0] = ep;
        for (int q = 0; q < 5; ++q) st->prof[q] = (double)sh.t_acc[q];
        if (pp) {
          pp->red_seq = rseq;
          pp->halo_seq = hseq;
        }
      }
      give_go(0);
      break;
    }
    give_go(1);  // the ring refills for the next SpMV while phase C runs
    // ---- phase C: fused update, chunks of 480 nodes dealt round-robin (w is complete: barrier above)
    acc_g = acc_rr = acc_bb = 0.0;
    for (int ch = blockIdx.x; ch < n_chunks; ch += grid) {
      const int32_t i = ch * kPConsumers + ctid;
      if (i >= n_nodes) continue;
      const double2 ui = u2[i], wi = w2[i], ri = r2[i], di = __ldg(d2 + i);
      double2 pn = ui, sn = wi;
      if (!first) {
        const double2 pi = p2[i], si = s2[i];
        pn = make_double2(ui.x + beta * pi.x, ui.y + beta * pi.y);
        sn = make_double2(wi.x + beta * si.x, wi.y + beta * si.y);
      }
      const double2 rn = make_double2(ri.x - alpha * sn.x, ri.y - alpha * sn.y);
      const double2 un = make_double2(di.x * rn.x, di.y * rn.y);
      p2[i] = pn;
      s2[i] = sn;
      if (!MULTI) {  // (MULTI: deferred to the next reduction, where it hides the cross-GPU round trip)
        const double2 xi = x2[i];
        x2[i] = make_double2(xi.x + alpha * pn.x, xi.y + alpha * pn.y);
      }
      r2[i] = rn;
      u2[i] = un;
      acc_g += rn.x * un.x + rn.y * un.y;
      acc_rr += rn.x * rn.x + rn.y * rn.y;
    }
    lap(2);
    barrier_and_push(++hseq);
    lap(3);
    first = false;
    alpha_prev = alpha;
    gamma_prev = gamma;
    ++it;
  }
}
