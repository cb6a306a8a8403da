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
//   * persistent CTAs (2 per SM, all co-resident: cooperative launch).  CTA b owns a CONTIGUOUS
//     range of SpMV tiles and the vector entries of exactly those rows, so w, r, p, s, x never cross
//     a CTA; only u (gathered by the neighbours' rows) and the three scalars do.
//   * per iteration:  phase A  w = A u over the CTA's tiles (the TMA ring of k_spmv_stream) + partial
//                              (delta, gamma, ||r||^2)
//                     barrier  flag array (one release store per CTA, one relaxed poll per flag, one
//                              acquire fence) -- no atomics -- then every CTA sums the per-CTA partials
//                              in the same fixed order: all CTAs (and all ranks) hold bit-identical
//                              scalars and take the same decisions without further communication
//                     phase C  the fused vector update on the CTA's own rows, interface values of u
//                              pushed straight into the neighbours' LL cells by the CTA that owns them
//                     barrier  (u complete before anybody gathers it)
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
  double *partials;  // [grid][kPQ]
  unsigned *flags;   // [0] = barrier epoch carried from launch to launch, [1 .. grid] = per-CTA flags
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
  int own_lo[kMaxRanks], own_hi[kMaxRanks];
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

// Grid-wide barrier over the consumer threads of all CTAs.  Every CTA publishes the epoch in its own
// flag (after a fence that covers the whole CTA's writes: bar.sync + cumulativity) and polls the
// others' flags with relaxed loads; one acquire fence at the end invalidates L1.  Bounded: a CTA that
// never arrives (cannot happen with a cooperative launch unless a peer GPU died inside a wait) makes
// the pollers give up after ~10 s and raise st->breakdown, which every later wait honours at once.
__device__ __forceinline__ void grid_barrier(unsigned *flags, unsigned ep, int grid, PcgState *st, int ctid,
                                             PersistShared &sh) {
  ptx::named_barrier(1, kPConsumers);
  if (ctid == 0) {
    __threadfence();
    st_relaxed_u32(flags + 1 + blockIdx.x, ep);
  }
  for (int t = ctid; t < grid; t += kPConsumers) {
    long long n = 0;
    while ((int)(ld_relaxed_u32(flags + 1 + t) - ep) < 0) {
      if (sh.fail) break;
      if (++n > 4000) {
        __nanosleep(128);
        if ((n & 1023) == 0 && *reinterpret_cast<volatile int *>(&st->breakdown)) sh.fail = 1;
        if (n > (1ll << 26)) {
          sh.fail = 1;
          st->breakdown = 2;
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
  const int t0 = (int)((long long)blockIdx.x * n_tiles / grid);
  const int t1 = (int)((long long)(blockIdx.x + 1) * n_tiles / grid);
  const int32_t node0 = min(t0 * T, n_nodes), node1 = min(t1 * T, n_nodes);
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
        for (int tile = t0; tile < t1; ++tile, ++j) {
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
  P2PDev *pp = a.pp;
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

  // this CTA's share of every neighbour's send list (ascending local DOFs per neighbour: checked by the host)
  if (pp && ctid < hd->n_nbr) {
    const int lo = hd->send_ptr[ctid], hi = hd->send_ptr[ctid + 1];
    auto lower = [&](int32_t v) {
      int l = lo, h = hi;
      while (l < h) {
        const int m = (l + h) >> 1;
        if (__ldg(a.send_idx + m) < v) l = m + 1; else h = m;
      }
      return l;
    };
    sh.own_lo[ctid] = lower(2 * node0);
    sh.own_hi[ctid] = lower(2 * node1);
  }

  auto give_go = [&](int v) {
    if (ctid == 0) {
      sh.go_val = v;
      ptx::mbar_arrive(&sh.go_bar);
    }
  };
  // interface values of u owned by this CTA -> the neighbours' ghost cells, exchange number `seq`
  auto push_halo = [&](unsigned seq) {
    ptx::named_barrier(1, kPConsumers);  // the CTA's rows of u are complete
    if (!pp) return;
    const int n_nbr = hd->n_nbr;
    for (int k = 0; k < n_nbr; ++k) {
      uint4 *dst = pp->ghost[hd->nbr_rank[k]] + (seq & 1);
      const int base = hd->send_ptr[k], off = hd->dst_off[k];
      for (int i = sh.own_lo[k] + ctid; i < sh.own_hi[k]; i += kPConsumers)
        ll_store(dst + 2 * (size_t)(off + (i - base)), ld_coherent(a.u + __ldg(a.send_idx + i)), seq);
    }
  };
  // w = A u over the CTA's tiles; returns this thread's share of (w, u)
  auto spmv_pass = [&]() -> double {
    double dot = 0.0;
    const uint4 *gcells = pp ? pp->ghost[me] + (hseq & 1) : nullptr;
    for (int tile = t0; tile < t1; ++tile, ++jc) {
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
      for (int q = 0; q < PASSES; ++q)
        xv[q] = (c[q] < n_nodes) ? ld_coherent2(a.u + 2 * (size_t)c[q]) : ghost_pair(gcells, c[q] - n_nodes, hseq);
      double a0[PASSES], a1[PASSES];
#pragma unroll
      for (int q = 0; q < PASSES; ++q) {
        a0[q] = v0[q].x * xv[q].x + v0[q].y * xv[q].y;
        a1[q] = v1[q].x * xv[q].x + v1[q].y * xv[q].y;
        for (int k = lane + 8; k < deg[q]; k += 8) {  // valence > 8
          const int32_t ck = is[sft[q] + k];
          const double2 xx = (ck < n_nodes) ? ld_coherent2(a.u + 2 * (size_t)ck) : ghost_pair(gcells, ck - n_nodes, hseq);
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
  for (int32_t i = node0 + ctid; i < node1; i += kPConsumers) u2[i] = x2[i];
  push_halo(++hseq);
  grid_barrier(a.flags, ++ep, grid, st, ctid, sh);
  give_go(1);
  (void)spmv_pass();
  ptx::named_barrier(1, kPConsumers);  // the CTA's rows of w are complete
  double acc_g = 0.0, acc_rr = 0.0, acc_bb = 0.0;
  for (int32_t i = node0 + ctid; i < node1; i += kPConsumers) {
    const double2 bi = b2[i], wi = w2[i], di = __ldg(d2 + i);
    const double2 ri = make_double2(bi.x - wi.x, bi.y - wi.y);
    const double2 ui = make_double2(di.x * ri.x, di.y * ri.y);
    r2[i] = ri;
    u2[i] = ui;
    acc_g += ri.x * ui.x + ri.y * ui.y;
    acc_rr += ri.x * ri.x + ri.y * ri.y;
    acc_bb += bi.x * bi.x + bi.y * bi.y;
  }
  // a diagonal that is not positive (k_extract_dinv) must stop EVERY rank: poison the reduction
  if (blockIdx.x == 0 && ctid == 0 && breakdown_in) acc_bb = __longlong_as_double(0x7ff8000000000000ll);
  push_halo(++hseq);
  grid_barrier(a.flags, ++ep, grid, st, ctid, sh);
  give_go(1);

  bool first = true;
  double alpha_prev = 0.0, gamma_prev = 0.0, bb = 0.0;
  while (true) {
    // ---- phase A
    const double dot = spmv_pass();
    double v[kPQ] = {dot, acc_g, acc_rr, acc_bb};
    block_reduce_q(v, sh, ctid);
    if (ctid < kPQ) a.partials[(size_t)blockIdx.x * kPQ + ctid] = sh.tot[ctid];
    grid_barrier(a.flags, ++ep, grid, st, ctid, sh);
    // ---- the rank's sums: every CTA adds the per-CTA partials in the same order
    ++rseq;
    if (!pp || blockIdx.x == 0) {
      double acc[kPQ] = {0.0, 0.0, 0.0, 0.0};
      for (int cta = ctid; cta < grid; cta += kPConsumers) {
        const double2 lo = ld_coherent2(a.partials + (size_t)cta * kPQ), hi = ld_coherent2(a.partials + (size_t)cta * kPQ + 2);
        acc[0] += lo.x;
        acc[1] += lo.y;
        acc[2] += hi.x;
        acc[3] += hi.y;
      }
      block_reduce_q(acc, sh, ctid);
    }
    if (pp) {  // all-reduce over the ranks through LL cells: CTA 0 pushes, every CTA gathers its own block
      const int par = (int)(rseq & 1);
      if (blockIdx.x == 0 && ctid < R) {
        uint4 *dst = pp->red[ctid] + ((size_t)par * R + me) * 4;
#pragma unroll
        for (int q = 0; q < kPQ; ++q) ll_store(dst + q, sh.tot[q], rseq);
      }
      if (ctid < R) {
        const uint4 *src = pp->red[me] + ((size_t)par * R + ctid) * 4;
#pragma unroll
        for (int q = 0; q < kPQ; ++q) sh.part[ctid][q] = ll_wait(src + q, rseq);
      }
      ptx::named_barrier(1, kPConsumers);
      if (ctid < kPQ) {
        double t = 0.0;
        for (int rk = 0; rk < R; ++rk) t += sh.part[rk][ctid];
        sh.tot[ctid] = t;
      }
      ptx::named_barrier(1, kPConsumers);
    }
    const double delta = sh.tot[0], gamma = sh.tot[1], rr = sh.tot[2];
    if (first) bb = sh.tot[3];
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
        if (pp) {
          pp->red_seq = rseq;
          pp->halo_seq = hseq;
        }
      }
      give_go(0);
      break;
    }
    give_go(1);  // the ring refills for the next SpMV while phase C runs
    // ---- phase C: fused update on the CTA's own rows
    acc_g = acc_rr = acc_bb = 0.0;
    if (first) {
      for (int32_t i = node0 + ctid; i < node1; i += kPConsumers) {
        const double2 ui = u2[i], wi = w2[i], ri = r2[i], xi = x2[i], di = __ldg(d2 + i);
        const double2 rn = make_double2(ri.x - alpha * wi.x, ri.y - alpha * wi.y);
        const double2 un = make_double2(di.x * rn.x, di.y * rn.y);
        p2[i] = ui;
        s2[i] = wi;
        x2[i] = make_double2(xi.x + alpha * ui.x, xi.y + alpha * ui.y);
        r2[i] = rn;
        u2[i] = un;
        acc_g += rn.x * un.x + rn.y * un.y;
        acc_rr += rn.x * rn.x + rn.y * rn.y;
      }
    } else {
      for (int32_t i = node0 + ctid; i < node1; i += kPConsumers) {
        const double2 ui = u2[i], wi = w2[i], ri = r2[i], xi = x2[i], pi = p2[i], si = s2[i], di = __ldg(d2 + i);
        const double2 pn = make_double2(ui.x + beta * pi.x, ui.y + beta * pi.y);
        const double2 sn = make_double2(wi.x + beta * si.x, wi.y + beta * si.y);
        const double2 rn = make_double2(ri.x - alpha * sn.x, ri.y - alpha * sn.y);
        const double2 un = make_double2(di.x * rn.x, di.y * rn.y);
        p2[i] = pn;
        s2[i] = sn;
        x2[i] = make_double2(xi.x + alpha * pn.x, xi.y + alpha * pn.y);
        r2[i] = rn;
        u2[i] = un;
        acc_g += rn.x * un.x + rn.y * un.y;
        acc_rr += rn.x * rn.x + rn.y * rn.y;
      }
    }
    push_halo(++hseq);
    grid_barrier(a.flags, ++ep, grid, st, ctid, sh);
    first = false;
    alpha_prev = alpha;
    gamma_prev = gamma;
    ++it;
  }
}
