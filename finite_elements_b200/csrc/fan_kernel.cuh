// k_assemble_fan: the default numeric-assembly kernel (variant 3).  Included by assemble.cu inside namespace fe,
// after FanOps (per-kind arithmetic of one fan step) and FanRec (the two record formats).
//
// Every WARP of the persistent grid is an independent software pipeline over 32-node chunks (chunk = global
// warp id, + total warps, ...); warps never synchronise with each other.
//  * Input ring per warp (2 stages, one mbarrier each): while chunk c is computed, lane 0 has already handed
//    chunk c+1 to the TMA engine -- the pointer slices (adj_ptr, fan_ptr, fan_hdr: 36 words each) and the chunk's
//    contiguous record range; the end points of the record range a refill needs are read by lane 0 at the top of
//    the trip and consumed after the walk (scalar instance: by cp.async a chunk earlier; kEpLdg below).
//  * Walk: one thread per node.  A node's first record is the seed of its fan (its first neighbour); every further
//    record is one element (self, previous neighbour, this neighbour).  A step gathers ONE 16-byte coordinate,
//    adds the element's share to the block towards the previous neighbour -- which arrives in registers from the
//    step before and is stored finished -- and keeps the share of the block towards this neighbour for the next
//    step.  The diagonal block is never evaluated: it is minus the sum of the row's finished blocks (FanOps).
//    Nodes with a single fan (all but bow-ties) run two steps per trip with no flag tests; the coordinates of the
//    next trip are in flight in a second pair of registers while this trip computes.
//  * Output: the warp's private sub-tile is the exact image of its 32 nodes' slice of `vals`; one thread hands it
//    to the TMA engine as a single bulk store.
// 10 KB of shared memory per warp at valence 7 -> 5 CTAs of 4 warps per SM, 96 registers.
//
// Measured alternatives (S16M plane stress, B200, profiles/r02_b_fan_kernel_search.md; round 1: 0.458 ms):
//   this kernel, 4-byte records 0.406 ms; 8-byte records 0.410 ms; coalesced 128-bit copy-out instead of the bulk
//   store 0.47 ms; TMA L2 prefetch of the coordinates a chunk touches first 0.49 ms; neighbour coordinates through cp.async into shared memory a chunk ahead (no register
//   gathers at all, 20 KB per warp, 10 warps per SM) 0.51 - 0.57 ms; 3-stage ring + first gathers a chunk ahead
//   (13 KB per warp, 16 warps) 0.57 ms.  The kernel is bound by issue slots and dependent-instruction latency at
//   20 warps per SM (ncu: 43 % issue-active, FP64 pipe 28 %, DRAM 2.29 GB in 0.41 ms), not by HBM: every design
//   that traded warps for deeper prefetch lost.
#pragma once
constexpr int kFanThreads = 128;
constexpr int kFanWarps = 4;

__host__ __device__ inline size_t fan_stage_bytes(int rec_cap, bool r4) {
  return ((size_t)(r4 ? 3 : 2) * kFanPtrInts * sizeof(int32_t) + (size_t)rec_cap * (r4 ? 4 : 8) + 15) / 16 * 16;
}
__host__ __device__ inline size_t fan_warp_bytes(int rec_cap, int warp_slot_bytes, bool r4) {
  return (32 + 2 * fan_stage_bytes(rec_cap, r4) + (size_t)warp_slot_bytes + 127) / 128 * 128;
}

template <int KC, bool R4>
__global__ void __launch_bounds__(kFanThreads, (KC == 2 ? 7 : 5)) k_assemble_fan(
    int32_t n_owned, const int32_t *__restrict__ fan_ptr, const typename FanRec<R4>::T *__restrict__ fan_rec,
    const uint32_t *__restrict__ fan_hdr, const int32_t *__restrict__ adj_ptr, const double2 *__restrict__ coords,
    const MatRow *__restrict__ tab, double *__restrict__ vals, int rec_cap, int warp_slot_bytes) {
  using Ops = FanOps<KC>;
  using Val = typename Ops::Val;
  using Slot = typename Ops::Slot;
  using RO = FanRec<R4>;
  using Rec = typename RO::T;
  constexpr int SPB = (KC == 2) ? 1 : 2;  // Slots per node-level block
  constexpr int kPtrSlices = R4 ? 3 : 2;  // adj_ptr, fan_ptr (, fan_hdr)
  // Two scheduling choices, settled per instance by measurement (r02 captures V, W; ms at S16M, 4-byte records):
  //   kEpLdg : lane 0 reads the refill's end points with plain loads at the top of a trip (consumed after the walk)
  //            instead of a cp.async group one chunk earlier.        2 DOF: 0.417 -> 0.407; scalar: 0.243 -> 0.257
  //   kEarly : the next chunk's first loads go out before this chunk's store / refill sequence instead of after it.
  //            2 DOF (with kEpLdg): 0.407 -> 0.439 (12 B of spills); scalar (cp.async end points): 0.243 -> 0.231
  //            2 DOF, 8-byte records (what a rank of a partition runs; no spills there): 0.418 -> 0.410
  constexpr bool kEpLdg = KC != 2, kEarly = KC == 2 || !R4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *wbase = smem_raw + (size_t)warp * fan_warp_bytes(rec_cap, warp_slot_bytes, R4);
  uint64_t *full = reinterpret_cast<uint64_t *>(wbase);
  const size_t stage_bytes = fan_stage_bytes(rec_cap, R4);
  int32_t *ep = reinterpret_cast<int32_t *>(wbase + 16);  // [2][2] record-range end points (LDGSTS)
  unsigned char *stage0 = wbase + 32;
  Slot *acc = reinterpret_cast<Slot *>(stage0 + 2 * stage_bytes);

  const int n_chunks = (n_owned + kFanChunk - 1) / kFanChunk;
  const int chunk_stride = gridDim.x * kFanWarps;
  if (lane == 0) {
    ptx::mbar_init(&full[0], 1);
    ptx::mbar_init(&full[1], 1);
    ptx::mbar_init_fence();
  }
  __syncwarp();

  // ---- lane 0: the TMA loads of a chunk.  The end points of its record range are fetched one
  //      chunk ahead with cp.async (global -> shared, no registers held across the compute loop).
  auto request_endpoints = [&](int chunk, int slot) {
    if (chunk < n_chunks) {
      const int32_t n0 = chunk * kFanChunk;
      ptx::cp_async4(ep + 2 * slot, fan_ptr + n0);
      ptx::cp_async4(ep + 2 * slot + 1, fan_ptr + min(n0 + kFanChunk, n_owned));
    }
    ptx::cp_async_commit();
  };
  auto issue = [&](int chunk, int stage, int32_t r0, int32_t r1) {
    const int32_t n0 = chunk * kFanChunk;
    const int32_t base = r0 & ~(RO::kAlign - 1);  // 16-byte aligned start of the record copy
    const uint32_t rec_bytes = (uint32_t)((r1 - base + RO::kAlign - 1) / RO::kAlign) * 16u;
    unsigned char *st = stage0 + stage * stage_bytes;
    ptx::mbar_expect_tx(&full[stage], (uint32_t)kPtrSlices * kFanPtrInts * 4u + rec_bytes);
    ptx::bulk_load(st, adj_ptr + n0, kFanPtrInts * 4u, &full[stage]);
    ptx::bulk_load(st + kFanPtrInts * 4, fan_ptr + n0, kFanPtrInts * 4u, &full[stage]);
    if (R4) ptx::bulk_load(st + 2 * kFanPtrInts * 4, fan_hdr + n0, kFanPtrInts * 4u, &full[stage]);
    if (rec_bytes) ptx::bulk_load(st + kPtrSlices * kFanPtrInts * 4, fan_rec + base, rec_bytes, &full[stage]);
  };
  int chunk = blockIdx.x * kFanWarps + warp;
  if (lane == 0 && chunk < n_chunks) {
    // chunks 0 and 1 of this warp: direct loads (start-up only); chunk 2's end points requested
    for (int q = 0; q < 2; ++q) {
      const int c = chunk + q * chunk_stride;
      if (c < n_chunks) {
        const int32_t n0 = c * kFanChunk;
        issue(c, q, __ldg(fan_ptr + n0), __ldg(fan_ptr + min(n0 + kFanChunk, n_owned)));
      }
    }
    if (!kEpLdg) request_endpoints(chunk + 2 * chunk_stride, 0);
  }

  // ---- per-thread state of the chunk about to be computed (filled by begin_chunk)
  // Coordinates travel through five register sets: the fan's first neighbour, and two pairs (A, B) that
  // alternate between "used by this trip" and "in flight for the next trip" (a trip = two fan steps).  Records are
  // re-read from the ring stage where they are needed (an LDS is cheaper than a register held across a trip).
  double2 p0 = make_double2(0.0, 0.0), pa1 = p0, pa2 = p0, pb1 = p0, pb2 = p0;
  double2 ps = make_double2(0.0, 0.0);
  const Rec *recs = nullptr;
  Slot *my = acc;
  int f = 0, fe = 0, deg = 0;
  int32_t self = 0;
  uint32_t hdr = 0;
  int cur_mat = -1, loaded_mat = -1;  // the material row in `m` survives from chunk to chunk
  MatRow m = {0.0, 0.0, 0.0, 0.0};
  auto fetch = [&](int i, double2 &p) {
    if (i < fe) p = __ldg(coords + RO::nbr(recs[i], self, n_owned));
  };
  // Waits for the chunk's ring slot and puts the first gathers in flight.
  auto begin_chunk = [&](int c, int jj) {
    const int stage = jj & 1;
    ptx::mbar_wait(&full[stage], (uint32_t)((jj >> 1) & 1));
    const int32_t n0 = c * kFanChunk;
    const int n_in = min(kFanChunk, n_owned - n0);
    const unsigned char *st = stage0 + stage * stage_bytes;
    const int32_t *a_sl = reinterpret_cast<const int32_t *>(st);
    const int32_t *f_sl = a_sl + kFanPtrInts;
    recs = reinterpret_cast<const Rec *>(st + kPtrSlices * kFanPtrInts * 4);
    const int32_t base = f_sl[0] & ~(RO::kAlign - 1);
    const int32_t out_lo = a_sl[0];
    f = fe = deg = 0;
    self = n0 + lane;
    if (lane < n_in) {
      ps = __ldg(coords + self);
      f = f_sl[lane] - base;
      fe = f_sl[lane + 1] - base;
      deg = a_sl[lane + 1] - a_sl[lane];
      my = acc + SPB * (a_sl[lane] - out_lo);
      if (R4) hdr = reinterpret_cast<const uint32_t *>(f_sl + kFanPtrInts)[lane];
    }
    fetch(f, p0);
    fetch(f + 1, pa1);
    fetch(f + 2, pa2);
  };

  int j = 0;  // ring position mod 4: stage = j & 1, barrier parity = (j >> 1) & 1
  if (chunk < n_chunks) begin_chunk(chunk, 0);
  for (; chunk < n_chunks; chunk += chunk_stride, j = (j + 1) & 3) {
    const int stage = j & 1;
    const int next = chunk + chunk_stride;

    // lane 0: end points of the record range this trip's refill will need; consumed after the walk, within the same
    // trip (a cp.async group for them shares its scoreboard with other loads: ncu r02 capture Q, 17 - 21 % of the stall
    // samples on an unrelated LDG of begin_chunk)
    int32_t ep0 = 0, ep1 = 0;
    if (kEpLdg && lane == 0 && next + chunk_stride < n_chunks) {
      const int32_t nr = (next + chunk_stride) * kFanChunk;
      ep0 = __ldg(fan_ptr + nr);
      ep1 = __ldg(fan_ptr + min(nr + kFanChunk, n_owned));
    }
    // ---- the fan walk of this thread's node
    if (f < fe) {
      if (R4) cur_mat = RO::first_mat(hdr);
      Val diag = Ops::zero(), X = Ops::zero(), Y;
      const Rec r0 = recs[f];
      int kself = RO::kself(r0, hdr);
      double2 ea = make_double2(p0.x - ps.x, p0.y - ps.y);
      // fan step of record i: element (self, neighbour of record i-1, neighbour of record i at `p`); cin arrives
      // holding the previous element's share of the block towards the previous neighbour and is stored finished,
      // cout = this element's share of the block towards this neighbour
      auto step = [&](int i, const double2 p, const double2 eprev, double2 &ecur, Val &cin, Val &cout) {
        const Rec rc = recs[i];
        ecur = make_double2(p.x - ps.x, p.y - ps.y);
        if (RO::new_mat(rc, hdr, cur_mat) || cur_mat != loaded_mat) {
          m = tab[cur_mat];
          loaded_mat = cur_mat;
        }
        Ops::step(eprev, ecur, m, cin, cout);
        Ops::store(my, deg, RO::k(recs[i - 1]), cin);
        Ops::diag_acc(diag, cin);
      };
      if (!RO::multi(r0)) {
        // a single fan (all nodes but bow-ties and some boundary corners): records f+1 .. fe-1 are its steps and
        // only the last one needs a flag test.  Two steps per trip; the pairs A and B swap roles every trip.
        int i = f + 1;
        double2 e1, e2;
        while (true) {
          if (i + 1 >= fe) break;
          fetch(i + 2, pb1);
          fetch(i + 3, pb2);
          step(i, pa1, ea, e1, X, Y);
          step(i + 1, pa2, e1, e2, Y, X);
          ea = e2;
          i += 2;
          if (i + 1 >= fe) {
            pa1 = pb1;
            break;
          }
          fetch(i + 2, pa1);
          fetch(i + 3, pa2);
          step(i, pb1, ea, e1, X, Y);
          step(i + 1, pb2, e1, e2, Y, X);
          ea = e2;
          i += 2;
        }
        if (i < fe) {  // an odd step left (its coordinate sits in pa1)
          step(i, pa1, ea, e1, X, Y);
          X = Y;
        }
        // X: the last element's block towards the last neighbour (a closed fan's first block waits in its slot)
        if (fe - f > 1) {
          const Rec rl = recs[fe - 1];
          Ops::diag_acc(diag, X);
          if (RO::add_first(rl)) Ops::add(X, Ops::load(my, deg, RO::k(rl)));
          Ops::store(my, deg, RO::k(rl), X);
        }
      } else {
        // general walk: several fans around the node
        for (int i = f + 1; i < fe; ++i) {
          const Rec rc = recs[i];
          const double2 p = __ldg(coords + RO::nbr(rc, self, n_owned));
          double2 e2 = make_double2(p.x - ps.x, p.y - ps.y);
          if (RO::seed(rc)) {  // a chain starts: its first neighbour, nothing carried
            kself = RO::kself(rc, hdr);
            X = Ops::zero();
          } else {
            step(i, p, ea, e2, X, Y);
            X = Y;
            if (RO::last(rc)) {
              Ops::diag_acc(diag, Y);
              if (RO::add_first(rc)) Ops::add(Y, Ops::load(my, deg, RO::k(rc)));
              Ops::store(my, deg, RO::k(rc), Y);
            }
          }
          ea = e2;
        }
      }
      Ops::store(my, deg, kself, diag);
    }

    // ---- the sub-tile is complete: the exact image of vals[dim^2 * out_lo ...)
    int32_t out_lo, out_len;  // node-level block range of this chunk (slice still in the ring slot)
    {
      const int32_t *a_sl = reinterpret_cast<const int32_t *>(stage0 + stage * stage_bytes);
      out_lo = a_sl[0];
      out_len = a_sl[min(kFanChunk, n_owned - chunk * kFanChunk)] - out_lo;
    }
    // kEarly: the next chunk's first loads go out NOW and travel while this chunk's store and the ring refill are issued
    if (kEarly && next < n_chunks) begin_chunk(next, (j + 1) & 3);
    ptx::fence_async_smem();  // generic smem accesses of this chunk ordered before the async proxy
    __syncwarp();
    if (KC == 2) {
      // 1 DOF per node: the destination is only 8-byte aligned -> plain coalesced copy
      const double *src = reinterpret_cast<const double *>(acc);
      double *dst = vals + out_lo;
      for (int q = lane; q < out_len; q += 32) dst[q] = src[q];
    } else if (lane == 0 && out_len > 0) {
      ptx::bulk_store(vals + 4 * (int64_t)out_lo, acc, (uint32_t)out_len * 32u);  // one TMA bulk store
    }
    if (lane == 0) {
      // this warp is done with ring slot `stage`: refill it with the chunk after the next one
      const int nn = next + chunk_stride;
      if (kEpLdg) {
        if (nn < n_chunks) issue(nn, stage, ep0, ep1);
      } else {  // (the end points were requested a whole chunk ago and sit in ep[stage])
        ptx::cp_async_wait_all();
        if (nn < n_chunks) issue(nn, stage, ep[2 * stage], ep[2 * stage + 1]);
        request_endpoints(nn + chunk_stride, stage ^ 1);
      }
    }
    // first gathers of the next chunk go out before we wait for the store to drain the sub-tile
    if (!kEarly && next < n_chunks) begin_chunk(next, (j + 1) & 3);
    if (KC != 2 && lane == 0) ptx::bulk_store_wait_read();
    __syncwarp();
  }
}

