// fe_plan: per-mesh symbolic data built once by fe_plan_create (plan.cu) and consumed by
// the numeric assembly kernels (assemble.cu).
#pragma once
#include "common.cuh"

namespace fe {
constexpr int kFan4FieldBits = 19, kFan4Shift = 32 - kFan4FieldBits;
enum : uint32_t { FAN4_SEED = 1, FAN4_MATSW = 2, FAN4_ADD_FIRST = 4, FAN4_LAST = 8, FAN4_MULTI = 16 };
constexpr int kTetStageNodes = 16;  // owned nodes per tile of the staged tetrahedral assembly (tet.cu)
constexpr int kTile = 128;  // nodes (= threads) per CTA of the tiled assembly kernels
}

struct fe_plan {
  fe_ctx *ctx = nullptr;
  int32_t n_nodes = 0, n_owned = 0, dim = 0;
  int64_t n_elems = 0;
  int64_t n_corners = 0;  // (element, local vertex) pairs whose vertex is an owned node
  int64_t nnzb = 0;       // node-level adjacency entries (incl. self) over owned rows
  int64_t nnz = 0;        // nnzb * dim * dim
  int32_t max_degree = 0;
  int32_t max_mat_id = 0;  // largest material index any element refers to (fe_assemble checks n_mat against it)
  int64_t bytes = 0;
  // node -> corners, counting-sorted by node, ascending element id within a node
  int32_t *corner_ptr = nullptr;  // [n_owned + 1]
  // corner record: .x = (elem << 2) | local_vertex
  //                .y = k0 | k1 << 8 | k2 << 16 | first0 << 24 | first1 << 25 | first2 << 26
  //   k_j   = position of the element's j-th vertex in this node's sorted adjacency list
  //   first_j = this corner is the first (lowest element id) contributor to that block
  int2 *corner_rec = nullptr;  // [n_corners]
  // node-level sorted unique adjacency (block-CSR pattern of the owned rows)
  int32_t *adj_ptr = nullptr;  // [n_owned + 1]
  int32_t *adj = nullptr;      // [nnzb]
  // connectivity padded to 16 B with the material id: one LDG.128 per corner visit
  int4 *conn4 = nullptr;  // [n_elems] {n0, n1, n2, mat_id}
  // fan-ordered corner records (plan.cu: fan_walk); valid when fan_ok
  bool fan_ok = false;
  int64_t n_fan = 0;
  int32_t fan_tile_max = 0;    // max records of one 32-node chunk (one warp of k_assemble_fan)
  int32_t *fan_ptr = nullptr;  // [n_owned + 1]
  int2 *fan_rec = nullptr;     // [n_fan]
  // the same records in 4 bytes (plan.cu: k_fan_compact) when the numbering is banded (|neighbour - node| <
  // 2^18, ghost columns included) and no node star holds more than two materials (ids < 4096): half the
  // record traffic of the assembly kernel.
  //   word        = k | FAN4_* flags << 8 | (neighbour - node) << 13   (signed 19-bit difference)
  //   fan_hdr[i]  = k_self | mat0 << 8 | mat1 << 20;  a node's walk starts on mat0 and FAN4_MATSW on a step
  //                 switches to the other material before the step is evaluated; FAN4_MULTI on a node's first
  //                 record: more than one fan around the node (the kernel's general loop)
  bool fan_compact_ok = false;
  uint32_t *fan_rec4 = nullptr;  // [n_fan]
  uint32_t *fan_hdr = nullptr;   // [n_owned]
  // linear tetrahedra (npe == 4, dim == 3; fe_tet_plan_create): per OFF-DIAGONAL block (node i, slot k) the
  // elements that hold both nodes, ascending, as (element << 4 | local vertex of i << 2 | local vertex of the
  // neighbour) -- what k_tet_assemble_slots walks; the diagonal block is summed from the same visits
  int32_t npe = 3;
  int32_t *corner_elem = nullptr;   // [n_corners] element of every corner, ascending within a node (variants 1, 2)
  int32_t *contrib_ptr = nullptr;   // [nnzb + 1]
  int32_t *contrib = nullptr;       // [12 n_elems restricted to owned rows]
  int64_t n_contrib = 0;
  bool tet_degenerate = false;      // an element lists a node twice: the slot kernel is not used
  // staged variant (k_tet_assemble_staged): tiles of kTetStageNodes consecutive owned nodes.  Per tile the distinct
  // nodes and elements its rows touch (an element belongs to the tile of every owned node it holds), the elements'
  // connectivity in tile-local node numbers, and the contribution codes re-based on the tile-local element
  // index -- everything the kernel stages in shared memory arrives as contiguous, coalesced slices.
  bool tet_stage_ok = false;         // false: a tile exceeds the 12-bit element / 16-bit node index or the caps
  int32_t n_tiles = 0;
  int32_t *tile_eptr = nullptr;      // [n_tiles + 1] -> tile_elist / tile_erec
  int32_t *tile_nptr = nullptr;      // [n_tiles + 1] -> tile_nodes
  int4 *tile_desc = nullptr;         // [2 n_tiles] (adj offset, adj length, element offset, elements), (node offset,
                                     // nodes, contribution offset, contributions): all a CTA needs to address its slices
  int32_t *tile_elist = nullptr;     // global element id (material lookup)
  ushort4 *tile_erec = nullptr;      // the element's four nodes as positions in the tile's node list
  int32_t *tile_nodes = nullptr;     // ascending global node ids
  uint16_t *contrib16 = nullptr;     // [n_contrib] tile-local element index << 4 | vi << 2 | vj
  uint8_t *tet_kself = nullptr;      // [n_owned] position of the node in its own adjacency row
  int32_t tile_elems_max = 0, tile_nodes_max = 0, tile_contrib_max = 0, tile_adj_max = 0;
};
