"""TEST INFRASTRUCTURE -- mints tests/golden/*.npz by running the reference's
UNMODIFIED modules (oracle/ref_loader.py) on the fixtures the reference's own
scripts define.  Run in the build container only:

    python -m oracle.make_golden

The fixtures carry both the flat inputs (coords, conn, materials, load / BC
records) and the reference's outputs (element matrices, augmented CSR, source
vector, spsolve solution, K / M blocks).  /root/reference is never read by the
tests; they read these files.
"""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader  # noqa: E402
from oracle.numpy_oracle import structured_mesh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def read_msh41_triangles(path, etype_wanted=2):
    """Minimal gmsh 4.1 ASCII reader: nodes in file order, 3-node triangles (type 2) or, with
    etype_wanted=4, 4-node tetrahedra (coordinates then keep z)."""
    with open(path) as fh:
        lines = [ln.strip() for ln in fh]
    i = lines.index("$Nodes") + 1
    nblocks, nnodes = int(lines[i].split()[0]), int(lines[i].split()[1])
    i += 1
    tags, xyz = [], []
    for _ in range(nblocks):
        _, _, _, nb = (int(t) for t in lines[i].split())
        i += 1
        tags.extend(int(lines[i + k]) for k in range(nb))
        i += nb
        xyz.extend([float(t) for t in lines[i + k].split()] for k in range(nb))
        i += nb
    assert len(tags) == nnodes
    tag_to_idx = {t: k for k, t in enumerate(tags)}
    coords = np.array(xyz, dtype=np.float64)[:, :(3 if etype_wanted == 4 else 2)]
    i = lines.index("$Elements") + 1
    nblocks = int(lines[i].split()[0])
    i += 1
    tris = []
    for _ in range(nblocks):
        _, _, etype, nb = (int(t) for t in lines[i].split())
        i += 1
        if etype == etype_wanted:
            for k in range(nb):
                t = [int(v) for v in lines[i + k].split()]
                tris.append([tag_to_idx[v] for v in t[1:(5 if etype_wanted == 4 else 4)]])
        i += nb
    return coords, np.array(tris, dtype=np.int32)


def csr_parts(m, prefix):
    m = m.tocsr()
    m.sum_duplicates()
    m.sort_indices()
    return {prefix + "_indptr": m.indptr.astype(np.int64), prefix + "_indices": m.indices.astype(np.int64),
            prefix + "_data": m.data.astype(np.float64), prefix + "_shape": np.array(m.shape, dtype=np.int64)}


def run(ns, spec, modal_k=0):
    kind = spec["kind"]
    bounds = spec["group_bounds"]
    groups = [dict(start=bounds[g], stop=bounds[g + 1], params=tuple(spec["group_params"][g]))
              for g in range(len(bounds) - 1)]
    ps = spec.get("plane")
    plane_strain = None if ps is None else ps == "strain"
    plane_stress = None if ps is None else ps == "stress"
    an, mesh, elems = ref_loader.build_reference_analysis(
        ns, spec["coords"], spec["conn"], groups, kind,
        node_loads=spec.get("node_loads", ()), node_bcs=spec.get("node_bcs", ()),
        elements_loads=spec.get("elements_loads", ()), edge_loads=spec.get("edge_loads", ()),
        edge_bcs=spec.get("edge_bcs", ()), element_bcs=spec.get("element_bcs", ()),
        plane_strain=plane_strain, plane_stress=plane_stress, magnet_loads=spec.get("magnet_loads", ()))
    out = {}
    if kind in ("elasticity", "elasticity3d"):
        out["ke"] = np.array([e.elementary_matrix(plane_strain, plane_stress) for e in elems])
        out["me"] = np.array([e.elementary_mass_matrix() for e in elems])
    else:
        out["ke"] = np.array([e.elementary_matrix() for e in elems], dtype=np.float64)
    if kind != "elasticity3d":   # tetrahedra have no element_to_node_factors in the reference
        out["factors"] = np.array([e.element_to_node_factors() for e in elems], dtype=np.float64)
    kaug = an.create_matrix()
    out.update(csr_parts(kaug, "kaug"))
    out["f"] = an.create_source_matrix()
    if spec.get("magnet_loads"):   # analysis.py:556-577 on its own (data, node rows), in emission order
        data, rows = an.source_c_matrix_magnet_loads()
        out["magnet_data"] = np.array(data, dtype=np.float64)
        out["magnet_rows"] = np.array(rows, dtype=np.int64)
    out["x"] = np.array(an.solve().result_vector, dtype=np.float64)
    # post-processing of the reference (results.py:809-830, :769-781, :121-152) on its own solution
    if kind == "elasticity":
        er = ns.results.ElasticityResults2D(mesh, list(out["x"]), plane_strain, plane_stress)
        out["strain"] = np.array([er.strain[e] for e in elems], dtype=np.float64)
        out["stress"] = np.array([er.stress[e] for e in elems], dtype=np.float64)
        out["energy"] = np.array([er.energy_per_element[e] for e in elems], dtype=np.float64)
    elif kind == "elasticity3d":
        er = ns.results.ElasticityResults3D(mesh, list(out["x"]), plane_strain, plane_stress)
        out["strain"] = np.array([er.strain[e] for e in elems], dtype=np.float64)
        out["stress"] = np.array([er.stress[e] for e in elems], dtype=np.float64)
        out["energy"] = np.array([er.energy_per_element[e] for e in elems], dtype=np.float64)
        out["disp_node1"] = np.array(list(er.displacement_vectors_per_node[mesh.nodes[1]]), dtype=np.float64)
    elif kind == "magnetic":
        mr = ns.results.MagneticResults(mesh, list(out["x"]))
        out["bfield"] = np.array([[mr.magnetic_field_per_element[e][0], mr.magnetic_field_per_element[e][1]]
                                  for e in elems], dtype=np.float64)
    bcs = an._boundary_conditions
    out["bc_dofs"] = np.array([an.positions[(mesh.node_to_index[b.application], b.dimension)] for b in bcs],
                              dtype=np.int64)
    out["bc_vals"] = np.array([b.value for b in bcs], dtype=np.float64)
    out.update(csr_parts(an.k_matrix_sparse(), "k"))
    if kind == "elasticity3d":
        out.update(csr_parts(an.m_matrix_sparse(), "m"))
        if modal_k:
            vals, vecs = an.modal_analysis('largest', modal_k)
            out["eig_largest"] = np.sort(np.real(vals))
    if kind == "elasticity":
        out.update(csr_parts(an.m_matrix_sparse(), "m"))
        if modal_k:
            vals, vecs = an.modal_analysis('largest', modal_k)
            out["eig_largest"] = np.sort(np.real(vals))
    return out


def spec_arrays(spec):
    """Flatten a spec into npz-storable arrays (+ one JSON blob for ragged records)."""
    e_count = len(spec["conn"])
    bounds = spec["group_bounds"]
    mat_id = np.zeros(e_count, dtype=np.int32)
    for g in range(len(bounds) - 1):
        mat_id[bounds[g]:bounds[g + 1]] = g
    if spec["kind"] in ("elasticity", "elasticity3d"):  # reference ctor order (E, nu, rho, t) -> flat (E, nu, t, rho)
        mat = np.array([[p[0], p[1], p[3], p[2]] for p in spec["group_params"]], dtype=np.float64)
    else:
        mat = np.array([[p[0], 0, 0, 0] for p in spec["group_params"]], dtype=np.float64)
    rec = {k: [list(map(_py, r)) for r in spec.get(k, ())]
           for k in ("node_loads", "node_bcs", "elements_loads", "edge_loads", "edge_bcs", "element_bcs",
                     "magnet_loads")}
    meta = dict(name=spec["name"], kind=spec["kind"], plane=spec.get("plane"), records=rec,
                group_bounds=[int(b) for b in bounds], source=spec.get("source", ""))
    return dict(coords=np.asarray(spec["coords"], dtype=np.float64), conn=np.asarray(spec["conn"], dtype=np.int32),
                mat_id=mat_id, mat=mat, meta=np.array(json.dumps(meta)))


def _py(v):
    if isinstance(v, (list, tuple, np.ndarray)):
        return [int(t) for t in v]
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    return v


def fixtures():
    steel = (210e9, 0.25, 7860, 1)
    # 1. scripts/Elasticity/beam2d_example_1.py:18-40 -- 2-triangle plate, first-seen node order
    yield dict(name="plate2_pstress", kind="elasticity", plane="stress",
               source="scripts/Elasticity/beam2d_example_1.py:18-40",
               coords=np.array([[3, 0], [3, 2], [0, 0], [0, 2]], float), conn=np.array([[0, 1, 2], [3, 2, 1]]),
               group_bounds=[0, 2], group_params=[(30e6, 0.25, 2.7, 0.5)],
               node_loads=[(1, -1000, 2)],
               node_bcs=[(0, 0, 2), (3, 0, 1), (3, 0, 2), (2, 0, 1), (2, 0, 2)]), 0
    # 2. scripts/Elasticity/beam2d_example_2.py:33-95 -- 18-triangle beam, 3 groups, plane strain
    coords, conn = structured_mesh(9, 1, h=1.0)
    left = [n for n in range(len(coords)) if coords[n, 0] == 0]
    tip = int(np.where((coords[:, 0] == 9) & (coords[:, 1] == 1))[0][0])
    yield dict(name="beam18_pstrain", kind="elasticity", plane="strain",
               source="scripts/Elasticity/beam2d_example_2.py:33-95",
               coords=coords, conn=conn, group_bounds=[0, 6, 12, 18], group_params=[steel] * 3,
               node_loads=[(tip, -10000000, 2)],
               node_bcs=[(n, 0, d) for n in left for d in (1, 2)]), 6
    # 3. scripts/Magnetic/finite_element_beam.py:14-71 -- 3-material bar, ElementsLoad on 2 elements
    mu0 = 4 * math.pi * 1e-7
    right = [n for n in range(len(coords)) if math.isclose(coords[n, 0], 9, abs_tol=1e-6)]
    yield dict(name="magbar18", kind="magnetic", plane=None,
               source="scripts/Magnetic/finite_element_beam.py:14-71",
               coords=coords, conn=conn, group_bounds=[0, 6, 12, 18],
               group_params=[(mu0 * 100000,), (mu0,), (mu0 * 50000,)],
               elements_loads=[([0, 1], 1e10, 1)],
               node_bcs=[(n, 0, 1) for n in right]), 0
    # 4. scripts/Elasticity/beam2d_example_3.py:35-106 -- gmsh cantilever sweep
    for lc in ("0.8", "0.5", "0.3", "0.18", "0.1"):
        path = os.path.join(ref_loader.REFERENCE_ROOT, "scripts", "InputFiles", "2D", f"beam_2d_{lc}.msh")
        c, t = read_msh41_triangles(path)
        tip = int(np.where((c[:, 0] == 10) & (c[:, 1] == 1))[0][0])
        left = [n for n in range(len(c)) if c[n, 0] == 0]
        yield dict(name=f"gmsh_beam_{lc}", kind="elasticity", plane="stress",
                   source=f"scripts/Elasticity/beam2d_example_3.py:35-106 on InputFiles/2D/beam_2d_{lc}.msh",
                   coords=c, conn=t, group_bounds=[0, len(t)], group_params=[(30e6, 0.25, 2.7, 1)],
                   node_loads=[(tip, -1000, 2)],
                   node_bcs=[(n, 0, d) for n in left for d in (1, 2)]), (8 if lc in ("0.5", "0.3") else 0)
    # 5. API semantics: jittered 6x4 mesh, two materials, every load / BC record type,
    #    duplicate keys (last-wins, analysis.py:42-43,:85-86), non-zero BC values
    coords, conn = structured_mesh(6, 4, jitter=0.2, seed=0)
    nx = 6
    nid = lambda i, j: j * (nx + 1) + i  # noqa: E731
    yield dict(name="semantics_elast", kind="elasticity", plane="stress", source="synthetic (SURVEY §8a-7..a-9)",
               coords=coords, conn=conn, group_bounds=[0, 20, 48],
               group_params=[(210e9, 0.25, 7860, 1.0), (70e9, 0.33, 2700, 0.5)],
               node_loads=[(nid(6, 4), -1000.0, 2), (nid(6, 3), 250.0, 1), (nid(6, 4), -400.0, 2)],
               elements_loads=[([40, 41, 42], 900.0, 2)],
               edge_loads=[(nid(6, 0), nid(6, 1), 300.0, 1), (nid(6, 1), nid(6, 2), 500.0, 1)],
               node_bcs=[(nid(0, j), 0.0, d) for j in range(5) for d in (1, 2)] + [(nid(0, 2), 1e-4, 1)],
               edge_bcs=[(nid(3, 0), nid(4, 0), 2e-4, 2)],
               element_bcs=[(5, 1e-3, 2)]), 5
    yield dict(name="semantics_elast_pstrain", kind="elasticity", plane="strain", source="synthetic",
               coords=coords, conn=conn, group_bounds=[0, 20, 48],
               group_params=[(210e9, 0.25, 7860, 1.0), (70e9, 0.33, 2700, 0.5)],
               node_loads=[(nid(6, 4), -1000.0, 2)],
               node_bcs=[(nid(0, j), 0.0, d) for j in range(5) for d in (1, 2)]), 0
    yield dict(name="semantics_mag", kind="magnetic", plane=None, source="synthetic",
               coords=coords, conn=conn, group_bounds=[0, 16, 32, 48],
               group_params=[(mu0 * 1e5,), (mu0,), (mu0 * 5e4,)],
               elements_loads=[([0, 1, 2, 3], 1e10, 1), ([2, 3, 14], -4e9, 1)],
               node_loads=[(nid(3, 2), 7.0, 1)],
               edge_loads=[(nid(2, 4), nid(3, 4), 11.0, 1)],
               node_bcs=[(nid(6, j), 0.0, 1) for j in range(5)],
               edge_bcs=[(nid(0, 0), nid(0, 1), 3.0, 1)],
               element_bcs=[(47, 2.0, 1)]), 0
    # 5b. MagnetLoad (loads.py:105-147, analysis.py:556-577): a 2x2-cell magnet block inside the jittered
    #     6x4 mesh (its centre node listed as non-contour), a second single-cell magnet whose
    #     non_contour_nodes cover one whole edge (that edge must drop out), plus an ordinary source
    cells = lambda i0, i1, j0, j1: [2 * (j * nx + i) + t for j in range(j0, j1) for i in range(i0, i1) for t in (0, 1)]  # noqa: E731
    yield dict(name="semantics_mag_magnet", kind="magnetic", plane=None,
               source="synthetic (MagnetLoad: loads.py:105-147, analysis.py:556-577)",
               coords=coords, conn=conn, group_bounds=[0, 16, 32, 48],
               group_params=[(mu0 * 1.05,), (mu0,), (mu0 * 5e4,)],
               magnet_loads=[(cells(2, 4, 1, 3), [nid(3, 2)], 2.5e5, 7.65e5),
                             (cells(0, 1, 3, 4), [nid(0, 3), nid(0, 4)], -1.0e5, 3.0e5)],
               elements_loads=[([46, 47], 3e6, 1)],
               node_bcs=[(nid(6, j), 0.0, 1) for j in range(5)] + [(nid(i, 0), 0.0, 1) for i in range(6)]), 0
    # 6. mid-size structured meshes (uniform: exact zeros in K; jittered: general geometry)
    for nm, jit in (("struct24x16", 0.0), ("struct24x16_jit", 0.2)):
        coords, conn = structured_mesh(24, 16, jitter=jit, seed=0)
        nx, ny = 24, 16
        h = 1.0 / ny
        yield dict(name=nm + "_pstress", kind="elasticity", plane="stress", source="synthetic (SURVEY §8d)",
                   coords=coords, conn=conn, group_bounds=[0, len(conn)], group_params=[steel],
                   node_loads=[(j * (nx + 1) + nx, -1000 * h, 2) for j in range(ny + 1)],
                   node_bcs=[(j * (nx + 1), 0, d) for j in range(ny + 1) for d in (1, 2)]), 8
        third = (len(conn) // 3) // 2 * 2
        yield dict(name=nm + "_mag", kind="magnetic", plane=None, source="synthetic (SURVEY §8d)",
                   coords=coords, conn=conn, group_bounds=[0, third, 2 * third, len(conn)],
                   group_params=[(mu0 * 1e5,), (mu0,), (mu0 * 5e4,)],
                   elements_loads=[([0, 1], 1e10, 1)],
                   node_bcs=[(j * (nx + 1) + nx, 0, 1) for j in range(ny + 1)]), 0


def fixtures_3d():
    """P1 tetrahedra (ElasticityTetrahedralElement3D, elements.py:663-876)."""
    from oracle.numpy_oracle import structured_tet_mesh
    steel, alu = (210e9, 0.25, 7860, 1), (70e9, 0.33, 2700, 1)

    def clamp_and_load(coords, xmax, zmax):
        left = [n for n in range(len(coords)) if coords[n, 0] == 0]
        tip = [n for n in range(len(coords)) if coords[n, 0] == xmax and coords[n, 2] == zmax]
        return ([(n, -1000.0, 3) for n in tip] + [(tip[0], 250.0, 1)],
                [(n, 0, d) for n in left for d in (1, 2, 3)])

    coords, conn = structured_tet_mesh(2, 2, 2, h=1.0, jitter=0.2, seed=3)
    loads, bcs = clamp_and_load(coords, 2.0, 2.0)
    yield dict(name="tet_cube2_jit", kind="elasticity3d", plane="stress", source="synthetic (Kuhn cube, 2 materials)",
               coords=coords, conn=conn, group_bounds=[0, 24, len(conn)], group_params=[steel, alu],
               node_loads=loads, node_bcs=bcs + [(26, 1e-4, 2)]), 6
    coords, conn = structured_tet_mesh(6, 2, 2, h=0.5, jitter=0.15, seed=4)
    loads, bcs = clamp_and_load(coords, 3.0, 1.0)
    yield dict(name="tet_beam6x2x2_jit", kind="elasticity3d", plane="stress", source="synthetic (Kuhn beam)",
               coords=coords, conn=conn, group_bounds=[0, len(conn)], group_params=[steel],
               node_loads=loads, node_bcs=bcs), 8
    # scripts/Elasticity/beam3d_example_2.py:52-98 on the two small gmsh beams (10 x 2 x 2)
    for lc in ("0.5", "1"):
        path = os.path.join(ref_loader.REFERENCE_ROOT, "scripts", "InputFiles", "3D", f"beam3d_{lc}.msh")
        c, t = read_msh41_triangles(path, etype_wanted=4)
        loads = [(n, -1000.0, 3) for n in range(len(c)) if c[n, 0] == 10 and c[n, 2] == 2]
        bcs = [(n, 0, d) for n in range(len(c)) if c[n, 0] == 0 for d in (1, 2, 3)]
        yield dict(name=f"gmsh_beam3d_{lc}", kind="elasticity3d", plane="stress",
                   source=f"scripts/Elasticity/beam3d_example_2.py:52-98 on InputFiles/3D/beam3d_{lc}.msh",
                   coords=c, conn=t, group_bounds=[0, len(t)], group_params=[(30e6, 0.25, 2.7, 1)],
                   node_loads=loads, node_bcs=bcs), 0


def main():
    ns = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    only3d = "--3d" in sys.argv   # (re)mint only the tetrahedral fixtures
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None   # ... or one fixture by name
    for spec, modal_k in (list(fixtures_3d()) if only3d else list(fixtures()) + list(fixtures_3d())):
        if only is not None and spec["name"] != only:
            continue
        out = run(ns, spec, modal_k)
        arrays = spec_arrays(spec)
        arrays.update({"ref_" + k: v for k, v in out.items()})
        path = os.path.join(OUT, spec["name"] + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{spec['name']:28s} E={len(spec['conn']):5d} N={len(spec['coords']):5d} "
              f"aug={out['kaug_shape'][0]:5d} nnz={len(out['kaug_data']):6d} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
