"""TEST / MEASUREMENT INFRASTRUCTURE -- not product code.

SURVEY §8d "CPU baseline (1)": the LITERAL reference (its unmodified modules, loaded by
oracle/ref_loader.py) timed on structured meshes of ~5 k / 20 k / 45 k triangles: element constructors,
k_matrix_data (analysis.py:324-339), create_matrix (:617-663), create_source_matrix (:665-708) and the
spsolve of solve() (:798-830), one core (the reference is single-threaded by construction).  The reference
cannot travel to the GPU box, so this runs in the build container and the result is committed:

    python -m oracle.time_literal_reference > profiles/r02_literal_reference_cpu.json
"""
import json
import os
import platform
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402
from oracle.numpy_oracle import structured_mesh  # noqa: E402


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor()


def run(ns, nx, ny):
    coords, conn = structured_mesh(nx, ny)
    h = 1.0 / ny
    groups = [dict(start=0, stop=len(conn), params=(210e9, 0.25, 7860, 1))]
    t0 = time.perf_counter()
    an, mesh, elems = ref_loader.build_reference_analysis(
        ns, coords, conn, groups, "elasticity",
        node_loads=[(j * (nx + 1) + nx, -1000 * h, 2) for j in range(ny + 1)],
        node_bcs=[(j * (nx + 1), 0, d) for j in range(ny + 1) for d in (1, 2)],
        plane_strain=False, plane_stress=True)
    t_ctor = time.perf_counter() - t0
    t0 = time.perf_counter()
    an.k_matrix_data()
    t_kdata = time.perf_counter() - t0
    t0 = time.perf_counter()
    k = an.create_matrix()
    t_create = time.perf_counter() - t0
    t0 = time.perf_counter()
    an.create_source_matrix()
    t_rhs = time.perf_counter() - t0
    t0 = time.perf_counter()
    an.solve()              # create_matrix + create_source_matrix + spsolve (analysis.py:798-830)
    t_solve_total = time.perf_counter() - t0
    n_el = len(conn)
    return dict(nx=nx, ny=ny, triangles=n_el, ndof=2 * len(coords), nnz_augmented=int(k.nnz),
                ctor_s=t_ctor, k_matrix_data_s=t_kdata, create_matrix_s=t_create, create_source_matrix_s=t_rhs,
                solve_total_s=t_solve_total, spsolve_s=max(t_solve_total - t_create - t_rhs, 0.0),
                k_matrix_data_us_per_element=1e6 * t_kdata / n_el,
                assembly_melem_per_s=n_el / t_create / 1e6,
                solve_melem_per_s=n_el / t_solve_total / 1e6)


def main():
    ns = ref_loader.load()
    import scipy
    rows = [run(ns, nx, ny) for nx, ny in ((70, 36), (142, 71), (212, 106))]
    print(json.dumps(dict(what="literal reference (unmodified /root/reference modules, volmdlr geometry restated by "
                               "oracle/ref_loader.py), structured plane-stress meshes, 1 core",
                          cpu=cpu_model(), cores_used=1, nproc=os.cpu_count(), numpy=np.__version__, scipy=scipy.__version__,
                          when=time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), rows=rows), indent=1))


if __name__ == "__main__":
    main()
