#!/usr/bin/env python3
"""tools/sweep.py -- per-kernel timings on device-resident CG data (needs a B200).
    python tools/sweep.py [--nx 256] [--reps 20] [--configs 0,1,2]   -> one JSON line per measurement"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import minife_b200 as mf

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=256)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--configs", default="0")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--flags", default="1")
ap.add_argument("--line-tiles", default="-1")
args = ap.parse_args()
n = args.nx
g = mf.Box([0, n, 0, n, 0, n])
mesh = mf.simple_mesh_description(g, g)
A = mf.CSRMatrix()
mf.generate_matrix_structure(mesh, A); mf.assemble_FE_data(mesh, A, None)
mf.impose_dirichlet(0.0, A, None, bc_rows=0); mf.impose_dirichlet(1.0, A, None, bc_rows=1); mf.make_local_matrix(A)
N, nnz = A.num_rows(), A.num_nonzeros()
it, nr, t = mf.cg_solve(A, None, None, None, args.iters, 2.2e-16)     # leaves realistic r, p, Ap on the device
print(json.dumps({"setup": {p: mesh.phase_ms(p) for p in ("structure", "assembly", "assembly_elem", "assembly_gather", "dirichlet0", "dirichlet1", "make_local")},
                  "cg": {"iters": it, "normr": nr, "WAXPY": t[0], "DOT": t[1], "MATVEC": t[2], "TOTAL": t[4]}}))
mvb = 12.0 * nnz + 4.0 * (N + 1) + 16.0 * N
for cfg in [int(c) for c in args.configs.split(",") if c != ""]:
    mesh.set_option("matvec_config", cfg)
    for fl in (0,):
        for what in ("matvec", "matvec_dot"):
            ms = mesh.time_kernel(what, args.reps)
            print(json.dumps({"kernel": what, "config": cfg, "ms": ms, "GBs": mvb / ms / 1e6}), flush=True)
mesh.set_option("matvec_config", 0)
for what, byts in (("stream_probe", 12.0 * nnz), ("stream_probe21", 12.0 * nnz), ("matvec_simple", mvb), ("waxpby", 24.0 * N), ("xr_update", 48.0 * N), ("dot_rr", 8.0 * N), ("dot_pap", 16.0 * N)):
    for mode in ((0, 1, 2) if what.startswith("dot") else (0,)):
        if mode == 1 and n > 64:
            continue
        mesh.set_dot_mode(mode)
        for nw in ((16, 24) if (what.startswith("dot") and mode == 0) else (0,)):
            mesh.set_option("dot_maps_warps", nw)
            ms = mesh.time_kernel(what, args.reps)
            print(json.dumps({"kernel": what, "dot_mode": mode, "dot_maps_warps": nw, "ms": ms, "GBs": byts / ms / 1e6}), flush=True)
mesh.set_option("dot_maps_warps", 0)
mesh.close()
