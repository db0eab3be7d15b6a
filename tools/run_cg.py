#!/usr/bin/env python3
"""tools/run_cg.py -- set up an nx^3 problem on cuda:0 and run one cg_solve (for ncu captures).
    python tools/run_cg.py [--nx 256] [--iters 5] [--config 4] [--dot-mode 0]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minife_b200 as mf

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=256)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--config", type=int, default=-1)
ap.add_argument("--dot-mode", type=int, default=0)
ap.add_argument("--solves", type=int, default=1)
args = ap.parse_args()
n = args.nx
g = mf.Box([0, n, 0, n, 0, n])
mesh = mf.simple_mesh_description(g, g)
if args.config >= 0:
    mesh.set_option("matvec_config", args.config)
mesh.set_dot_mode(args.dot_mode)
A = mf.CSRMatrix()
mf.generate_matrix_structure(mesh, A); mf.assemble_FE_data(mesh, A, None)
mf.impose_dirichlet(0.0, A, None, bc_rows=0); mf.impose_dirichlet(1.0, A, None, bc_rows=1); mf.make_local_matrix(A)
for _ in range(args.solves):
    it, nr, t = mf.cg_solve(A, None, None, None, args.iters, 2.2e-16)
print(json.dumps({"iters": it, "normr": nr, "WAXPY": t[0], "DOT": t[1], "MATVEC": t[2], "TOTAL": t[4], "launches": mesh.launch_count()}))
mesh.close()
