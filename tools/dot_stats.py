#!/usr/bin/env python3
"""tools/dot_stats.py --nx N [--iters K] -- exact-dot diagnostics of one cg_solve on cuda:0 (mfe_dot_stats)."""
import argparse, ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minife_b200 as mf
ap = argparse.ArgumentParser(); ap.add_argument("--nx", type=int, default=256); ap.add_argument("--iters", type=int, default=20); ap.add_argument("--warps", type=int, default=0)
args = ap.parse_args()
n = args.nx
g = mf.Box([0, n, 0, n, 0, n]); mesh = mf.simple_mesh_description(g, g)
A = mf.CSRMatrix(); mf.generate_matrix_structure(mesh, A); mf.assemble_FE_data(mesh, A, None)
mf.impose_dirichlet(0.0, A, None, bc_rows=0); mf.impose_dirichlet(1.0, A, None, bc_rows=1); mf.make_local_matrix(A)
mesh.set_option('dot_maps_warps', args.warps); mesh.set_option('x_overlap', 0)
L = mf.lib(); st = np.zeros(16 + 1024, np.uint64)
L.mfe_dot_stats(mesh._h, st.ctypes.data_as(C.c_void_p), len(st), 1)
it, nr, t = mf.cg_solve(A, None, None, None, args.iters, 2.2e-16)
L.mfe_dot_stats(mesh._h, st.ctypes.data_as(C.c_void_p), len(st), 0)
print("nx", n, "iters", it, "dots", int(st[0]), "dirty sum/max", int(st[1]), int(st[2]), "full replays", int(st[3]), "slow gaps", int(st[4]),
      "walk ms total/max", st[5] / 1e6, st[6] / 1e6, "phases us/dot: prologue %.1f gaps %.1f serial %.1f" % (st[11] / max(st[0], 1) / 1e3, st[12] / max(st[0], 1) / 1e3, st[14] / max(st[0], 1) / 1e3), "slowest dot #", int(st[7]), "slow pieces", int(st[8]), "cg DOT_s", t[1], "TOTAL_s", t[4])
for i in range(min(int(st[0]), 40)):
    w = int(st[16 + 2 * i])
    try:
        print("  dot %3d: dirty %4d slow gaps %3d walk %.3f ms" % (i, w & 0xffffffff, w >> 32, st[17 + 2 * i] / 1e6))
    except BrokenPipeError:
        break
for what in ("dot_rr", "dot_pap", "r_update", "xr_update", "waxpby", "matvec_dot"):
    print("  time_kernel %-10s %.2f us" % (what, 1e3 * mesh.time_kernel(what, 20)))
mesh.close()
