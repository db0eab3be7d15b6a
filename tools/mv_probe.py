#!/usr/bin/env python3
"""tools/mv_probe.py --n NXxNYxNZ --np P --rank R -- SpMV of ONE rank of a P-rank partition on cuda:0 with the halo supplied
by the caller (test_no_exchange): isolates the kernel variants (all tiles / fused one-launch / two launches) from the
communication."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minife_b200 as mf
ap = argparse.ArgumentParser(); ap.add_argument("--n", default="400x400x400"); ap.add_argument("--np", type=int, default=2); ap.add_argument("--rank", type=int, default=0)
args = ap.parse_args()
n = [int(v) for v in args.n.split("x")]
g = mf.Box([0, n[0], 0, n[1], 0, n[2]])
boxes = mf.box_partition(0, args.np, 2, g)
mesh = mf.simple_mesh_description(g, boxes[args.rank], rank=args.rank, nranks=args.np, all_boxes=boxes)
A = mf.CSRMatrix(); mf.generate_matrix_structure(mesh, A); mf.assemble_FE_data(mesh, A, None)
mf.impose_dirichlet(0.0, A, None, bc_rows=0); mf.impose_dirichlet(1.0, A, None, bc_rows=1); mf.make_local_matrix(A)
mesh.set_option("test_no_exchange", 1)
nr, nnz = A.num_rows(), A.num_nonzeros()
byts = 12.0 * nnz + 4.0 * (nr + 1) + 16.0 * nr
print("box", boxes[args.rank], "rows %d nnz %d  make_local %.1f ms" % (nr, nnz, mesh.phase_ms("make_local")))
for ring, ov in [(1, o_) for o_ in (0, 1, 2, 4)]:
    mesh.set_option("overlap", ov if ov < 3 else 1)
    mesh.set_option("dbg_no_ext", {3: 1, 4: 2}.get(ov, 0))      # 3: fused without the halo-pointer select, 4: fused without the boundary pass
    for what in ("matvec_op",):
        ms = mesh.time_kernel(what, 20)
        print("ring %d overlap %d %-10s %.4f ms  %.0f GB/s" % (ring, ov, what, ms, byts / ms / 1e6))
mesh.close()
