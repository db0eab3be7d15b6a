"""tools/mv_prof_case.py -- the case ncu was pointed at when the fused SpMV was 35 % slow: rank 0 of a
2-rank 400^3 run on one GPU (halo supplied by the caller), fused kernel (overlap 1) then the two-launch form (overlap 2).
    ncu --set full --clock-control none -k regex:k_matvec -c 6 -o prof python tools/mv_prof_case.py"""
import os, sys
sys.path.insert(0, "/root/repo")
import minife_b200 as mf
n=[400,400,400]; g = mf.Box([0, n[0], 0, n[1], 0, n[2]])
boxes = mf.box_partition(0, 2, 2, g)
mesh = mf.simple_mesh_description(g, boxes[0], rank=0, nranks=2, all_boxes=boxes)
A = mf.CSRMatrix(); mf.generate_matrix_structure(mesh, A); mf.assemble_FE_data(mesh, A, None)
mf.impose_dirichlet(0.0, A, None, bc_rows=0); mf.impose_dirichlet(1.0, A, None, bc_rows=1); mf.make_local_matrix(A)
mesh.set_option("test_no_exchange", 1)
for ov in (1, 2):
    mesh.set_option("overlap", ov)
    print(ov, mesh.time_kernel("matvec_op", 3))
mesh.close()
