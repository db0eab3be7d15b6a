#!/usr/bin/env python3
"""tools/asm_probe.py --nx N -- setup-phase timings on cuda:0 for both assembly modes (mfe_phase_ms)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minife_b200 as mf
ap = argparse.ArgumentParser(); ap.add_argument("--nx", type=int, default=256); ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
n = args.nx
g = mf.Box([0, n, 0, n, 0, n])
for mode in (0, 1):
    for rep in range(args.reps):
        mesh = mf.simple_mesh_description(g, g)
        mesh.set_option("assembly_mode", mode)
        A = mf.CSRMatrix(); mf.generate_matrix_structure(mesh, A); mf.assemble_FE_data(mesh, A, None)
        mf.impose_dirichlet(0.0, A, None, bc_rows=0); mf.impose_dirichlet(1.0, A, None, bc_rows=1); mf.make_local_matrix(A)
        print("nx %d mode %d rep %d: structure %.2f assembly %.2f (elem %.2f gather %.2f) dirichlet %.2f + %.2f ms" % (
            n, mode, rep, mesh.phase_ms("structure"), mesh.phase_ms("assembly"), mesh.phase_ms("assembly_elem"), mesh.phase_ms("assembly_gather"),
            mesh.phase_ms("dirichlet0"), mesh.phase_ms("dirichlet1")), flush=True)
        mesh.close()
