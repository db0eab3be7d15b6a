#!/usr/bin/env python3
"""tests/golden/make_golden.py -- regenerate the golden fixtures from the UNMODIFIED reference (oracle/_ref, built
from /root/reference by oracle/Makefile).  Needs /root/reference (authoring container only); the fixtures are
committed so that the tests run where the reference is absent.

    python tests/golden/make_golden.py [--big]     # --big adds the 128^3 and 256^3 CG histories (minutes)
"""
import argparse
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Ref, ref_dump_mpi, EPS  # noqa: E402


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def serial_case(n, full):
    r = Ref(*n)
    r.generate_matrix_structure()
    rows, ro, cols, _ = r.csr()
    r.assemble_FE_data()
    coefs_prebc, b_prebc = r.csr()[3].copy(), r.b()
    r.impose_dirichlet(0); r.impose_dirichlet(1)
    coefs, b = r.csr()[3].copy(), r.b()
    r.make_local_matrix()
    h = r.cg_solve(200, EPS)
    x = r.x()
    d = dict(n=np.array(n), nrows=r.nrows(), nnz=r.nnz(), num_iters=h["num_iters"], final_normr=h["final_normr"],
             normr=h["normr"], pap=h["pap"], alpha=h["alpha"], crc_row_offsets=crc(ro), crc_cols=crc(cols), crc_coefs_prebc=crc(coefs_prebc),
             crc_b_prebc=crc(b_prebc), crc_coefs=crc(coefs), crc_b=crc(b), crc_x=crc(x))
    if full:
        d.update(row_offsets=ro, cols=cols, coefs_prebc=coefs_prebc, b_prebc=b_prebc, coefs=coefs, b=b, x=x)
    return d


def history_case(n):
    r = Ref(n, n, n).setup()
    h = r.cg_solve(200, EPS)
    return dict(n=np.array([n] * 3), num_iters=h["num_iters"], final_normr=h["final_normr"], normr=h["normr"], pap=h["pap"], alpha=h["alpha"], crc_x=crc(r.x()))


def overlap_case(n, npr):
    """mv_overlap_comm_comp=1 (driver.hpp:278-282): rearrange_matrix_local_external + matvec_overlap change the summation
    order of the rows that have external columns, hence the history."""
    d = ref_dump_mpi(*n, npr, overlap=True)
    out = dict(n=np.array(n), np=npr, normr=d["normr_hist"][0], pap=d["pap_hist"][0])
    for r in range(npr):
        out["crc_x_%d" % r] = crc(d["x"][r]); out["x_%d" % r] = d["x"][r]
    return out


def mpi_case(n, npr, imbalance=0):
    d = ref_dump_mpi(*n, npr, imbalance=imbalance)
    out = dict(n=np.array(n), np=npr)
    if imbalance:
        out["imbalance"] = imbalance
    for r in range(npr):
        for k in ("box", "rows", "row_offsets", "cols_global", "cols_local", "external_index", "external_local_index", "elements_to_send",
                  "neighbors", "recv_length", "send_length", "num_cols"):
            out["%s_%d" % (k, r)] = d[k][r]
        out["crc_coefs_%d" % r] = crc(d["coefs"][r]); out["crc_b_%d" % r] = crc(d["b"][r]); out["crc_x_%d" % r] = crc(d["x"][r])
    out["normr"] = d["normr_hist"][0]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    for n, full in (((1, 1, 2), True), ((4, 4, 4), True), ((5, 3, 7), True), ((17, 3, 5), False), ((20, 30, 40), False)):
        np.savez_compressed(os.path.join(HERE, "serial_%dx%dx%d.npz" % n), **serial_case(n, full))
    for n in (32, 64):
        np.savez_compressed(os.path.join(HERE, "history_%d.npz" % n), **history_case(n))
    # the last three are the reference CI's own `mpirun -np 2` cases (.github/workflows/compare-ref-kokkos.yml:13-25,52)
    for n, npr in (((2, 2, 2), 2), ((2, 2, 2), 8), ((8, 8, 8), 8), ((6, 6, 6), 4), ((5, 3, 7), 3), ((1, 1, 2), 2), ((17, 3, 5), 2), ((20, 30, 40), 2)):
        np.savez_compressed(os.path.join(HERE, "mpi_%dx%dx%d_np%d.npz" % (n + (npr,))), **mpi_case(n, npr))
    # boxes changed by the reference's own add_imbalance (load_imbalance percent), ref/src/driver.hpp:137-139
    for n, npr, imb in (((12, 12, 12), 8, 30), ((10, 9, 8), 4, 60), ((9, 14, 11), 6, 40)):
        np.savez_compressed(os.path.join(HERE, "mpi_%dx%dx%d_np%d_imb%d.npz" % (n + (npr, imb))), **mpi_case(n, npr, imb))
    for n, npr in (((8, 8, 8), 8), ((20, 30, 40), 2), ((12, 10, 9), 4), ((5, 3, 7), 3)):
        np.savez_compressed(os.path.join(HERE, "mpiov_%dx%dx%d_np%d.npz" % (n + (npr,))), **overlap_case(n, npr))
    if args.big:
        for n in (128, 256):
            np.savez_compressed(os.path.join(HERE, "history_%d.npz" % n), **history_case(n))
            print("history", n, "done", flush=True)


if __name__ == "__main__":
    main()
