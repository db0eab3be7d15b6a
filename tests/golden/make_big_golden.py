#!/usr/bin/env python3
"""tests/golden/make_big_golden.py -- goldens of the UNMODIFIED reference at sizes whose arrays cannot be committed:
the reference (oracle/_ref/ref_dump_mpi.x = ref/src headers + the MPI shim, REF_DUMP_LEAN=1) runs the whole pipeline
at np ranks and leaves, per rank, CRC-32s of row_offsets / cols / coefs / b / x, the index maps, and -- rank 0 -- the
four per-iteration histories of the 200-iteration solve.  Stored as tests/golden/big_<nx>x<ny>x<nz>_np<N>.npz
(a few KB each: the maps are stored as CRCs too).

    python tests/golden/make_big_golden.py 256x256x512:2 256x256x256:8 400x400x400:1 ...

Needs /root/reference (authoring container only).  Memory: about 0.45 KB per mesh node over all ranks together
(256^3: 8 GB, 400^3: 30 GB); 512^3 (65 GB) does not fit the 62 GB of this container at any np.
"""
import os
import subprocess
import sys
import tempfile
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "mpi_shim"))
import launch  # noqa: E402

REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump_mpi.x")
CRC_NAMES = ("row_offsets", "cols_global", "coefs_prebc", "b_prebc", "coefs", "b", "cols_local", "x")
MAPS = ("external_index", "external_local_index", "elements_to_send", "neighbors", "recv_length", "send_length")


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def big_case(n, npr, max_iter=200, timeout=4 * 3600):
    out = dict(n=np.array(n), np=npr)
    with tempfile.TemporaryDirectory() as d:
        os.environ["REF_DUMP_LEAN"] = "1"
        t0 = time.time()
        rc = launch.launch(npr, [REF_DUMP, str(n[0]), str(n[1]), str(n[2]), d, str(max_iter)], timeout=timeout, stdout=subprocess.DEVNULL)
        if rc != 0:
            raise RuntimeError("ref_dump_mpi.x failed rc=%d" % rc)
        out["wall_s"] = time.time() - t0
        for r in range(npr):
            f = lambda name, ext, dt: np.fromfile(os.path.join(d, "%s.%d.%d.%s" % (name, npr, r, ext)), dtype=dt)
            c = f("crc", "u32", np.uint32)
            for k, v in zip(CRC_NAMES, c):
                out["crc_%s_%d" % (k, r)] = v
            out["box_%d" % r] = f("box", "i32", np.int32)
            out["num_cols_%d" % r] = f("num_cols", "i32", np.int32)[0]
            out["sizes_%d" % r] = f("sizes", "i64", np.int64)          # nrows, nnz
            for m in MAPS:
                a = f(m, "i32", np.int32)
                if m in ("neighbors", "recv_length", "send_length"):
                    out["%s_%d" % (m, r)] = a
                else:
                    out["crc_%s_%d" % (m, r)] = crc(a); out["len_%s_%d" % (m, r)] = len(a)
        for h in ("normr", "rtrans", "pap", "alpha"):
            out[h] = np.fromfile(os.path.join(d, "%s_hist.%d.0.f64" % (h, npr)))
    return out


def main():
    for spec in sys.argv[1:]:
        dims, npr = spec.split(":")
        n = tuple(int(v) for v in dims.split("x"))
        fn = os.path.join(HERE, "big_%dx%dx%d_np%d.npz" % (n + (int(npr),)))
        if os.path.exists(fn):
            print("have", fn, flush=True)
            continue
        d = big_case(n, int(npr))
        np.savez_compressed(fn, **d)
        print("%s: %d iterations, final normr %.17g, %.0f s" % (os.path.basename(fn), len(d["normr"]) - 1, d["normr"][-1], d["wall_s"]), flush=True)


if __name__ == "__main__":
    main()
