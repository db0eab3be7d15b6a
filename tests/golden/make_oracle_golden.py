#!/usr/bin/env python3
"""tests/golden/make_oracle_golden.py -- records of the ORACLE (oracle/minife_oracle.c, the plain-C restatement that is
pinned bit for bit to the unmodified reference wherever the reference can run) at sizes the reference itself cannot hold
in this container: 512^3 needs about 65 GB with the reference's std::map / std::set bookkeeping at any np, the oracle
needs 51 GB.  One process simulates all np ranks (serial-order dot per rank, rank-ordered sum: Vector_functions.hpp:194-226).

    python tests/golden/make_oracle_golden.py 512x512x512:8          # -> tests/golden/oracle_512x512x512_np8.npz (~30 min, 1 thread)
    python tests/golden/make_oracle_golden.py --check 256x256x256:8  # same run, compared with the REFERENCE's own record
                                                                     #    tests/golden/big_256x256x256_np8.npz instead of stored
    python tests/golden/make_oracle_golden.py --wide 512x512x512:1   # one rank, 3.6e9 nonzeros: the oracle built with 64-bit row
                                                                     #    offsets (oracle/Makefile oracle-wide); --wide --check
                                                                     #    400x400x400:1 pins that build to the reference's record

The file holds the same keys as tests/golden/big_*.npz (make_big_golden.py) except the CRCs of row_offsets / cols / coefs
(copying those arrays out would not fit the memory of the container next to the oracle's own): sizes, boxes, CRC-32 of b
before / after the Dirichlet conditions and of x, the index maps, and the four per-iteration histories of the
200-iteration solve.  `source` = "oracle".
bench.py falls back to such a file when no reference record exists for its configuration and says so in `parity`.
"""
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib  # noqa: E402
from oracle_lib import Oracle, EPS  # noqa: E402

MAPS = ("external_index", "external_local_index", "elements_to_send", "neighbors", "recv_length", "send_length")


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


class WideOracle(Oracle):
    """oracle/libminife_oracle_wide.so: the same source compiled with -DMFO_WIDE_OFFSETS (64-bit row offsets, `make -C oracle
    oracle-wide`), for one rank holding 2^31 nonzeros or more (512^3 at np = 1)."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            import subprocess
            subprocess.run(["make", "-s", "-C", os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"), "oracle-wide"], check=True)
            Oracle._lib = None                             # let the base class bind its prototypes to the wide library
            saved = oracle_lib.ORACLE_SO
            oracle_lib.ORACLE_SO = saved.replace("libminife_oracle.so", "libminife_oracle_wide.so")
            try:
                cls._lib = Oracle.lib.__func__(Oracle)
            finally:
                oracle_lib.ORACLE_SO = saved
                Oracle._lib = None
        return cls._lib


def oracle_case(n, npr, max_iter=200, wide=False):
    t0 = time.time()
    out = dict(n=np.array(n), np=npr, source="oracle (64-bit row offsets)" if wide else "oracle")
    o = (WideOracle if wide else Oracle)(n[0], n[1], n[2], npr)
    o.generate_matrix_structure()
    for r in range(npr):
        out["box_%d" % r] = o.box(r)
        out["sizes_%d" % r] = np.array([o.nrows(r), o.nnz(r)], np.int64)
    print("structure %.0f s" % (time.time() - t0), flush=True)
    o.assemble_FE_data()
    for r in range(npr):
        out["crc_b_prebc_%d" % r] = crc(o.b(r))
    print("assembly %.0f s" % (time.time() - t0), flush=True)
    o.impose_dirichlet(0)
    o.impose_dirichlet(1)
    for r in range(npr):
        out["crc_b_%d" % r] = crc(o.b(r))
    o.make_local_matrix()
    for r in range(npr):
        out["num_cols_%d" % r] = np.int32(o.num_cols(r))
        for m in MAPS:
            a = o.get_map(m, r)
            if m in ("neighbors", "recv_length", "send_length"):
                out["%s_%d" % (m, r)] = a
            else:
                out["crc_%s_%d" % (m, r)] = crc(a); out["len_%s_%d" % (m, r)] = len(a)
    print("make_local %.0f s" % (time.time() - t0), flush=True)
    h = o.cg_solve(max_iter, EPS)
    for k in ("normr", "rtrans", "pap", "alpha"):
        out[k] = h[k]
    for r in range(npr):
        out["crc_x_%d" % r] = crc(o.x(r))
    out["wall_s"] = time.time() - t0
    print("cg %d iterations, final normr %.17g, %.0f s" % (h["num_iters"], h["normr"][-1], out["wall_s"]), flush=True)
    return out


def main():
    check, wide = "--check" in sys.argv[1:], "--wide" in sys.argv[1:]
    for spec in [a for a in sys.argv[1:] if not a.startswith("--")]:
        dims, npr = spec.split(":")
        n = tuple(int(v) for v in dims.split("x")); npr = int(npr)
        if check:
            ref = np.load(os.path.join(HERE, "big_%dx%dx%d_np%d.npz" % (n + (npr,))))
            d = oracle_case(n, npr, wide=wide)
            bad = []
            for k, v in d.items():
                if k in ("wall_s", "source"):
                    continue
                a, b = np.asarray(v), np.asarray(ref[k])
                if k in ("pap", "alpha"):
                    a, b = a[1:], b[1:]                   # iteration 0 is never written by the reference
                same = a.shape == b.shape and (np.array_equal(a.view(np.int64), b.view(np.int64)) if a.dtype == np.float64 else np.array_equal(a, b))
                if not same:
                    bad.append(k)
            print("%s np=%d: oracle vs the reference's record: %s" % (dims, npr, "all %d keys equal (histories bit for bit)" % (len(d) - 2) if not bad else "DIFFERENT: %s" % bad), flush=True)
            if bad:
                sys.exit(1)
            continue
        fn = os.path.join(HERE, "oracle_%dx%dx%d_np%d.npz" % (n + (npr,)))
        if os.path.exists(fn):
            print("have", fn, flush=True)
            continue
        np.savez_compressed(fn, **oracle_case(n, npr, wide=wide))
        print("wrote", fn, flush=True)


if __name__ == "__main__":
    main()
