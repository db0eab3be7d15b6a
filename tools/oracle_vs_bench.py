#!/usr/bin/env python3
"""tools/oracle_vs_bench.py -- ties records made on the CPU after a GPU run to that run: the bench line carries the final
residual of its 200-iteration solve with 17 digits (config.final_resid); the record (tests/golden/big_*.npz from the
reference, oracle_*.npz from the oracle) holds the whole residual history.  Prints, per bench line under profiles/, whether
the 200th residual of the record of the same mesh and np is the same double.  No GPU needed."""
import glob, json, os, re
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
print("%-52s %-18s %-32s %-24s %s" % ("bench line", "mesh, np", "record", "final residual (bench)", "record's 200th residual"))
for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_*.json"))):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        m = re.search(r"miniFE (\d+)x(\d+)x(\d+)", l["config"]["workload"]); npr = int(l["n_gpus"]); fr = l["config"]["final_resid"]
    except Exception:
        continue
    dims = tuple(int(v) for v in m.groups())
    for kind in ("big", "oracle"):
        g = os.path.join(ROOT, "tests", "golden", "%s_%dx%dx%d_np%d.npz" % ((kind,) + dims + (npr,)))
        if os.path.exists(g):
            r = float(np.load(g)["normr"][-1])
            print("%-52s %-18s %-32s %-24r %s" % (os.path.basename(f), "%dx%dx%d np=%d" % (dims + (npr,)), os.path.basename(g), fr, "same double" if r == fr else "DIFFERENT %r" % r))
