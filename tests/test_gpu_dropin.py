"""Drop-in test (-m gpu): the UNMODIFIED reference driver (ref/src/main.cpp + driver.hpp), compiled against the shadow
operator headers of minife_b200/host and linked to libminife_b200.so (build/miniFE_b200*.x, built where the
reference tree exists), must reproduce the reference executable: same YAML numbers and -- with MINIFE_DEBUG -- the
same A/b/x dump files the reference's own CI diffs between variants (.github/workflows/compare-ref-kokkos.yml)."""
import glob
import os
import re
import subprocess
import tempfile

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "build", "miniFE_b200.x")
OURS_DBG = os.path.join(ROOT, "build", "miniFE_b200_debug.x")
REF = os.path.join(ROOT, "oracle", "_ref", "miniFE_ref_serial.x")
REF_DBG = os.path.join(ROOT, "oracle", "_ref", "miniFE_ref_serial_debug.x")
STANDALONE = os.path.join(ROOT, "build", "minife_b200_standalone.x")
OURS_FUSED = os.path.join(ROOT, "build", "miniFE_b200_fused.x")
REF_FUSED = os.path.join(ROOT, "oracle", "_ref", "miniFE_ref_serial_fused.x")
OURS_KERNELS = os.path.join(ROOT, "build", "miniFE_b200_kernels.x")
OURS_ELL_DBG = os.path.join(ROOT, "build", "miniFE_b200_ell_debug.x")
REF_ELL_DBG = os.path.join(ROOT, "oracle", "_ref", "miniFE_ref_serial_ell_debug.x")


def run(exe, args, cwd):
    out = subprocess.run([exe] + args, cwd=cwd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (exe, out.stdout[-2000:], out.stderr[-2000:])
    y = glob.glob(os.path.join(cwd, "*.yaml"))
    return out.stdout, (open(y[0]).read() if y else "")


def grab(text, key):
    m = re.search(r"%s:\s*(\S+)" % re.escape(key), text)
    assert m, key
    return m.group(1)


@pytest.mark.skipif(not (os.path.exists(OURS) and os.path.exists(REF)), reason="drop-in binaries not built (need the reference tree at build time)")
@pytest.mark.parametrize("args", [["nx=10", "verify_solution=1"], ["nx=32"], ["nx=20", "ny=30", "nz=40"], ["nx=48", "mv_overlap_comm_comp=1"]])
def test_unmodified_driver_reproduces_reference_yaml(args):
    with tempfile.TemporaryDirectory() as a, tempfile.TemporaryDirectory() as b:
        so, yo = run(OURS, args, a)
        sr, yr = run(REF, [x for x in args if not x.startswith("mv_overlap")], b)
        for key in ("Iterations", "Final Resid Norm", "Global Nrows", "Global NNZ", "Total CG Flops", "MATVEC Flops"):
            assert grab(yo, key) == grab(yr, key), key
        res_o = re.findall(r"(Initial Residual = \S+|Iteration = \d+\s+Residual = \S+|Final Resid Norm: \S+)", so)
        res_r = re.findall(r"(Initial Residual = \S+|Iteration = \d+\s+Residual = \S+|Final Resid Norm: \S+)", sr)
        if any(a.startswith("mv_overlap") for a in args):      # driver.hpp:279-282 prints no "Final Resid Norm:" line on that branch
            res_r = [l for l in res_r if not l.startswith("Final")]
        assert res_o == res_r and len(res_o) >= 3
        if "verify_solution=1" in args:     # device verify_solution (shadow of ref/fem/verify_solution.hpp) prints the reference's verdict
            msg = "solution matches analytic solution to within"
            assert msg in sr and msg in so


@pytest.mark.skipif(not (os.path.exists(OURS_DBG) and os.path.exists(REF_DBG)), reason="debug drop-in binaries not built")
@pytest.mark.parametrize("args", [["nx=1", "ny=1", "nz=2"], ["nx=17", "ny=3", "nz=5"], ["nx=20", "ny=30", "nz=40"]])
def test_debug_dumps_identical_to_reference(args):
    """The CI cases of the reference (compare-ref-kokkos.yml:13-25); files are compared byte for byte."""
    with tempfile.TemporaryDirectory() as a, tempfile.TemporaryDirectory() as b:
        run(OURS_DBG, args, a)
        run(REF_DBG, args, b)
        names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(b, "*.mtx.*")) + glob.glob(os.path.join(b, "*.vec.*")))
        assert len(names) >= 7, names
        for n in names:
            assert open(os.path.join(a, n)).read() == open(os.path.join(b, n)).read(), n


@pytest.mark.skipif(not os.path.exists(STANDALONE), reason="standalone driver not built")
def test_standalone_driver():
    with tempfile.TemporaryDirectory() as a:
        out = subprocess.run([STANDALONE, "nx=32"], cwd=a, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        assert grab(out.stdout, "Iterations") == "106" and grab(out.stdout, "Global NNZ") == "912673"
        assert abs(float(grab(out.stdout, "Final Resid Norm")) - 1.2517e-16) < 1e-20


@pytest.mark.skipif(not (os.path.exists(OURS_FUSED) and os.path.exists(REF_FUSED)), reason="MINIFE_FUSED drop-in binaries not built")
@pytest.mark.parametrize("args", [["nx=12"], ["nx=20", "ny=30", "nz=40"]])
def test_minife_fused_build_reproduces_reference(args):
    """-DMINIFE_FUSED (matvec_and_dot + fused_waxpby, ref/src/cg_solve.hpp:172-213): same iteration count and residuals, and the
    YAML carries the MATVECDOT rows with a nonzero time (the SpMV + p.Ap dot interval)."""
    with tempfile.TemporaryDirectory() as a, tempfile.TemporaryDirectory() as b:
        so, yo = run(OURS_FUSED, args, a)
        sr, yr = run(REF_FUSED, args, b)
        for key in ("Iterations", "Final Resid Norm", "Global NNZ", "MATVECDOT Flops"):
            assert grab(yo, key) == grab(yr, key), key
        assert float(grab(yo, "MATVECDOT Time")) > 0.0
        res_o = re.findall(r"(Initial Residual = \S+|Iteration = \d+\s+Residual = \S+)", so)
        res_r = re.findall(r"(Initial Residual = \S+|Iteration = \d+\s+Residual = \S+)", sr)
        assert res_o == res_r and len(res_o) >= 2


@pytest.mark.skipif(not os.path.exists(OURS_KERNELS), reason="MINIFE_KERNELS drop-in binary not built")
def test_minife_kernels_build_runs_time_kernels():
    """-DMINIFE_KERNELS=1 (driver.hpp:257-270 -> time_kernels): the reference's own ref/src cannot be built in this mode
    (time_kernels.hpp:75 needs the basic/ Vector), so the check is against the oracle's kernels: the driver reports
    500 'iterations' and Final Resid Norm = xdotp = 500 x dot(A*p, p) with p = x0 (x0 = 0 except x0[0] = 0.9)."""
    import numpy as np
    from oracle_lib import Oracle
    with tempfile.TemporaryDirectory() as a:
        so, yo = run(OURS_KERNELS, ["nx=10"], a)
        assert grab(yo, "Iterations") == "500"
        o = Oracle(10, 10, 10).setup()
        p = np.zeros(o.num_cols()); p[0] = 0.9
        x = o.matvec(p)
        d = Oracle.dot(x, p[:len(x)])
        want = 0.0
        for _ in range(500):
            want += d
        assert abs(float(grab(yo, "Final Resid Norm")) - want) <= 1e-5 * abs(want)


@pytest.mark.skipif(not (os.path.exists(OURS_ELL_DBG) and os.path.exists(REF_ELL_DBG)), reason="ELL drop-in binaries not built")
@pytest.mark.parametrize("args", [["nx=1", "ny=1", "nz=2"], ["nx=6"], ["nx=17", "ny=3", "nz=5"], ["nx=20", "ny=30", "nz=40"]])
def test_ell_matrix_build_reproduces_reference(args):
    """-DMINIFE_ELL_MATRIX (ref/src/ELLMatrix.hpp): the unmodified driver instantiated with ELLMatrix against
    the shadow headers.  Same YAML numbers (Global NNZ counts the padding: rows x 27), same residual print-out, and
    byte-identical A / b / x dumps (the ELL matvec's padding terms 0.0 * x[0] are applied on the device)."""
    with tempfile.TemporaryDirectory() as a, tempfile.TemporaryDirectory() as b:
        so, yo = run(OURS_ELL_DBG, args, a)
        sr, yr = run(REF_ELL_DBG, args, b)
        for key in ("Iterations", "Final Resid Norm", "Global Nrows", "Global NNZ", "Total CG Flops", "MATVEC Flops"):
            assert grab(yo, key) == grab(yr, key), key
        res = r"(Initial Residual = \S+|Iteration = \d+\s+Residual = \S+|Final Resid Norm: \S+)"
        assert re.findall(res, so) == re.findall(res, sr)
        names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(b, "*.mtx.*")) + glob.glob(os.path.join(b, "*.vec.*")))
        assert len(names) >= 7, names
        for n in names:
            assert open(os.path.join(a, n)).read() == open(os.path.join(b, n)).read(), n


OURS_NP = os.path.join(ROOT, "build", "miniFE_b200_debug_np.x")
REF_NP = os.path.join(ROOT, "oracle", "_ref", "miniFE_ref_mpi_debug.x")


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(not (os.path.exists(OURS_NP) and os.path.exists(REF_NP)), reason="multi-rank drop-in binaries not built")
@pytest.mark.parametrize("nproc,args", [(2, ["nx=1", "ny=1", "nz=2"]), (2, ["nx=17", "ny=3", "nz=5"]), (2, ["nx=20", "ny=30", "nz=40"]),
                                        (2, ["nx=20", "ny=30", "nz=40", "mv_overlap_comm_comp=1"]), (4, ["nx=20", "ny=30", "nz=40"]),
                                        (8, ["nx=24", "ny=20", "nz=16"])])
def test_multi_rank_dropin_dumps_identical_to_reference(nproc, args):
    """The reference CI's own `mpirun -np 2 miniFE.x` parity job (compare-ref-kokkos.yml:51-52,79,95): the UNMODIFIED main.cpp
    + driver.hpp built with -DHAVE_MPI against the shadow headers (tests/Makefile; MPI = the test shim), one process per
    rank, rank r on GPU r -- against the reference built the same way on the CPU.  Every A/b/x dump of every rank is compared
    byte for byte, and the YAML iteration count / residual too."""
    if _gpus() < nproc:
        pytest.skip("needs %d GPUs, this box has %d" % (nproc, _gpus()))
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle", "mpi_shim"))
    import launch
    with tempfile.TemporaryDirectory() as a, tempfile.TemporaryDirectory() as b:
        with open(os.path.join(a, "stdout.txt"), "w") as fa:
            assert launch.launch(nproc, [OURS_NP] + args, cwd=a, timeout=900, stdout=fa) == 0, open(os.path.join(a, "stdout.txt")).read()[-2000:]
        with open(os.path.join(b, "stdout.txt"), "w") as fb:
            assert launch.launch(nproc, [REF_NP] + args, cwd=b, timeout=900, stdout=fb) == 0
        names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(b, "*.mtx.*")) + glob.glob(os.path.join(b, "*.vec.*")))
        assert len(names) >= 7 * nproc, names
        for n in names:
            assert open(os.path.join(a, n)).read() == open(os.path.join(b, n)).read(), n
        ya, yb = open(glob.glob(os.path.join(a, "*.yaml"))[0]).read(), open(glob.glob(os.path.join(b, "*.yaml"))[0]).read()
        for key in ("Iterations", "Final Resid Norm", "Global Nrows", "Global NNZ"):
            assert grab(ya, key) == grab(yb, key), key
