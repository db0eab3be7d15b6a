#!/usr/bin/env python3
"""bench.py -- miniFE hot path on B200: CG GFLOP/s (reference flop model) and HBM GB/s at nx=ny=nz=256 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--nx 256] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

One "step" = one full cg_solve (200 iterations of matvec + 2 dot + 3 waxpby, ref/src/cg_solve.hpp:69-229) on the
assembled 256^3-per-GPU system, everything resident in HBM (`value`); `e2e` is the same solve through the
reference-facing call with HOST buffers (x0 uploaded from pinned memory, x downloaded).  Weak scaling: 256^3
elements per GPU, partitioned with miniFE's own box_partition.  Inputs are far larger than L2 (7.4 GB of traffic
per iteration vs 126 MB), so no explicit L2 flush is needed.  Rank 0 prints ONE JSON line.
"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
EPS = float(np.finfo(np.float64).eps)
MAX_ITERS = 200


def weak_dims(nx, ngpus):
    """256^3 elements per GPU; box_partition splits z, then y, then x (ref/utils/BoxPartition.cpp:96)."""
    d = [nx, nx, nx]
    axis, n = 2, ngpus
    while n > 1:
        if n % 2:
            raise SystemExit("--gpus must be a power of two")
        d[axis] *= 2
        axis = (axis + 2) % 3
        n //= 2
    return d


def flop_model(dims, nnz_global, iters):
    """ref/src/driver.hpp:322-341 (n = nx*ny*nz elements, the reference's own convention)."""
    n = dims[0] * dims[1] * dims[2]
    return 2.0 * nnz_global * (iters + 1) + 2.0 * n * (2 * iters) + 3.0 * n * (3 * iters + 2)


def algorithmic_bytes(nrows, nnz, iters):
    """matvec 12*nnz + 4*(N+1) + 16*N; dot(r,r) 8N; dot(Ap,p) 16N; waxpby 24N."""
    mv = 12.0 * nnz + 4.0 * (nrows + 1) + 16.0 * nrows
    return dict(matvec=mv, total=mv * (iters + 1) + (8.0 + 16.0) * nrows * iters + 24.0 * nrows * (3 * iters + 2))


class ClockSampler:
    """SM clock, power and throttle reasons of one GPU DURING the timed region (B200_PROFILING.md recipe).  Sampled through
    NVML inside this process (a thread, one sample per 250 ms; no child process polling the driver).  Falls back to an
    `nvidia-smi -lms 500` child without pynvml."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period=0.25):
        self.index, self.rows, self.proc, self.period, self.how = index, [], None, period, None
        self._stop = threading.Event()

    def _nvml_loop(self, nv, h):
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.rows.append((float(sm), float(self.max_sm), pw, [n for n, bit in names if rs & bit]))
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates all GPUs of the box; CUDA_VISIBLE_DEVICES may remap the CUDA index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.how = "nvml"
            self.t = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.t.start()
            return
        except Exception:
            self.how = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi"
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.rows.append((float(r[1]), float(r[2]), float(r[3]), [nm for nm, v in zip(names, r[5:9]) if v.lower().startswith("active")]))
            except (ValueError, IndexError):
                continue

    def stop(self):
        self._stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.how is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"]}
        try:
            self.t.join(timeout=2)
        except Exception:
            pass
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "how": self.how}
        sm = [r[0] for r in self.rows]; mx = [r[1] for r in self.rows]; pw = [r[2] for r in self.rows]
        reasons = sorted({x for r in self.rows for x in r[3]})
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)), "samples": len(sm), "reasons": reasons,
                "how": "%s, one sample per %.0f ms during the timed solves" % (self.how, 1e3 * self.period if self.how == "nvml" else 500)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own OpenMP build (oracle/_ref/miniFE_ref_openmp.x, compiled unmodified from
# /root/reference/openmp/src) on all host threads, on a bounded sample of the workload.
# ------------------------------------------------------------------------------------------------------------
def run_reference_cpu(dims, threads=None):
    """ONE run of the reference's own OpenMP executable (openmp/src, unmodified) on `dims`; its YAML report supplies the
    200-iteration CG timing (the sample), Mat-struc-gen Time and FE assembly Time."""
    if isinstance(dims, int):
        dims = [dims] * 3
    exe = os.path.join(ROOT, "oracle", "_ref", "miniFE_ref_openmp.x")
    kind = "reference"
    threads = threads or os.cpu_count() or 1
    if not os.path.exists(exe):
        return None
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="spread")
    with tempfile.TemporaryDirectory() as d:
        t0 = time.time()
        out = subprocess.run([exe, "nx=%d" % dims[0], "ny=%d" % dims[1], "nz=%d" % dims[2]], cwd=d, env=env, capture_output=True, text=True, timeout=3000)
        wall = time.time() - t0
        ytxt = ""
        for fn in os.listdir(d):
            if fn.endswith(".yaml"):
                with open(os.path.join(d, fn)) as f:
                    ytxt = f.read()

    def grab(key):
        m = re.search(r"%s:\s*([-+0-9.eE]+|inf)" % re.escape(key), ytxt)
        return float(m.group(1)) if m and m.group(1) != "inf" else None
    if out.returncode != 0 or grab("Total CG Mflops") is None:
        return None
    return dict(kind=kind, cores=threads, dims=list(dims), wall_s=wall, iters=grab("Iterations"), cg_mflops=grab("Total CG Mflops"), cg_time=grab("Total CG Time"),
                matvec_mflops=grab("MATVEC Mflops"), struct_time=grab("Mat-struc-gen Time"), assembly_time=grab("FE assembly Time"),
                final_resid=grab("Final Resid Norm"))


def run_oracle_port_cpu(nx):
    """Fallback when oracle/_ref is absent: the plain-C restatement (1 thread), on a smaller cube."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    o = Oracle(nx, nx, nx).setup()
    nnz = o.nnz()
    t0 = time.time()
    h = o.cg_solve(MAX_ITERS, EPS)
    dt = time.time() - t0
    return dict(kind="port", cores=1, dims=[nx] * 3, wall_s=dt, iters=h["num_iters"], cg_mflops=1e-6 * flop_model([nx] * 3, nnz, h["num_iters"]) / dt, cg_time=dt,
                struct_time=None, assembly_time=None, matvec_mflops=None, final_resid=h["final_normr"] if "final_normr" in h else None)


def cpu_sample_dims(args, dims):
    """What the CPU arm runs: the SAME mesh as the GPU arm when the reference can hold it (int32 ordinals, host memory:
    anything up to 256x256x512), else the per-GPU share nx^3 of it -- said so in `sample`."""
    if args.cpu_nx:
        return [args.cpu_nx] * 3
    nodes = (dims[0] + 1) * (dims[1] + 1) * (dims[2] + 1)
    return list(dims) if nodes <= 34_500_000 else [min(args.nx, 256)] * 3


def cpu_baseline_obj(r, dims):
    same = list(r["dims"]) == list(dims)
    what = "reference OpenMP build (openmp/src unmodified) on all host threads" if r["kind"] == "reference" else "plain-C oracle port, 1 thread"
    return {"value": r["cg_mflops"] / 1e3, "unit": "GFLOP/s", "cores": r["cores"], "kind": r["kind"], "same_config": same,
            "nx": r["dims"][0], "ny": r["dims"][1], "nz": r["dims"][2], "cg_iterations": int(r["iters"]), "cg_time_s": r["cg_time"],
            "matvec_gflops": (r["matvec_mflops"] / 1e3) if r.get("matvec_mflops") else None,
            "struct_gen_time_s": r.get("struct_time"), "fe_assembly_time_s": r.get("assembly_time"), "final_resid": r.get("final_resid"), "wall_s": r["wall_s"],
            "sample": "%s, ONE run of nx=%d ny=%d nz=%d (%s): its own full 200-iteration cg_solve timing (%d iterations, %.1f s), %.1f s wall in all" % (
                what, r["dims"][0], r["dims"][1], r["dims"][2], "the same mesh as the GPU arm" if same else "a bounded sample of the %dx%dx%d workload" % tuple(dims),
                int(r["iters"]), r["cg_time"], r["wall_s"])}


def reference_arm(args, rank):
    """--impl reference: the reference's CPU implementation, timed on the host cores (rank 0 only).  The executable runs ONCE
    per invocation: its 200-iteration cg_solve is the sample, --steps/--warmup are echoed but add nothing (repeating a
    40-second CPU solve 25 times would not make the number better)."""
    if rank != 0:
        return 0
    dims = [args.nx] * 3 if args.strong else weak_dims(args.nx, args.gpus)
    r = run_reference_cpu(cpu_sample_dims(args, dims)) or run_oracle_port_cpu(96)
    cb = cpu_baseline_obj(r, dims)
    gf, ms = cb["value"], r["cg_time"] * 1e3
    line = {"impl": "reference", "metric": "cg_gflops", "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "miniFE %dx%dx%d hex8 Poisson (%s), assembled on device, 200-iteration unpreconditioned CG" % (
                           dims[0], dims[1], dims[2], "strong scaling: fixed global mesh" if args.strong else "%d^3 elements per GPU" % args.nx),
                       "sample": cb["sample"], "parallelism": "reference OpenMP build on %d host threads (no GPU)" % r["cores"], "flop_model": "ref/src/driver.hpp:322-341",
                       "runs": 1},
            "cpu_baseline": cb,
            "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def golden_for(dims, world):
    """The unmodified reference's record of this very configuration, if tests/golden holds one: per-iteration histories and
    CRC-32s of the solution (tests/golden/make_big_golden.py: ref/src over the MPI shim at np = world; history_<n>.npz: serial)."""
    g = os.path.join(ROOT, "tests", "golden")
    fn = os.path.join(g, "big_%dx%dx%d_np%d.npz" % (dims[0], dims[1], dims[2], world))
    if os.path.exists(fn):
        return fn, np.load(fn)
    if world == 1 and dims[0] == dims[1] == dims[2]:
        fn = os.path.join(g, "history_%d.npz" % dims[0])
        if os.path.exists(fn):
            return fn, np.load(fn)
    # sizes the reference cannot hold in the authoring container: the record of the oracle (tests/golden/make_oracle_golden.py),
    # which equals the reference's record bit for bit wherever both exist
    fn = os.path.join(g, "oracle_%dx%dx%d_np%d.npz" % (dims[0], dims[1], dims[2], world))
    if os.path.exists(fn):
        return fn, np.load(fn)
    return None, None


def check_parity(mesh, mf, L, dims, rank, world, dist, torch, iters, hist):
    """Compare the solve just timed (free-running, no reference in the loop) with the golden record of the reference at the
    same np: iteration count, the whole normr history and the CRC-32 of this rank's solution, all bit for bit.  A mismatch
    ends the bench with an error; without a golden the solve is at least verified against the analytic solution."""
    import zlib
    fn, g = golden_for(dims, world)
    if g is None:
        # No reference record at this size (512^3 needs > 62 GB on the CPU side).  Size-independent checks instead: the
        # order-reproducing dot must return the bits of the literal serial loop on this very solution vector (a one-thread
        # device loop, MFE_DOT_SERIAL), on every rank; verify_solution's error is reported (after 200 iterations the solve is
        # far from converged at this size -- so is the reference's -- hence no verdict).
        x = np.zeros(int(L.mfe_nrows(mesh._h)))
        L.mfe_export_vector(mesh._h, b"x", x)
        d_exact = mf.dot(x, x, ctx=mesh)
        mesh.set_dot_mode(1)
        d_serial = mf.dot(x, x, ctx=mesh)
        mesh.set_dot_mode(0)
        nbad = int(np.float64(d_exact).view(np.int64) != np.float64(d_serial).view(np.int64))
        if dist:
            t = torch.tensor([nbad], device="cuda")
            dist.all_reduce(t)
            nbad = int(t)
        if nbad:
            raise SystemExit("PARITY FAILURE: exact dot %.17g != serial loop %.17g on the solution vector" % (d_exact, d_serial))
        rc, gmax, _ = mf.verify_solution(mesh, None, 0.06, False, info=True)
        return ("no golden for %dx%dx%d np=%d (the reference needs more host memory than this container has at that size); x.x by the "
                "order-reproducing dot == the literal serial loop bit for bit on every rank (%.17g); verify_solution max error %.3g after %d iterations" % (
                    dims[0], dims[1], dims[2], world, d_exact, gmax, iters))
    want = g["normr"]
    bad = []
    if iters != len(want) - 1:
        bad.append("iterations %d vs %d" % (iters, len(want) - 1))
    k = min(iters, len(want) - 1)
    if not np.array_equal(hist[:k + 1].view(np.int64), want[:k + 1].view(np.int64)):
        i = int(np.argmax(hist[:k + 1].view(np.int64) != want[:k + 1].view(np.int64)))
        bad.append("normr history differs from iteration %d (%.17g vs %.17g)" % (i, hist[i], want[i]))
    x = np.zeros(int(L.mfe_nrows(mesh._h)))
    L.mfe_export_vector(mesh._h, b"x", x)
    key = "crc_x_%d" % rank if ("crc_x_%d" % rank) in g.files else ("crc_x" if "crc_x" in g.files else None)
    if key is not None and np.uint32(zlib.crc32(x.tobytes())) != np.uint32(g[key]):
        bad.append("CRC-32 of the solution on rank %d differs" % rank)
    nbad = len(bad)
    if dist:
        t = torch.tensor([nbad], device="cuda")
        dist.all_reduce(t)
        nbad = int(t)
    if nbad:
        raise SystemExit("PARITY FAILURE against %s: %s" % (os.path.basename(fn), "; ".join(bad) or "on another rank"))
    if os.path.basename(fn).startswith("oracle_"):
        return ("oracle golden np=%d ok (%s: %d residual norms and the solution CRC on every rank bit-equal to the CPU oracle's record; the reference "
                "itself does not fit the authoring container at this size, the oracle equals it bit for bit at every size where both run)" % (world, os.path.basename(fn), k + 1))
    return "golden np=%d ok (%s: %d residual norms and the solution CRC on every rank bit-equal to the unmodified reference)" % (world, os.path.basename(fn), k + 1)


def traffic_for(dims, world):
    """dram__bytes_read + dram__bytes_write of one SpMV launch from an `ncu --set full` capture of THIS configuration
    (profiles/matvec_traffic_<nx>x<ny>x<nz>_np<N>.json), else None."""
    fn = os.path.join(ROOT, "profiles", "matvec_traffic_%dx%dx%d_np%d.json" % (dims[0], dims[1], dims[2], world))
    if os.path.exists(fn):
        with open(fn) as f:
            return json.load(f).get("dram_bytes_per_launch")
    return None


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--nx", type=int, default=256, help="elements per GPU per dimension (weak scaling), or of the whole mesh with --strong")
    ap.add_argument("--strong", action="store_true", help="strong scaling: the global mesh is nx^3 whatever --gpus is (BASELINE.json configs[2]: --nx 400 --strong)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-nx", type=int, default=0, help="force the CPU baseline onto an nx^3 cube (default 0: the GPU arm's own mesh when the reference can hold it)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dot-mode", type=int, default=0)
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"], help="N > 1: halo exchange and dot all-reduce over NVLink peer stores "
                    "(default; falls back to NCCL when CUDA IPC is unavailable) or over NCCL send/recv + allgather")
    ap.add_argument("--no-overlap", action="store_true", help="peer mode: do not overlap the halo exchange with the interior SpMV tiles")
    ap.add_argument("--overlap", type=int, default=1, help="peer mode: 1 = one fused SpMV launch (pack, interior tiles, wait, boundary tiles), 2 = the same as four launches, 0 = none")
    ap.add_argument("--x-overlap", type=int, default=1, help="1: x += alpha*p runs on a second, low-priority stream beside the dot kernels; 0: fused into the r update")
    ap.add_argument("--op-timing", type=int, default=2, help="CUDA events inside the timed solves: 2 = around every SpMV launch (roofline), 1 = every op class, 0 = none")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
        args.gpus = world

    import torch
    import minife_b200 as mf
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    dims = [args.nx] * 3 if args.strong else weak_dims(args.nx, world)
    gbox = mf.Box([0, dims[0], 0, dims[1], 0, dims[2]])
    boxes = mf.box_partition(0, world, 2, gbox)
    mesh = mf.simple_mesh_description(gbox, boxes[rank], rank=rank, nranks=world, all_boxes=boxes, device=local_rank)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(mf.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        mesh.comm_init(bytes(idt.cpu().numpy().tobytes()))
    mesh.set_dot_mode(args.dot_mode)
    mesh.set_op_timing(args.op_timing)
    mesh.set_option("comm_mode", 1 if args.comm == "peer" else 0)
    mesh.set_option("overlap", 0 if args.no_overlap else args.overlap)
    mesh.set_option("x_overlap", args.x_overlap)

    A = mf.CSRMatrix()
    t_setup = time.time()
    mf.generate_matrix_structure(mesh, A)
    mf.assemble_FE_data(mesh, A, None)
    mf.impose_dirichlet(0.0, A, None, bc_rows=0)
    mf.impose_dirichlet(1.0, A, None, bc_rows=1)
    mf.make_local_matrix(A)
    t_setup = time.time() - t_setup
    nrows, nnz = A.num_rows(), A.num_nonzeros()
    if world > 1:
        t = torch.tensor([nrows, nnz], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        nrows_g, nnz_g = int(t[0]), int(t[1])
    else:
        nrows_g, nnz_g = nrows, nnz

    L = mf.lib()
    xpin = torch.zeros(nrows, dtype=torch.float64).pin_memory()
    xnp = xpin.numpy()
    zero = np.zeros(nrows)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    hist_buf = np.full(MAX_ITERS + 1, np.nan)

    def solve(host_buffers, want_hist=False):
        """one step; returns (iters, normr, cg_times)"""
        if host_buffers:
            xnp[:] = 0.0
        else:
            L.mfe_import_vector(mesh._h, b"x", zero)       # x0 = 0, outside the timed region
        barrier()
        t0 = time.perf_counter()
        it, nr = C.c_int(0), C.c_double(0)
        times = np.zeros(5)
        rc = L.mfe_cg_solve(mesh._h, xnp.ctypes.data_as(C.c_void_p) if host_buffers else None, MAX_ITERS, EPS, C.byref(it), C.byref(nr),
                            times.ctypes.data_as(C.c_void_p), hist_buf.ctypes.data_as(C.c_void_p) if want_hist else None, None, None, None)
        if rc not in (0, 2):
            raise SystemExit("cg_solve failed: %s" % L.mfe_last_error().decode())
        barrier()
        wall = time.perf_counter() - t0
        return it.value, nr.value, times, wall

    op_s = []

    def timed(host_buffers, steps):
        dev_s, wall_s, mv_s, its, last = [], [], [], [], None
        for _ in range(steps):
            it, nr, times, wall = solve(host_buffers)
            # device time of the solve: CUDA events on the library stream (cg_times[TOTAL]); max over ranks
            tt = torch.tensor([times[4], wall, times[2], times[0], times[1]], dtype=torch.float64, device="cuda")
            if dist:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dev_s.append(float(tt[0])); wall_s.append(float(tt[1])); mv_s.append(float(tt[2])); its.append(it); last = nr
            op_s.append((float(tt[3]), float(tt[4])))
        return dev_s, wall_s, mv_s, its, last

    for _ in range(args.warmup):
        solve(False)
    # per-op-class breakdown (reference YAML rows WAXPY / DOT / MATVEC): one extra, untimed solve with events around every op
    mesh.set_op_timing(1)
    it1, _, t1, _ = solve(False, want_hist=True)
    parity = check_parity(mesh, mf, L, dims, rank, world, dist, torch, it1, hist_buf)
    cg_times = {"WAXPY_s": float(t1[0]), "DOT_s": float(t1[1]), "MATVEC_s": float(t1[2]), "TOTAL_s": float(t1[4]),
                "note": "one untimed solve with CUDA events around every op (costs ~3 %); the timed solves record events around the SpMV only"}
    mesh.set_op_timing(args.op_timing)
    launches0 = mesh.launch_count()
    L.mfe_dot_stats(mesh._h, None, 0, 1)             # reset the dot diagnostics
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_s, wall_s, mv_s, its, normr = timed(False, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = mesh.launch_count() - launches0
    # exact-dot diagnostics of the timed solves (mfe_dot_stats): walk time and, multi-rank, the time spent in the all-reduce
    st = np.zeros(16, np.uint64)
    L.mfe_dot_stats(mesh._h, st.ctypes.data_as(C.c_void_p), 16, 0)
    ndots = max(int(st[0]), 1)
    dstat = torch.tensor([float(st[5]) / ndots / 1e3, float(st[9]) / ndots / 1e3, float(st[10]) / 1e3, float(st[1]) / ndots, float(st[4]), float(st[8]) / ndots,
                          float(t1[0]), float(t1[1]), float(t1[2])], dtype=torch.float64, device="cuda")
    if dist:
        allst = [torch.zeros_like(dstat) for _ in range(world)]
        dist.all_gather(allst, dstat)
    else:
        allst = [dstat]
    dot_diag = {"per_rank_walk_us": [round(float(t[0]), 2) for t in allst], "per_rank_allreduce_wait_us_avg": [round(float(t[1]), 2) for t in allst],
                "per_rank_allreduce_wait_us_max": [round(float(t[2]), 1) for t in allst], "per_rank_dirty_leaves_avg": [round(float(t[3]), 1) for t in allst],
                "slow_gaps": [int(t[4]) for t in allst], "per_rank_slow_pieces_per_dot": [round(float(t[5]), 1) for t in allst],
                "per_rank_op_timed_solve_s": {"WAXPY": [round(float(t[6]), 5) for t in allst], "DOT": [round(float(t[7]), 5) for t in allst],
                                              "MATVEC": [round(float(t[8]), 5) for t in allst]}}
    e_dev, e_wall, _, e_its, _ = timed(True, max(1, min(args.steps, 3)))

    iters = its[-1]
    flops = flop_model(dims, nnz_g, iters)
    byts = algorithmic_bytes(nrows_g, nnz_g, iters)
    t_step = float(np.mean(dev_s))
    value = flops / t_step / 1e9
    hbm_gbs = byts["total"] / t_step / 1e9
    e2e_t = float(np.mean(e_wall))
    e2e_value = flop_model(dims, nnz_g, e_its[-1]) / e2e_t / 1e9

    # roofline of the dominant kernel (SpMV): algorithmic bytes of one launch on THIS rank / its mean duration,
    # from the CUDA events the library records around every matvec launch inside the timed solves
    peak, peak_src = measured_peaks()
    mv_bytes_rank = 12.0 * nnz + 4.0 * (nrows + 1) + 16.0 * nrows
    mv_ms = float(np.mean(mv_s)) / (iters + 1) * 1e3
    if mv_ms <= 0.0:                                   # --op-timing 0: isolated launches on the same device-resident data
        mv_ms = mesh.time_kernel("matvec_dot", 20)
    achieved = mv_bytes_rank / (mv_ms * 1e-3) / 1e9
    traffic = traffic_for(dims, world)
    # assembly: algorithmic output 12*nnz + 8*N bytes written once; FP64 work ~6 kflop per element
    asm_ms = mesh.phase_ms("assembly")
    nelem_rank = float(nnz) / 27.0                      # ~ elements per rank (interior rows hold 27 entries)
    asm_bytes = 12.0 * nnz + 8.0 * nrows
    assembly = {"kernel": "k_assemble (elements + row-owner gather)", "bound": "hbm", "algorithmic_bytes": asm_bytes, "ms": asm_ms,
                "achieved": asm_bytes / (asm_ms * 1e-3) / 1e9 if asm_ms > 0 else None, "peak": peak, "unit": "GB/s",
                "frac": (asm_bytes / (asm_ms * 1e-3) / 1e9 / peak) if asm_ms > 0 else None,
                "fp64_gflop": 6.0e-6 * nelem_rank, "structure_ms": mesh.phase_ms("structure"),
                "dirichlet_ms": mesh.phase_ms("dirichlet0") + mesh.phase_ms("dirichlet1"), "make_local_ms": mesh.phase_ms("make_local")}
    roofline = {"kernel": "k_matvec (CSR SpMV, TMA-staged)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": mv_bytes_rank, "ms_per_launch": mv_ms,
                "cg_iteration": {"achieved": hbm_gbs / world, "frac": hbm_gbs / world / peak, "frac_of_8TBs": hbm_gbs / world / 8000.0}}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = run_reference_cpu(cpu_sample_dims(args, dims)) or run_oracle_port_cpu(96)
            cpu = cpu_baseline_obj(r, dims)
        line = {
            "metric": "cg_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "miniFE %dx%dx%d hex8 Poisson (%s), assembled on device, 200-iteration unpreconditioned CG" % (
                           dims[0], dims[1], dims[2], "strong scaling: fixed global mesh" if args.strong else "%d^3 elements per GPU" % args.nx),
                       "rows": nrows_g, "nnz": nnz_g, "cg_iterations": iters, "final_resid": normr, "parallelism": "box_partition x%d, 1 rank per GPU" % world,
                       "comm": (None if world == 1 else ("NCCL send/recv + allgather" if args.comm == "nccl" else
                                "NVLink peer stores (halo + dot all-reduce)%s" % ("" if args.no_overlap else ", halo overlapped with the interior SpMV tiles"))),
                       "l2": "inputs larger than L2 (%.1f GB per iteration per GPU vs 126 MB)" % ((12.0 * nnz + 116.0 * nrows) / 1e9), "dot": ["exact (order-reproducing)", "serial", "tree"][args.dot_mode],
                       "flop_model": "ref/src/driver.hpp:322-341"},
            "hbm_gbs": hbm_gbs, "hbm_gbs_per_gpu": hbm_gbs / world, "pct_of_8TBs_per_gpu": 100.0 * hbm_gbs / world / 8000.0,
            "ms_per_cg_iteration": t_step * 1e3 / max(iters, 1), "cg_times": cg_times,
            "setup_ms": {"structure": mesh.phase_ms("structure"), "assembly": mesh.phase_ms("assembly"), "assembly_elem": mesh.phase_ms("assembly_elem"),
                         "assembly_gather": mesh.phase_ms("assembly_gather"), "dirichlet0": mesh.phase_ms("dirichlet0"), "dirichlet1": mesh.phase_ms("dirichlet1"),
                         "make_local": mesh.phase_ms("make_local"), "wall_all": t_setup * 1e3},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "GFLOP/s", "h2d_bytes_per_step": 8 * nrows, "d2h_bytes_per_step": 8 * nrows + 4 * 8 * (MAX_ITERS + 1) + 128,
                    "ms_per_step": e2e_t * 1e3, "what": "mfe_cg_solve with host x (pinned): H2D x0 + 200-iteration solve + D2H x and histories, wall clock, max over ranks"},
            "gpu_launches": int(launches),
            "dot_diag": dot_diag,
            "roofline": roofline,
            "assembly": assembly,
            "parity": parity,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    mesh.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
