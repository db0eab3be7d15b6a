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
PDL_DEFAULT = 0


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
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


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
def run_reference_cpu(nx, threads=None):
    exe = os.path.join(ROOT, "oracle", "_ref", "miniFE_ref_openmp.x")
    kind = "reference"
    threads = threads or os.cpu_count() or 1
    if not os.path.exists(exe):
        return None
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="spread")
    with tempfile.TemporaryDirectory() as d:
        t0 = time.time()
        out = subprocess.run([exe, "nx=%d" % nx, "ny=%d" % nx, "nz=%d" % nx], cwd=d, env=env, capture_output=True, text=True, timeout=1800)
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
    return dict(kind=kind, cores=threads, wall_s=wall, iters=grab("Iterations"), cg_mflops=grab("Total CG Mflops"), cg_time=grab("Total CG Time"),
                matvec_mflops=grab("MATVEC Mflops"), struct_time=grab("Mat-struc-gen Time"), assembly_time=grab("FE assembly Time"),
                final_resid=grab("Final Resid Norm"))


def run_oracle_port_cpu(nx):
    """Fallback when oracle/_ref is absent: the plain-C restatement (1 thread)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    o = Oracle(nx, nx, nx).setup()
    nnz = o.nnz()
    t0 = time.time()
    h = o.cg_solve(MAX_ITERS, EPS)
    dt = time.time() - t0
    return dict(kind="port", cores=1, wall_s=dt, iters=h["num_iters"], cg_mflops=1e-6 * flop_model([nx] * 3, nnz, h["num_iters"]) / dt, cg_time=dt)


def reference_arm(args, rank):
    """--impl reference: the reference's CPU implementation, timed on the host cores (rank 0 only)."""
    if rank != 0:
        return 0
    sample = args.cpu_nx
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        r = run_reference_cpu(sample) or run_oracle_port_cpu(min(sample, 96))
        if i >= args.warmup:
            vals.append(r)
        last = r
    gf = float(np.mean([v["cg_mflops"] for v in vals])) / 1e3
    ms = float(np.mean([v["cg_time"] for v in vals])) * 1e3
    dims = [args.nx] * 3 if args.strong else weak_dims(args.nx, args.gpus)
    desc = "reference %s build, nx=ny=nz=%d (bounded sample of the %dx%dx%d workload), %d host threads, full 200-iteration cg_solve per step" % (
        "OpenMP (openmp/src, unmodified)" if last["kind"] == "reference" else "C port", sample, dims[0], dims[1], dims[2], last["cores"])
    line = {"impl": "reference", "metric": "cg_gflops", "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "miniFE %dx%dx%d hex8 Poisson (%s), assembled on device, 200-iteration unpreconditioned CG" % (
                           dims[0], dims[1], dims[2], "strong scaling: fixed global mesh" if args.strong else "%d^3 elements per GPU" % args.nx),
                       "sample": desc, "parallelism": "reference OpenMP build on %d host threads (no GPU)" % last["cores"], "flop_model": "ref/src/driver.hpp:322-341"},
            "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": last["cores"], "kind": last["kind"], "sample": desc},
            "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--nx", type=int, default=256, help="elements per GPU per dimension (weak scaling), or of the whole mesh with --strong")
    ap.add_argument("--strong", action="store_true", help="strong scaling: the global mesh is nx^3 whatever --gpus is (BASELINE.json configs[2]: --nx 400 --strong)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-nx", type=int, default=128, help="size of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dot-mode", type=int, default=0)
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"], help="N > 1: halo exchange and dot all-reduce over NVLink peer stores "
                    "(default; falls back to NCCL when CUDA IPC is unavailable) or over NCCL send/recv + allgather")
    ap.add_argument("--pdl", type=int, default=PDL_DEFAULT, help="programmatic dependent launch of the CG-loop kernels (1/0)")
    ap.add_argument("--no-overlap", action="store_true", help="peer mode: do not overlap the halo exchange with the interior SpMV tiles")
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
    mesh.set_option("overlap", 0 if args.no_overlap else 1)
    mesh.set_option("pdl", args.pdl)

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

    def solve(host_buffers):
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
                            times.ctypes.data_as(C.c_void_p), None, None, None, None)
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
    _, _, t1, _ = solve(False)
    cg_times = {"WAXPY_s": float(t1[0]), "DOT_s": float(t1[1]), "MATVEC_s": float(t1[2]), "TOTAL_s": float(t1[4]),
                "note": "one untimed solve with CUDA events around every op (costs ~3 %); the timed solves record events around the SpMV only"}
    mesh.set_op_timing(args.op_timing)
    launches0 = mesh.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_s, wall_s, mv_s, its, normr = timed(False, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = mesh.launch_count() - launches0
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
    traffic = None
    tf = os.path.join(ROOT, "profiles", "r01_matvec_traffic.json")
    if os.path.exists(tf):
        with open(tf) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    roofline = {"kernel": "k_matvec (CSR SpMV, TMA-staged)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": mv_bytes_rank, "ms_per_launch": mv_ms,
                "cg_iteration": {"achieved": hbm_gbs / world, "frac": hbm_gbs / world / peak, "frac_of_8TBs": hbm_gbs / world / 8000.0}}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = run_reference_cpu(args.cpu_nx) or run_oracle_port_cpu(min(args.cpu_nx, 96))
            cpu = {"value": r["cg_mflops"] / 1e3, "unit": "GFLOP/s", "cores": r["cores"], "kind": r["kind"],
                   "sample": "%s, nx=ny=nz=%d, full 200-iteration cg_solve (%d iterations run), %.1f s wall; its FE assembly %.3g s, structure %.3g s" % (
                       "reference OpenMP build (openmp/src unmodified) on all host threads" if r["kind"] == "reference" else "plain-C oracle port, 1 thread",
                       args.cpu_nx if r["kind"] == "reference" else min(args.cpu_nx, 96), int(r["iters"]), r["wall_s"],
                       r.get("assembly_time") or float("nan"), r.get("struct_time") or float("nan"))}
        line = {
            "metric": "cg_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "miniFE %dx%dx%d hex8 Poisson (%s), assembled on device, 200-iteration unpreconditioned CG" % (
                           dims[0], dims[1], dims[2], "strong scaling: fixed global mesh" if args.strong else "%d^3 elements per GPU" % args.nx),
                       "rows": nrows_g, "nnz": nnz_g, "cg_iterations": iters, "final_resid": normr, "parallelism": "box_partition x%d, 1 rank per GPU" % world,
                       "comm": (None if world == 1 else ("NCCL send/recv + allgather" if args.comm == "nccl" else
                                "NVLink peer stores (halo + dot all-reduce)%s" % ("" if args.no_overlap else ", halo overlapped with the interior SpMV tiles"))),
                       "l2": "inputs larger than L2 (%.1f GB per iteration per GPU vs 126 MB)" % ((12.0 * nnz + 116.0 * nrows) / 1e9), "dot": ["exact (order-reproducing)", "serial", "tree"][args.dot_mode],
                       "flop_model": "ref/src/driver.hpp:322-341", "pdl": bool(args.pdl)},
            "hbm_gbs": hbm_gbs, "hbm_gbs_per_gpu": hbm_gbs / world, "pct_of_8TBs_per_gpu": 100.0 * hbm_gbs / world / 8000.0,
            "ms_per_cg_iteration": t_step * 1e3 / max(iters, 1), "cg_times": cg_times,
            "setup_ms": {"structure": mesh.phase_ms("structure"), "assembly": mesh.phase_ms("assembly"), "assembly_elem": mesh.phase_ms("assembly_elem"),
                         "assembly_gather": mesh.phase_ms("assembly_gather"), "dirichlet0": mesh.phase_ms("dirichlet0"), "dirichlet1": mesh.phase_ms("dirichlet1"),
                         "make_local": mesh.phase_ms("make_local"), "wall_all": t_setup * 1e3},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "GFLOP/s", "h2d_bytes_per_step": 8 * nrows, "d2h_bytes_per_step": 8 * nrows + 4 * 8 * (MAX_ITERS + 1) + 128,
                    "ms_per_step": e2e_t * 1e3, "what": "mfe_cg_solve with host x (pinned): H2D x0 + 200-iteration solve + D2H x and histories, wall clock, max over ranks"},
            "gpu_launches": int(launches),
            "roofline": roofline,
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
