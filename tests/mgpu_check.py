#!/usr/bin/env python3
"""Multi-GPU parity check, one rank per GPU (run under torch.distributed.run; needs N GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/mgpu_check.py [--sizes 8x8x8,20x30x40,64x64x64,tests/golden/mpi_12x12x12_np8_imb30.npz]

Every rank builds its share with minife_b200 (halo exchange + rank-ordered dot all-reduce, over NVLink peer stores with
and without SpMV overlap and over NCCL) and compares with the
CPU oracle run at the SAME number of ranks: CSR, index maps, exchange_externals, matvec, CG history and solution,
all bit-exact.  Exit code 0 = all ranks agree."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import minife_b200 as mf
    from oracle_lib import Oracle, EPS
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="2x2x2,8x8x8,5x3x7,20x30x40,64x64x64")
    ap.add_argument("--modes", default="peer+overlap,peer+overlap2,peer,nccl", help="communication modes to check: peer+overlap (NVLink peer "
                    "stores, ONE fused SpMV launch: pack, interior tiles, flag wait, boundary tiles), peer+overlap2 (the same as four "
                    "launches), peer (no overlap), nccl (send/recv + allgather)")
    ap.add_argument("--big", default="", help="comma-separated tests/golden/big_*.npz files (np must equal the world size): the run is "
                    "compared with the UNMODIFIED reference's record at that size -- CRC-32 of every array after every phase, the index "
                    "maps, the four per-iteration histories and the solution -- instead of with the in-process oracle")
    ap.add_argument("--big-modes", default="peer+overlap")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    failures = []
    oracles = {}

    def bits(a):
        return np.ascontiguousarray(a, np.float64).view(np.int64)

    def check(cond, what):
        if not cond:
            failures.append(what)
            print("[rank %d] MISMATCH: %s" % (rank, what), flush=True)

    for size, mode in [(s_, m_) for s_ in args.sizes.split(",") for m_ in args.modes.split(",")]:
        gold_boxes = None
        if size.endswith(".npz"):      # a golden file of the reference run with add_imbalance: use ITS boxes (np must match)
            g = np.load(size)
            if int(g["np"]) != world:
                continue
            n = tuple(int(v) for v in g["n"])
            gold_boxes = np.stack([g["box_%d" % r] for r in range(world)])
            size = os.path.basename(size)[:-4]
        else:
            n = tuple(int(v) for v in size.split("x"))
        if min(n) * 1 < 1 or (n[0] * n[1] * n[2]) < world:
            continue
        gbox = mf.Box([0, n[0], 0, n[1], 0, n[2]])
        boxes = mf.box_partition(0, world, 2, gbox) if gold_boxes is None else [mf.Box(b) for b in gold_boxes]
        if any((b.ranges[1] - b.ranges[0]) * (b.ranges[3] - b.ranges[2]) * (b.ranges[5] - b.ranges[4]) == 0 for b in boxes):
            continue      # the reference itself refuses ranks without elements (main.cpp:116-130)
        mesh = mf.simple_mesh_description(gbox, boxes[rank], rank=rank, nranks=world, all_boxes=boxes, device=local)
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(mf.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        mesh.comm_init(bytes(idt.cpu().numpy().tobytes()))
        mesh.set_option("comm_mode", 0 if mode == "nccl" else 1)
        mesh.set_option("overlap", {"peer+overlap": 1, "peer+overlap2": 2}.get(mode, 0))
        size = "%s[%s]" % (size, mode)
        A = mf.CSRMatrix()
        mf.generate_matrix_structure(mesh, A)
        b = mf.Vector(int(A.rows[0]), A.num_rows()); x = mf.Vector(int(A.rows[0]), A.num_rows())
        mf.assemble_FE_data(mesh, A, b)
        mf.impose_dirichlet(0.0, A, b, bc_rows=0); mf.impose_dirichlet(1.0, A, b, bc_rows=1)
        mf.make_local_matrix(A)
        okey = (n, None if gold_boxes is None else gold_boxes.tobytes())
        if okey not in oracles:
            oracles.clear()
            oracles[okey] = Oracle(*n, world, boxes=gold_boxes).setup()
            oracles[okey]._cg = oracles[okey].cg_solve(200, EPS)
        o = oracles[okey]
        rows, ro, cols, coefs = o.csr(rank)
        check(np.array_equal(A.row_offsets, ro) and np.array_equal(A.packed_cols, cols), "%s CSR structure" % size)
        check(np.array_equal(bits(A.packed_coefs), bits(coefs)) and np.array_equal(bits(b.coefs), bits(o.b(rank))), "%s A/b values" % size)
        for m in ("external_index", "external_local_index", "elements_to_send", "neighbors", "recv_length", "send_length"):
            check(np.array_equal(getattr(A, m), o.get_map(m, rank)), "%s %s" % (size, m))
        # exchange_externals + matvec on a global pattern: x_global[row] = f(row)
        ncols = A.num_cols
        first = int(A.rows[0])
        f = lambda r: np.sin(0.37 * r) + 1e-3 * r
        xv = np.zeros(ncols); xv[:A.num_rows()] = f(first + np.arange(A.num_rows()))
        want_ext = f(o.get_map("external_index", rank).astype(np.float64))
        order = o.get_map("external_local_index", rank) - A.num_rows()
        want_tail = np.zeros(ncols - A.num_rows()); want_tail[order] = want_ext
        xe = xv.copy()
        mf.exchange_externals(A, xe)
        check(np.array_equal(bits(xe[A.num_rows():]), bits(want_tail)), "%s exchange_externals" % size)
        y = np.zeros(A.num_rows())
        mf.matvec(A, xv, y)
        xo = xv.copy(); xo[A.num_rows():] = want_tail
        check(np.array_equal(bits(y), bits(o.matvec(xo, rank))), "%s matvec" % size)
        # free-running CG against the oracle at np = world
        want = o._cg
        iters, normr, times, h = mf.cg_solve(A, b, x, mf.matvec_std(), 200, EPS, history=True)
        check(iters == want["num_iters"], "%s iterations %d vs %d" % (size, iters, want["num_iters"]))
        k = min(iters, want["num_iters"])
        check(np.array_equal(bits(h["normr"][:k + 1]), bits(want["normr"][:k + 1])), "%s residual history" % size)
        check(np.array_equal(bits(x.coefs), bits(o.x(rank))), "%s solution" % size)
        # verify_solution on the device-resident x, MAX over ranks (ref/fem/verify_solution.hpp:148-154)
        whole = n[0] * n[1] * n[2] <= 20 * 30 * 40
        vrc, vmax, _ = mf.verify_solution(mesh, None, 0.06, whole, info=True)
        wrc, wmax, _ = o.verify_solution(0.06, whole, info=True)
        check(vrc == wrc and abs(vmax - wmax) <= 1e-12 * max(wmax, 1e-30) + 1e-15, "%s verify_solution %r vs %r" % (size, (vrc, vmax), (wrc, wmax)))
        # time_kernels (ref/src/time_kernels.hpp:97-124): back-to-back matvecs with no dot between them -- consecutive halo
        # epochs with nothing that re-synchronises the ranks (a neighbour may run one epoch ahead)
        xk = mf.Vector(first, A.num_rows()); xk.coefs[:] = f(first + np.arange(A.num_rows()))
        nk = 6
        xdotp, _ = mf.time_kernels(A, b, xk, None, nk)
        wx, wxdotp = o.time_kernels(f, nk)
        check(np.array_equal(bits(xk.coefs), bits(wx[rank])) and bits(np.array([xdotp]))[0] == bits(np.array([wxdotp]))[0], "%s time_kernels" % size)
        # mv_overlap_comm_comp=1 (driver.hpp:278-282): rearrange_matrix_local_external + matvec_overlap, against what the
        # UNMODIFIED reference computes on that branch (tests/golden/mpiov_*.npz)
        gov = os.path.join(ROOT, "tests", "golden", "mpiov_%dx%dx%d_np%d.npz" % (n + (world,)))
        if gold_boxes is None and os.path.exists(gov):
            gv = np.load(gov)
            mf.rearrange_matrix_local_external(A)
            x2 = mf.Vector(first, A.num_rows())
            it2, _, _, h2 = mf.cg_solve(A, b, x2, mf.matvec_overlap(), 200, EPS, history=True)
            check(it2 == len(gv["normr"]) - 1 and np.array_equal(bits(h2["normr"]), bits(gv["normr"])), "%s overlap-order residual history" % size)
            check(np.array_equal(bits(x2.coefs), bits(gv["x_%d" % rank])), "%s overlap-order solution" % size)
        if rank == 0:
            rel = np.abs(h["normr"][:k + 1] - want["normr"][:k + 1]) / np.maximum(want["normr"][:k + 1], 1e-300)
            print("size %s np=%d: iters %d (oracle %d), final resid %.6g, max rel diff of history %.3g" % (size, world, iters, want["num_iters"], normr, rel.max()), flush=True)
        mesh.close()

    # ---- large sizes: against the unmodified reference's golden record (tests/golden/make_big_golden.py) ----
    import zlib
    crc = lambda a: np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))
    for path, mode in [(p_, m_) for p_ in args.big.split(",") if p_ for m_ in args.big_modes.split(",")]:
        g = np.load(path)
        if int(g["np"]) != world:
            continue
        n = tuple(int(v) for v in g["n"])
        tag = "%s[%s]" % (os.path.basename(path)[:-4], mode)
        gbox = mf.Box([0, n[0], 0, n[1], 0, n[2]])
        boxes = mf.box_partition(0, world, 2, gbox)
        check(np.array_equal(boxes[rank].ranges, g["box_%d" % rank]), "%s box" % tag)
        mesh = mf.simple_mesh_description(gbox, boxes[rank], rank=rank, nranks=world, all_boxes=boxes, device=local)
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(mf.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        mesh.comm_init(bytes(idt.cpu().numpy().tobytes()))
        mesh.set_option("comm_mode", 0 if mode == "nccl" else 1)
        mesh.set_option("overlap", {"peer+overlap": 1, "peer+overlap2": 2}.get(mode, 0))
        A = mf.CSRMatrix()
        mf.generate_matrix_structure(mesh, A)
        nrows, nnz = A.num_rows(), A.num_nonzeros()
        check((nrows, nnz) == tuple(int(v) for v in g["sizes_%d" % rank]), "%s sizes" % tag)
        b = mf.Vector(int(A.rows[0]), nrows); x = mf.Vector(int(A.rows[0]), nrows)
        # (records of the oracle, tests/golden/oracle_*.npz, carry no CRCs of cols / coefs / row_offsets: `have`)
        have = lambda k: ("%s_%d" % (k, rank)) in g.files
        want = lambda k: g["%s_%d" % (k, rank)]
        check((not have("crc_row_offsets") or crc(A.row_offsets) == want("crc_row_offsets")) and
              (not have("crc_cols_global") or crc(A.packed_cols) == want("crc_cols_global")), "%s structure CRC" % tag)
        mf.assemble_FE_data(mesh, A, b)
        check((not have("crc_coefs_prebc") or crc(A.packed_coefs) == want("crc_coefs_prebc")) and crc(b.coefs) == want("crc_b_prebc"), "%s assembled A/b CRC" % tag)
        mf.impose_dirichlet(0.0, A, b, bc_rows=0); mf.impose_dirichlet(1.0, A, b, bc_rows=1)
        check((not have("crc_coefs") or crc(A.packed_coefs) == want("crc_coefs")) and crc(b.coefs) == want("crc_b"), "%s A/b after Dirichlet CRC" % tag)
        mf.make_local_matrix(A)
        check((not have("crc_cols_local") or crc(A.packed_cols) == want("crc_cols_local")) and A.num_cols == int(want("num_cols")), "%s local columns CRC" % tag)
        for m in ("neighbors", "recv_length", "send_length"):
            check(np.array_equal(getattr(A, m), g["%s_%d" % (m, rank)]), "%s %s" % (tag, m))
        for m in ("external_index", "external_local_index", "elements_to_send"):
            a = getattr(A, m)
            check(len(a) == int(g["len_%s_%d" % (m, rank)]) and crc(a) == g["crc_%s_%d" % (m, rank)], "%s %s CRC" % (tag, m))
        iters, normr, times, h = mf.cg_solve(A, b, x, mf.matvec_std(), 200, EPS, history=True)
        check(iters == len(g["normr"]) - 1, "%s iterations %d vs %d" % (tag, iters, len(g["normr"]) - 1))
        k = min(iters, len(g["normr"]) - 1)
        for name in ("normr", "rtrans", "pap", "alpha"):
            check(np.array_equal(bits(h[name][1 if name in ("pap", "alpha") else 0:k + 1]), bits(g[name][1 if name in ("pap", "alpha") else 0:k + 1])), "%s %s history" % (tag, name))
        check(crc(x.coefs) == g["crc_x_%d" % rank], "%s solution CRC" % tag)
        if rank == 0:
            print("big %s np=%d: iters %d, final resid %.17g (golden %.17g)" % (tag, world, iters, h["normr"][k], g["normr"][k]), flush=True)
        mesh.close()
    t = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("MGPU_CHECK %s (%d mismatches over %d ranks)" % ("OK" if int(t) == 0 else "FAILED", int(t), world), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if int(t) == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
