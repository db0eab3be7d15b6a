"""world_size-2/4 `gloo` test of the N>1 host logic on CPU (not gpu): every rank derives its halo index maps from
the box table alone (halo_maps.h, no messages -- the reference needs MPI_Irecv/Send for this), then the ranks run
exchange_externals over torch.distributed with exactly the send/recv lists the NCCL path uses, and a rank-ordered
all-gather dot.  Received externals and the global dot must equal the oracle's."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, so, q, golden=None):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from oracle_lib import Oracle
        H = C.CDLL(so)
        H.hs_create.restype = C.c_void_p; H.hs_create.argtypes = [C.c_int] * 5
        H.hs_create_boxes.restype = C.c_void_p; H.hs_create_boxes.argtypes = [C.c_int] * 5 + [C.c_void_p]
        boxes = None
        if golden:               # boxes left by the reference's add_imbalance (not an RCB partition)
            g = np.load(golden)
            boxes = np.ascontiguousarray(np.stack([g["box_%d" % r] for r in range(world)]), np.int32)
        H.hs_structure.argtypes = [C.c_void_p]; H.hs_make_local.argtypes = [C.c_void_p]
        H.hs_nrows.restype = C.c_longlong; H.hs_nrows.argtypes = [C.c_void_p]
        H.hs_num_cols.restype = C.c_longlong; H.hs_num_cols.argtypes = [C.c_void_p]
        H.hs_get_map.restype = C.c_longlong; H.hs_get_map.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        h = H.hs_create(*n, world, rank) if boxes is None else H.hs_create_boxes(*n, world, rank, boxes.ctypes.data_as(C.c_void_p))
        H.hs_structure(h); H.hs_make_local(h)

        def gm(name):
            c = H.hs_get_map(h, name.encode(), None); a = np.zeros(c, np.int32)
            if c:
                H.hs_get_map(h, name.encode(), a.ctypes.data_as(C.c_void_p))
            return a
        nrows, ncols = H.hs_nrows(h), H.hs_num_cols(h)
        nb, rl, sl, ets = gm("neighbors"), gm("recv_length"), gm("send_length"), gm("elements_to_send")
        o = Oracle(*n, world, boxes=boxes).setup()
        first = o.first_row(rank)
        x = np.zeros(ncols); x[:nrows] = np.cos(0.1 * (first + np.arange(nrows))) + 2.0
        # exchange_externals (ref/src/exchange_externals.hpp:90-153) with the product's lists
        reqs, ro, so_ = [], nrows, 0
        recv_bufs = []
        for i, p in enumerate(nb):
            buf = torch.zeros(int(rl[i]), dtype=torch.float64); recv_bufs.append((ro, buf))
            reqs.append(dist.irecv(buf, src=int(p))); ro += int(rl[i])
        for i, p in enumerate(nb):
            seg = torch.from_numpy(x[ets[so_:so_ + int(sl[i])]].copy()); so_ += int(sl[i])
            reqs.append(dist.isend(seg, dst=int(p)))
        for r in reqs:
            r.wait()
        for off, buf in recv_bufs:
            x[off:off + len(buf)] = buf.numpy()
        ext = o.get_map("external_index", rank); loc = o.get_map("external_local_index", rank)
        want = np.zeros(ncols - nrows); want[loc - nrows] = np.cos(0.1 * ext.astype(np.float64)) + 2.0
        ok = np.array_equal(x[nrows:].view(np.int64), want.view(np.int64))
        # the NVLink peer-store exchange (peer.cuh): the SENDER chooses where its values land in the receiver's
        # external tail (send_dst_offset, derived from the box table alone) -- must give the same tail
        sdo = gm("send_dst_offset")
        tail = np.zeros(ncols - nrows)
        reqs, so_, inbox = [], 0, []
        for i, p in enumerate(nb):
            buf = torch.zeros(1 + int(rl[i]), dtype=torch.float64); inbox.append(buf)
            reqs.append(dist.irecv(buf, src=int(p)))
        for i, p in enumerate(nb):
            seg = np.concatenate([[float(sdo[i])], x[ets[so_:so_ + int(sl[i])]]]); so_ += int(sl[i])
            reqs.append(dist.isend(torch.from_numpy(seg.copy()), dst=int(p)))
        for r in reqs:
            r.wait()
        for buf in inbox:
            off = int(buf[0]); tail[off:off + len(buf) - 1] = buf[1:].numpy()
        ok = ok and np.array_equal(tail.view(np.int64), want.view(np.int64))
        # dot: local serial sum, all-gather, sum in rank order (Vector_functions.hpp:218-222 through the MPI shim)
        local = torch.tensor([Oracle.dot(x[:nrows], x[:nrows])], dtype=torch.float64)
        parts = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, local)
        g = float(parts[0])
        for p in parts[1:]:
            g = g + float(p)
        xs = [np.cos(0.1 * (o.first_row(r) + np.arange(o.nrows(r)))) + 2.0 for r in range(world)]
        wg = None
        for r in range(world):
            l = Oracle.dot(xs[r], xs[r]); wg = l if wg is None else wg + l
        ok = ok and g == wg
        q.put((rank, bool(ok)))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:   # noqa
        q.put((rank, "error: %r" % (e,)))


@pytest.mark.parametrize("n,world,golden", [((8, 8, 8), 2, None), ((6, 5, 7), 2, None), ((8, 8, 8), 4, None),
                                            ((10, 9, 8), 4, os.path.join(HERE, "golden", "mpi_10x9x8_np4_imb60.npz"))])
def test_halo_exchange_and_ordered_dot_over_gloo(harness, oracle_built, n, world, golden):
    so = os.path.join(HERE, "libcpu_harness.so")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, so, q, golden)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)], res
