#!/usr/bin/env python3
"""oracle/mpi_shim/launch.py -- TEST INFRASTRUCTURE.

`mpirun -np N` for binaries built against oracle/mpi_shim/mpi.h: forks N copies
of the command, wiring every pair of ranks with one AF_UNIX socketpair whose
descriptors are passed through MPISHIM_FDS (comma separated, index = peer rank).

usage: launch.py N [--cwd DIR] -- cmd args...
"""
import os
import socket
import subprocess
import sys


def launch(nranks, cmd, cwd=None, timeout=None, stdout=None):
    pairs = {}
    for i in range(nranks):
        for j in range(i + 1, nranks):
            a, b = socket.socketpair(socket.AF_UNIX, socket.SOCK_STREAM)
            for s in (a, b):
                s.setsockopt(socket.SOL_SOCKET, socket.SO_SNDBUF, 1 << 20)
                s.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, 1 << 20)
            pairs[(i, j)] = a
            pairs[(j, i)] = b
    procs = []
    for r in range(nranks):
        fds = [pairs[(r, p)].fileno() if p != r else -1 for p in range(nranks)]
        env = dict(os.environ)
        env["MPISHIM_RANK"] = str(r)
        env["MPISHIM_SIZE"] = str(nranks)
        env["MPISHIM_FDS"] = ",".join(str(f) for f in fds)
        procs.append(subprocess.Popen(cmd, env=env, cwd=cwd, stdout=stdout,
                                      pass_fds=[f for f in fds if f >= 0]))
    for s in pairs.values():
        s.close()
    rc = 0
    for p in procs:
        try:
            rc |= p.wait(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    return rc


def main(argv):
    n = int(argv[1])
    rest = argv[2:]
    cwd = None
    if rest and rest[0] == "--cwd":
        cwd = rest[1]
        rest = rest[2:]
    if rest and rest[0] == "--":
        rest = rest[1:]
    return launch(n, rest, cwd=cwd)


if __name__ == "__main__":
    sys.exit(main(sys.argv))
