"""CPU, world_size 2 and 3 (gloo): the host-side logic of the multi-GPU path -- the
shared-memory rendezvous the per-GPU processes use to exchange CUDA IPC handles and
reduction partials, and the z-slab partition -- exercised without any GPU."""
import ctypes as C
import os
import socket
import subprocess
import sys

import pytest

import helpers as H


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_group_rendezvous_and_partition_with_gloo(world):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(H.ROOT, "tests", "group_worker.py")],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r}:\n{o}"
        assert f"ok {r}" in o


def test_partition_matches_reference_rule():
    """floor(n/P) planes each, the remainder one each to the LAST ranks
    (runtime/grid_space_mpi.h:741-777)."""
    import physis_b200
    lib = physis_b200.load_runtime()
    lib.__PSB200Partition.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int)] * 2
    for n in (1, 7, 8, 9, 100, 512, 1023):
        for world in (1, 2, 3, 4, 8):
            want_off, want_len = [], []
            base, rem = divmod(n, world)
            pos = 0
            for r in range(world):
                ln = base + (1 if r >= world - rem else 0)
                want_off.append(pos)
                want_len.append(ln)
                pos += ln
            for r in range(world):
                off, ln = C.c_int(), C.c_int()
                lib.__PSB200Partition(n, 0, world, r, C.byref(off), C.byref(ln))
                assert (off.value, ln.value) == (want_off[r], want_len[r]), (n, world, r)
    # staggered grid (N+1) on an N-plane domain: same cuts, last rank takes the extra plane
    for world in (2, 4, 8):
        for r in range(world):
            a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            lib.__PSB200Partition(512, 512, world, r, C.byref(a), C.byref(b))
            lib.__PSB200Partition(513, 512, world, r, C.byref(c), C.byref(d))
            assert a.value == c.value
            assert d.value == b.value + (1 if r == world - 1 else 0)
