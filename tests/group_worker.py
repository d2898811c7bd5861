"""Worker of tests/test_group_cpu.py: one rank of a CPU-only process group.
Checks the b200 runtime's host-side group logic (shared-memory rendezvous, barrier,
all-gather, z partition) against torch.distributed (gloo) on the same ranks."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import physis_b200
    lib = physis_b200.load_runtime()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # 1. rendezvous / barrier / all-gather round trips inside the C library
    lib.__PSB200GroupSelfTest.restype = C.c_int
    bad = lib.__PSB200GroupSelfTest()
    t = torch.tensor([bad])
    dist.all_reduce(t)
    assert int(t) == 0, f"group self test failed: {int(t)} mismatches"
    # 2. every rank's slab, gathered with gloo, tiles the dimension exactly
    lib.__PSB200Partition.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int)] * 2
    for n, dn in ((512, 0), (513, 512), (7, 7), (10, 0), (1025, 1024)):
        off, ln = C.c_int(), C.c_int()
        lib.__PSB200Partition(n, dn, world, rank, C.byref(off), C.byref(ln))
        mine = torch.tensor([off.value, ln.value])
        allv = [torch.zeros(2, dtype=torch.long) for _ in range(world)]
        dist.all_gather(allv, mine)
        pos = 0
        for o, l in (tuple(int(x) for x in v) for v in allv):
            assert o == pos, (n, dn, allv)
            pos += l
        assert pos == n
        base = (dn or n) // world
        for r, v in enumerate(allv):
            extra = int(v[1]) - base
            assert extra in (0, 1) or (r == world - 1 and dn and extra == 1 + (1 if dn % world else 0)) or (r == world - 1 and dn), (n, dn, allv)
    dist.barrier()
    dist.destroy_process_group()
    print("ok", rank)


if __name__ == "__main__":
    main()
