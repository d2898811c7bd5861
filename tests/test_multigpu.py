"""GPU, N >= 2: z-slab decomposition + NVLink halo exchange, one process per GPU (SPMD).
Every rank compares the full result of each program with the CPU oracle, for the fused
(in-kernel peer store) and the copy-based halo exchange and both flag mechanisms."""
import os
import socket
import subprocess
import sys

import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, cases, options=""):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PHYSIS_B200_OPTIONS=options)
        procs.append(subprocess.Popen([sys.executable, os.path.join(H.ROOT, "tests", "mgpu_worker.py")] + cases,
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=600)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} of {world} ({options}):\n{o[-3000:]}"


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("options", ["halo_push=1", "halo_push=0", "halo_push=1,sync_mode=1",
                                     "halo_push=1,sync_mode=0", "halo_push=0,sync_mode=0"])
def test_two_gpus_match_oracle(options):
    _run(2, ["diffusion", "pair", "himeno", "pstag", "api"], options)


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
def test_two_gpus_reference_system_tests():
    _run(2, ["golden"], "halo=2")


@pytest.mark.skipif(_ngpus() < 4, reason="needs at least 4 GPUs")
def test_four_gpus_match_oracle():
    _run(4, ["diffusion", "pair", "himeno", "pstag", "api"], "halo_push=1")


@pytest.mark.skipif(_ngpus() < 3, reason="needs at least 3 GPUs")
def test_three_gpus_uneven_slabs():
    _run(3, ["diffusion", "pair", "pstag", "api"], "halo_push=0")


@pytest.mark.skipif(_ngpus() < 3, reason="needs at least 3 GPUs")
def test_three_gpus_fused_pairs_uneven_slabs():
    _run(3, ["pair", "diffusion"], "halo_push=1")


def test_single_process_group_of_one():
    # WORLD_SIZE=1 through the same launcher path
    _run(1, ["diffusion", "pair", "api"], "")
