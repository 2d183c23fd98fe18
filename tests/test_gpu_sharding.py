"""Subtree sharding of one tree over the GPUs of the box, inside the driver-run suite: tests/multi_gpu_check.py
under torchrun at every even world size the box offers (skipped below two GPUs).  The check compares the assembled sharded
iterates with the NumPy oracle and with the single-GPU loop, for the peer-memory exchange and for the NCCL all-gather."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_loop_under_torchrun(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs, the box has {_gpus()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + world), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    if world > 2:
        cmd.append("--quick")
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    sys.stdout.write(out.stdout[-4000:])
    assert out.returncode == 0 and "MULTI_GPU_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_cut_stage_reported_by_the_device():
    """rb_shard_info is the single source of the cut (the host never re-derives the rule): consistent with the sweep plan the
    device built, and the host-side ownership masks derived from it partition the layouts"""
    import numpy as np
    import raocp_b200 as r
    from oracle import problems
    for name in ("shard", "cfg2", "chain2010", "cfg5"):
        s = problems.spec(name)
        solver = r.core.Solver(problems.build(s, r.core), verbose=False)
        flat, dev = solver.cache.flat_problem, solver.cache.device_solver
        cut, first, width, chain = dev.shard_info()
        assert flat.shard_cut == cut and first == flat.stage_off[cut] and width == flat.stage_off[cut + 1] - first
        assert chain < 0 or chain > cut
        for world in (2, 3):
            if width < world:
                continue
            pm = sum(flat.shard_masks(rk, world)[0].astype(int) for rk in range(world))
            dm = sum(flat.shard_masks(rk, world)[1].astype(int) for rk in range(world))
            assert np.all(pm == 1) and np.all(dm == 1)
