"""Host-side logic of the subtree sharding (CPU only, world_size 2 over gloo): the ranks' ownership masks partition the
compact layouts and a masked all-reduce re-assembles an iterate -- the gather the multi-GPU path uses."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import raocp_b200 as r
from raocp_b200.core.flatten import FlatProblem
from oracle import problems


def test_masks_partition_every_layout():
    """for ANY cut stage the device may report (rb_shard_info; the rule itself lives in rb_create only) the ranks' masks
    partition both compact layouts, and every rank > 0 owns one contiguous node range per stage below the cut"""
    flat = FlatProblem(problems.build(problems.spec("cfg2"), r.core))
    with pytest.raises(Exception):
        flat.shard_cut_stage()            # not decided on the host
    for t_c in (2, 3, 4):
        flat.shard_cut = t_c
        width = int(flat.stage_off[t_c + 1] - flat.stage_off[t_c])
        for world in (2, 3, 8):
            if width < world:
                with pytest.raises(Exception):
                    flat.shard_owned_nodes(0, world)
                continue
            pm = sum(flat.shard_masks(rk, world)[0].astype(int) for rk in range(world))
            dm = sum(flat.shard_masks(rk, world)[1].astype(int) for rk in range(world))
            assert np.all(pm == 1) and np.all(dm == 1)
            own = [flat.shard_owned_nodes(rk, world) for rk in range(world)]
            assert np.all(sum(o.astype(int) for o in own) == 1)
            assert own[0][: flat.stage_off[t_c]].all()
            for o in own[1:]:
                assert not o[: flat.stage_off[t_c]].any()
                for t in range(t_c, flat.num_stages):
                    idx = np.flatnonzero(o[flat.stage_off[t]: flat.stage_off[t + 1]])
                    assert idx.size > 0 and np.array_equal(idx, np.arange(idx[0], idx[-1] + 1))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat = FlatProblem(problems.build(problems.spec("cfg2"), r.core), shard=(rank, world))
    flat.shard_cut = 3          # what rb_shard_info reports for cfg2 (a GPU test checks host == device)
    rng = np.random.default_rng(11)
    truth_p, truth_d = rng.standard_normal(flat.np_), rng.standard_normal(flat.nd_)
    pm, dm = flat.shard_masks(rank, world)
    # a rank only knows its own entries (garbage elsewhere), like after a sharded solve
    mine_p = np.where(pm, truth_p, 1e30 * (rank + 1))
    mine_d = np.where(dm, truth_d, -1e30 * (rank + 1))
    p = torch.from_numpy(np.where(pm, mine_p, 0.0))
    d = torch.from_numpy(np.where(dm, mine_d, 0.0))
    dist.all_reduce(p)
    dist.all_reduce(d)
    out[rank] = bool(np.array_equal(p.numpy(), truth_p) and np.array_equal(d.numpy(), truth_d))
    dist.destroy_process_group()


def test_masked_all_reduce_reassembles_iterates_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(rk, world, 29533, out)) for rk in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(out[rk] for rk in range(world))
