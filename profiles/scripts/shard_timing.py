"""Developer aid: where the pipelined sharded iteration spends its time (CUDA events on the main stream, plain launches).

    RB_SHARD_TIMING=1 [RAOCP_SHARD_XCHG=kernel] python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29515 profiles/scripts/shard_timing.py [cfg3|wide]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "raocp-toolbox_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import raocp_b200 as r  # noqa: E402
from oracle import problems  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
spec = problems.wide_spec(world) if which == "wide" else problems.spec("cfg3")
problem = problems.build(spec, r.core)
solver = r.core.Solver(problem, device=local, verbose=False, shard=(rank, world))
solver.cache.device_solver.shard_init()
alpha = solver.compute_step_size()
dist.barrier()
for _ in range(2):
    solver.chock(spec["x0"][:, :1], max_iters=400, tol=0.0, alpha=alpha)
dist.barrier()
dist.destroy_process_group()
