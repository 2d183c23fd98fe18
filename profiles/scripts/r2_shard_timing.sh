#!/bin/bash
# where the pipelined sharded iteration spends its time, per exchange form    usage: r2_shard_timing.sh N
N=${1:-2}
for which in cfg3 wide; do
for mode in "" kernel; do
  echo "== $which, RAOCP_SHARD_XCHG='$mode'"
  RB_SHARD_TIMING=1 RAOCP_SHARD_XCHG=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 profiles/scripts/shard_timing.py $which > gpurun_out/st_${which}_${mode:-top}.log 2>&1
  grep -a "shard timing, rank 0" gpurun_out/st_${which}_${mode:-top}.log | tail -1; grep -a "Exception\|rror" gpurun_out/st_${which}_${mode:-top}.log | head -3
done
done
