#!/bin/bash
# quick sharded pass: timing breakdown + quick parity check + bench    usage: r2_shard_quick.sh TAG N
TAG=${1:-r2q}; N=${2:-2}
O=gpurun_out
for which in cfg3 wide; do
  RB_SHARD_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 profiles/scripts/shard_timing.py $which > $O/${TAG}_st_${which}.log 2>&1
  echo "$which: $(grep -a 'shard timing, rank 0' $O/${TAG}_st_${which}.log | tail -1 | sed 's/.*iterations\] //' | cut -c1-200)"; grep -a "Exception" $O/${TAG}_st_${which}.log | head -2
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py --quick > $O/${TAG}_mg${N}.log 2>&1
grep -a "MULTI\|FAIL" $O/${TAG}_mg${N}.log | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu --no-parity --ttt-iters 0 > $O/${TAG}_bench${N}.json 2> $O/${TAG}_bench${N}.err
python - <<PY
import json
d=json.load(open("$O/${TAG}_bench${N}.json"))
st=d.get("sharded_cfg3_strong",{})
print("weak", round(d["value"]), "us/it", round(d["ms_per_step"]*1e3,1), "warm", round(d["warm"]["value"]), "e2e", round(d["e2e"]["value"]), "| cfg3 strong cold", round(st.get("value",0)), "warm", round(st.get("warm_value",0)), "| replicas", round(d.get("replicas",{}).get("value",0)))
PY
