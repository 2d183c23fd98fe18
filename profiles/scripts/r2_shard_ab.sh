#!/bin/bash
# sharded-loop ablations on N GPUs: exchange inside the tree kernel (default) / one launch / three launches    usage: r2_shard_ab.sh TAG N
TAG=${1:-r2ab}; N=${2:-2}
O=gpurun_out
for mode in tree kernel split; do
  m=$mode; [ $mode = tree ] && m=""
  RAOCP_SHARD_XCHG=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu --no-parity --ttt-iters 0 > $O/${TAG}_${mode}_${N}.json 2> $O/${TAG}_${mode}_${N}.err
  python - <<PY
import json
d=json.load(open("$O/${TAG}_${mode}_${N}.json"))
st=d.get("sharded_cfg3_strong",{})
print("$mode: weak", round(d["value"]), "us/it", round(d["ms_per_step"]*1e3,1), "warm", round(d["warm"]["value"]), "| cfg3 strong cold", round(st.get("value",0)), "warm", round(st.get("warm_value",0)))
PY
done
