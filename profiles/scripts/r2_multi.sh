#!/bin/bash
# multi-GPU pass: sharded parity check + bench at N GPUs    usage: r2_multi.sh TAG N [extra bench flags]
TAG=${1:-r2mg}; N=${2:-2}
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py $([ $N -gt 2 ] && echo --quick) > $O/${TAG}_mg${N}.log 2>&1
grep -a "rank 0\|MULTI\|skipped" $O/${TAG}_mg${N}.log | cut -c1-260
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu --no-parity --ttt-iters 0 $3 > $O/${TAG}_bench${N}.json 2> $O/${TAG}_bench${N}.err
python - <<PY
import json
d=json.load(open("$O/${TAG}_bench${N}.json"))
print("headline", round(d["value"]), d["scaling"], "ms/step", round(d["ms_per_step"]*1e3,1), "us; warm", round(d["warm"]["value"]), "e2e", round(d["e2e"]["value"]))
print("strong", {k:(round(v,1) if isinstance(v,float) else v) for k,v in d.get("sharded_cfg3_strong",{}).items() if k!="note"})
print("replicas", round(d.get("replicas",{}).get("value",0)))
print(d["config"]["workload"][:200])
PY
# cfg4: the batch of 4096 initial states split over the GPUs (instance-parallel, no collective)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload cfg4 --batch $((4096 / N)) --steps 100 --warmup 10 --no-cpu --no-parity --ttt-iters 0 > $O/${TAG}_cfg4_${N}.json 2>> $O/${TAG}_bench${N}.err
python - <<PY
import json
d=json.load(open("$O/${TAG}_cfg4_${N}.json"))
print("cfg4 x 4096 over $N GPUs:", round(d["value"]), "instance-it/s =", round(d["value"]/4096,1), "batch-it/s; per GPU batch", d["config"]["instances_per_gpu"], "ms/step", round(d["ms_per_step"],3))
PY
tail -2 $O/${TAG}_bench${N}.err
