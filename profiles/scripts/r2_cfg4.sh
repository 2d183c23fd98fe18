#!/bin/bash
# cfg4 iteration: panel tests + cfg4 x 4096 bench   usage: r2_cfg4.sh TAG [quick]
TAG=${1:-r2c4}
O=gpurun_out
python -m pytest tests/test_gpu_batch_panels.py tests/test_gpu_at_size.py -x -q -m gpu -k "panel or cfg4" 2>&1 | tail -3
python bench.py --workload cfg4 --batch 4096 --steps 200 --no-cpu --no-parity --ttt-iters 0 > $O/${TAG}_cfg4.json 2> $O/${TAG}.err
python - <<PY
import json
d=json.load(open("$O/${TAG}_cfg4.json"))
print("inst-it/s cold", round(d["value"]), "warm", round(d["warm"]["value"]), "e2e", round(d["e2e"]["value"]), "batch-it/s", round(d["value"]/4096,1), "iteration frac", round(d["roofline"]["iteration"]["frac"],3))
print({k:[round(x*1e3,1) for x in (v if isinstance(v,list) else [v])] for k,v in d["roofline"]["launch_ms_all"].items() if k != "note"})
PY
tail -2 $O/${TAG}.err
