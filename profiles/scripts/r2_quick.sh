#!/bin/bash
# quick GPU pass: GPU tests + default bench line (+ optional extra bench flags as ablation)   usage: r2_quick.sh TAG ["extra flags"]
TAG=${1:-r2q}
O=gpurun_out
(time python -m pytest tests -x -q -m gpu) > $O/${TAG}_pytest.log 2>&1; tail -4 $O/${TAG}_pytest.log
python bench.py --no-cpu --ttt-iters 0 > $O/${TAG}_bench.json 2> $O/${TAG}.err
python - <<PY
import json
d=json.load(open("$O/${TAG}_bench.json"))
print("cold", round(d["value"]), "warm", round(d["warm"]["value"]), "e2e", round(d["e2e"]["value"]), "parity", d.get("parity_check",{}).get("pass"))
print(d["roofline"]["launch_ms_all"])
PY
if [ -n "$2" ]; then
python bench.py --no-cpu --ttt-iters 0 --no-parity $2 > $O/${TAG}_bench_abl.json 2>> $O/${TAG}.err
python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_abl.json"))
print("ablation $2: cold", round(d["value"]), "warm", round(d["warm"]["value"]))
print(d["roofline"]["launch_ms_all"])
PY
fi
