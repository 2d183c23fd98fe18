#!/bin/bash
# perf iteration: sweeps parity tests only + bench (default and one ablation)   usage: r2_perf.sh TAG ["ablation flags"] [workload]
TAG=${1:-r2p}
WL=${3:-cfg3}
O=gpurun_out
(python -m pytest tests/test_gpu_sweeps.py -x -q -m gpu) > $O/${TAG}_pytest.log 2>&1; tail -1 $O/${TAG}_pytest.log
show() { python - "$1" "$2" <<PY
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[2], "cold", round(d["value"]), "warm", round(d["warm"]["value"]), "e2e", round(d["e2e"]["value"]), "parity", d.get("parity_check",{}).get("pass"))
print("   ", {k:[round(x*1e3,1) for x in (v if isinstance(v,list) else [v])] for k,v in d["roofline"]["launch_ms_all"].items() if k != "note"})
PY
}
python bench.py --workload $WL --no-cpu --ttt-iters 0 > $O/${TAG}_bench.json 2> $O/${TAG}.err; show $O/${TAG}_bench.json default
if [ -n "$2" ]; then
python bench.py --workload $WL --no-cpu --ttt-iters 0 --no-parity $2 > $O/${TAG}_bench_abl.json 2>> $O/${TAG}.err; show $O/${TAG}_bench_abl.json "$2"
fi
