#!/bin/bash
# cfg4 A/B of a compile-time switch: default build (tests + bench), then the build with $1 (bench only)
O=gpurun_out
bash profiles/scripts/r2_cfg4.sh r2t
RB_NVCC_EXTRA="$1" python __graft_entry__.py > /dev/null 2>&1
python bench.py --workload cfg4 --batch 4096 --steps 200 --no-cpu --no-parity --ttt-iters 0 > $O/r2t_abl.json 2>/dev/null
python - <<PY
import json
d=json.load(open("$O/r2t_abl.json"))
print("with $1:", round(d["value"]/4096,1), "batch-it/s", {k:[round(x*1e3,1) for x in (v if isinstance(v,list) else [v])] for k,v in d["roofline"]["launch_ms_all"].items() if k!="note"})
PY
