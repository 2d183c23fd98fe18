#!/bin/bash
# cfg4 x 4096: launch-bound variants of the panel dual kernels (compile-time), one bench line each
O=gpurun_out
for v in "3 3 2" "4 4 2" "4 4 3" "5 5 2"; do
  set -- $v
  RB_NVCC_EXTRA="-DRB_BP_LEAF_MINB=$1 -DRB_BP_RISK_MINB=$2 -DRB_BP_XU_MINB=$3" python __graft_entry__.py > /dev/null 2>&1
  python bench.py --workload cfg4 --batch 4096 --steps 100 --no-cpu --no-parity --ttt-iters 0 > $O/r2p_$1$2$3.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("$O/r2p_$1$2$3.json"))
print("leaf/risk/xu minb $v:", round(d["value"]/4096,1), "batch-it/s", {k:[round(x*1e3,1) for x in (v if isinstance(v,list) else [v])] for k,v in d["roofline"]["launch_ms_all"].items()})
PY
done
