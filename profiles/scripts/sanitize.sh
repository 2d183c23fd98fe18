#!/bin/bash
# compute-sanitizer pass over every kernel family switch.  One B200, under gpurun:
#   bash profiles/scripts/sanitize.sh r2   ->  gpurun_out/r2_sanitizer/*.log + summary.txt (copy into profiles/sanitizer/)
TAG=${1:-r2}
O=gpurun_out/${TAG}_sanitizer
mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
S=profiles/scripts/sanitize_case.py
: > $O/summary.txt
run() {   # tool, label, args...
  local tool=$1 label=$2; shift 2
  local log=$O/${tool}_${label}.log
  timeout 900 $CS --tool $tool --error-exitcode 77 --print-limit 20 python $S "$@" > $log 2>&1
  local rc=$?
  local errs=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $log | tail -1)
  echo "$tool $label rc=$rc :: $errs :: $(grep sanitize_case $log | tail -1)" | tee -a $O/summary.txt
}
for tool in memcheck racecheck synccheck initcheck; do
  # pipelined default: MMA chain walkers, fused tree kernel (cooperative, flag hand-off), chain / risk / lane dual passes
  run $tool chain2010_p1_t2_m1 chain2010 4 1 2 1
  # four warps per chain tile (named barriers, producer warps), per-level tree kernels, forward split
  run $tool chain2010_p3_t1_m2 chain2010 4 3 1 2
  # unpipelined loop, global-memory stage kernels, warp-per-chain walker
  run $tool chain2010_p0_t0_m0 chain2010 4 0 0 0
  # cfg5's sizes: BIG chain walkers (fragments in shared memory), wide-row tree kernels
  run $tool chain6432_p1_t2_m1 chain6432 3 1 2 1
  # small trees: cfg1 (31 nodes), a batch of 3 instances (instance-major kernels), a batch of 70 (panel kernels, batch.cu)
  run $tool cfg1_p1_t2_m1 cfg1 6 1 2 1
  run $tool mini3_b3_p1_t2_m1 mini3 4 1 2 1 3
  run $tool mini3_b70_panels mini3 4 1 2 1 70
done
cat $O/summary.txt
