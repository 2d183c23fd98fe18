#!/bin/bash
# r2 first GPU pass: forced rebuild on the box, GPU tests, default bench line, ncu launch list + full capture of the critical path
O=gpurun_out
(time python __graft_entry__.py --force) > $O/r2a_build.log 2>&1; tail -3 $O/r2a_build.log
(time python -m pytest tests -x -q -m gpu) > $O/r2a_pytest.log 2>&1; tail -5 $O/r2a_pytest.log
python bench.py > $O/r2a_final.json 2> $O/r2a.err; tail -c 600 $O/r2a_final.json
ncu --set full --clock-control none --import-source on -k regex:'k_chain_mma|k_tree_fused|k_dual_chain' -s 40 -c 5 -f -o $O/r2a_crit \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-parity --ttt-iters 0 > $O/r2a_ncu.log 2>&1
ls -la $O/r2a_*
