#!/bin/bash
# Round-end measurement pass (one B200): forced rebuild on the box, GPU tests, bench lines of every configuration, reference arm,
# ncu launch lists and full captures (never a bench value from a run under ncu).  Outputs: gpurun_out/$TAG_*.
TAG=${1:-r2}
O=gpurun_out
(time python __graft_entry__.py --force) > $O/${TAG}_build.log 2>&1; tail -3 $O/${TAG}_build.log
(time python -m pytest tests -x -q -m gpu) > $O/${TAG}_pytest.log 2>&1; tail -4 $O/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
python bench.py > $O/${TAG}_final.json 2> $O/${TAG}.err
python bench.py --steps 20 --warmup 5 > $O/${TAG}_final_driver_flags.json 2>> $O/${TAG}.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_final_ref.json 2>> $O/${TAG}.err
python bench.py --no-cpu --workload cfg1 --ttt-iters 200000 > $O/${TAG}_cfg1.json 2>> $O/${TAG}.err
python bench.py --no-cpu --workload cfg2 --ttt-iters 200000 > $O/${TAG}_cfg2.json 2>> $O/${TAG}.err
python bench.py --no-cpu --workload cfg4 --batch 4096 --steps 200 > $O/${TAG}_cfg4.json 2>> $O/${TAG}.err
python bench.py --no-cpu --workload cfg4 --batch 4096 --steps 50 --no-panels --no-parity > $O/${TAG}_cfg4_instance_major.json 2>> $O/${TAG}.err
python bench.py --workload cfg5 --steps 1000 > $O/${TAG}_cfg5.json 2>> $O/${TAG}.err
python bench.py --no-cpu --workload cfg5 --steps 1000 --mma-one-warp --no-parity --ttt-iters 0 > $O/${TAG}_cfg5_one_warp.json 2>> $O/${TAG}.err
python bench.py --no-cpu --no-dedup --no-parity --ttt-iters 0 > $O/${TAG}_cfg3_nodedup.json 2>> $O/${TAG}.err
python bench.py --no-cpu --mma-four-warps --no-parity --ttt-iters 0 > $O/${TAG}_cfg3_four_warps.json 2>> $O/${TAG}.err
python bench.py --no-cpu --overlap --no-parity --ttt-iters 0 > $O/${TAG}_cfg3_launch_overlap.json 2>> $O/${TAG}.err
python bench.py --no-cpu --batch 8 --steps 500 --no-parity --ttt-iters 0 > $O/${TAG}_b8.json 2>> $O/${TAG}.err
python profiles/scripts/determinism.py 12 > $O/${TAG}_determinism.log 2>&1; tail -1 $O/${TAG}_determinism.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-parity --ttt-iters 0 > $O/${TAG}_ncu_l.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_cfg4_launches.csv \
    python bench.py --workload cfg4 --batch 4096 --steps 5 --warmup 3 --no-cpu --no-parity --ttt-iters 0 > $O/${TAG}_ncu_l4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_chain_mma|k_tree_fused|k_dual_chain' -s 40 -c 5 -f -o $O/${TAG}_cfg3_crit \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-parity --ttt-iters 0 > $O/${TAG}_ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_bp_dual|k_bp_kproj|k_bp_top' -s 30 -c 6 -f -o $O/${TAG}_cfg4_dual \
    python bench.py --workload cfg4 --batch 4096 --steps 3 --warmup 3 --no-cpu --no-parity --ttt-iters 0 > $O/${TAG}_ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_chain_mma|k_tree_fused|k_dual_chain' -s 30 -c 4 -f -o $O/${TAG}_cfg5_crit \
    python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu --no-parity --ttt-iters 0 > $O/${TAG}_ncu_c5.log 2>&1
ls $O/${TAG}_* | head -60
