#!/bin/bash
# Round-end measurement pass (one B200): GPU tests, bench lines of every configuration, reference arm, ncu launch list of the
# default bench command and one full ncu capture (never a bench value from a run under ncu).  Outputs: gpurun_out/$TAG_*.
TAG=${1:-r1d}
O=gpurun_out
(time python -m pytest tests -x -q -m gpu) > $O/${TAG}_pytest.log 2>&1; tail -2 $O/${TAG}_pytest.log
python bench.py > $O/${TAG}_final.json 2> $O/${TAG}.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_final_ref.json 2>> $O/${TAG}.err
python bench.py --no-cpu --workload cfg1 --ttt-iters 200000 > $O/${TAG}_cfg1.json 2>> $O/${TAG}.err
python bench.py --no-cpu --workload cfg2 --ttt-iters 200000 > $O/${TAG}_cfg2.json 2>> $O/${TAG}.err
python bench.py --no-cpu --workload cfg4 --batch 4096 --steps 200 > $O/${TAG}_cfg4.json 2>> $O/${TAG}.err
python bench.py --no-cpu --workload cfg5 --steps 1000 > $O/${TAG}_cfg5.json 2>> $O/${TAG}.err
python bench.py --no-cpu --batch 8 --steps 500 > $O/${TAG}_b8.json 2>> $O/${TAG}.err
python bench.py --no-cpu --batch 32 --steps 200 > $O/${TAG}_b32.json 2>> $O/${TAG}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --ttt-iters 0 > $O/${TAG}_ncu_l.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_cfg5_launches.csv \
    python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu --ttt-iters 0 > $O/${TAG}_ncu_l5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_chain_mma|k_dual_chain' -s 12 -c 3 -f -o $O/${TAG}_cfg5_chain \
    python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu --ttt-iters 0 > $O/${TAG}_ncu_c5.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
ls $O/${TAG}_*
