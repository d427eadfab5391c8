#!/bin/bash
# Round 2, validation on one GPU: whole GPU suite, smoke, bench line (with the 108 M-dof netgen check and its CPU reference),
# reference arm, ncu launch list + one full capture of the dominant kernel in the bench command
TAG=${1:-r2z}
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/${TAG}_pytest_gpu.log 2>&1; tail -4 $O/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
( time timeout 1500 python bench.py --steps 20 --warmup 5 ) > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; head -c 3000 $O/${TAG}_bench_n1.json; echo; tail -4 $O/${TAG}_bench_n1.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; head -c 1500 $O/${TAG}_bench_reference.json; echo; tail -4 $O/${TAG}_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_ncu_launches_bench_n1.csv \
    python bench.py --steps 10 --warmup 3 --no-full-solve --no-cpu-baseline --no-netgen-check > $O/${TAG}_ncu_launches.log 2>&1
python tools/launch_summary.py $O/${TAG}_ncu_launches_bench_n1.csv "python bench.py --steps 10 --warmup 3 --no-full-solve --no-cpu-baseline --no-netgen-check" > $O/${TAG}_ncu_launches_bench_n1_summary.txt 2>/dev/null; head -16 $O/${TAG}_ncu_launches_bench_n1_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sell_spmv -s 2 -c 1 -f -o $O/${TAG}_prof_sell_bench \
    python bench.py --steps 10 --warmup 3 --no-full-solve --no-cpu-baseline --no-netgen-check > $O/${TAG}_ncu_full.log 2>&1
python tools/ncu_summary.py $O/${TAG}_prof_sell_bench.ncu-rep > $O/${TAG}_ncu_spmv_sell_bench.txt 2>/dev/null; head -8 $O/${TAG}_ncu_spmv_sell_bench.txt; tail -1 $O/${TAG}_ncu_spmv_sell_bench.txt
