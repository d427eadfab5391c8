#!/bin/bash
# One GPU call: parity tests, bench line, the other configs, ncu launch list and one full capture of the SpMV kernel.
# usage (under gpurun): bash tools/gpu_validate.sh [tag]
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt
nproc >> $O/${TAG}_gpu.txt; free -g | head -2 >> $O/${TAG}_gpu.txt
( time timeout 1200 python -m pytest tests -m gpu -q ) > $O/${TAG}_pytest_gpu.log 2>&1
tail -3 $O/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
# memory checker over the smoke solve and the reorder tests (racecheck/initcheck are too slow for the suite; memcheck is the gate)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $O/${TAG}_sanitizer_smoke.log 2>&1; echo "sanitizer smoke rc=$?"; tail -3 $O/${TAG}_sanitizer_smoke.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_reorder.py -m gpu -q -x -k "rcm or reorder_equals or reorder_long" > $O/${TAG}_sanitizer_reorder.log 2>&1; echo "sanitizer reorder rc=$?"; tail -3 $O/${TAG}_sanitizer_reorder.log
timeout 600 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; cat $O/${TAG}_bench_n1.json
timeout 300 python bench.py --impl reference > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_ref.err; cat $O/${TAG}_bench_reference_arm.json
timeout 900 python tools/bench_configs.py c1 c2 c4 c5 --out $O/${TAG}_configs.jsonl > $O/${TAG}_configs.log 2>&1; tail -5 $O/${TAG}_configs.log
if [ -f oracle/_ref/ngs/env.sh ] && ls integration/_build/_ngsb200*.so >/dev/null 2>&1; then
    ( source oracle/_ref/ngs/env.sh; timeout 900 python integration/run_ngsolve_dropin.py --maxh 0.03 > $O/${TAG}_dropin.json 2> $O/${TAG}_dropin.err ); cat $O/${TAG}_dropin.json; tail -3 $O/${TAG}_dropin.err
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_ncu_launches_bench_n1.csv \
    python bench.py --steps 10 --warmup 3 --no-full-solve --no-cpu-baseline > $O/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sell_spmv -s 1 -c 2 -f -o $O/${TAG}_prof_sell_bench \
    python bench.py --steps 10 --warmup 3 --no-full-solve --no-cpu-baseline > $O/${TAG}_ncu_full.log 2>&1
ncu -i $O/${TAG}_prof_sell_bench.ncu-rep --page raw --csv > $O/${TAG}_prof_sell_bench_raw.csv 2>/dev/null
ls -la $O
