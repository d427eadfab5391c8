#!/bin/bash
# Round 2, second GPU call (one B200): reorder + drop-in tests, netgen 13.6 M dofs with whole rows + the reference's OWN
# device path (ngscuda) on the same system, CG update kernel A/B (grid-stride vs chunked), ncu --set full of the two kernels.
TAG=${1:-r2b}
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_reorder.py tests/test_gpu_dropin.py -m gpu -q ) > $O/${TAG}_pytest_new.log 2>&1; tail -25 $O/${TAG}_pytest_new.log
( time timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dropin.py --deselect tests/test_gpu_reorder.py ) > $O/${TAG}_pytest_gpu.log 2>&1; tail -5 $O/${TAG}_pytest_gpu.log
timeout 900 python tools/netgen_scale.py --nref 2 --cache /dev/shm/ng2 --full --gpu-reference --out $O/${TAG}_netgen_14M.json > /dev/null 2> $O/${TAG}_netgen_14M.err; cut -c1-3500 $O/${TAG}_netgen_14M.json; tail -3 $O/${TAG}_netgen_14M.err
# drop-in at C1 size (maxh 0.03: 0.96 M dofs) with the automatic reorder mode
( source oracle/_ref/ngs/env.sh; timeout 600 python integration/run_ngsolve_dropin.py --maxh 0.03 > $O/${TAG}_dropin.json 2> $O/${TAG}_dropin.err ); cut -c1-2500 $O/${TAG}_dropin.json; tail -3 $O/${TAG}_dropin.err
for opt in cg_chunked=0 cg_chunked=1; do
  timeout 300 python bench.py --no-cpu-baseline --no-full-solve --opt $opt > $O/${TAG}_bench_$opt.json 2> $O/${TAG}_bench_$opt.err
  python -c "import json,sys;d=json.load(open('$O/${TAG}_bench_$opt.json'));print('$opt', d['value'], d['roofline']['avg_launch_ms'], d['roofline']['cg_update_kernels_ms_per_iteration'], d['e2e']['value'])"; tail -2 $O/${TAG}_bench_$opt.err
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sell_spmv -s 3 -c 1 -f -o $O/${TAG}_prof_sell_netgen14M \
    python tools/netgen_scale.py --nref 2 --cache /dev/shm/ng2 --modes 1 --spmv-only --reps 3 > $O/${TAG}_ncu_sell.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cg_fused_kernel -s 4 -c 1 -f -o $O/${TAG}_prof_cg_fused \
    python bench.py --steps 10 --warmup 3 --no-full-solve --no-cpu-baseline > $O/${TAG}_ncu_cg_fused.log 2>&1
for r in sell_netgen14M cg_fused; do
    ncu -i $O/${TAG}_prof_$r.ncu-rep --page raw --csv > $O/${TAG}_prof_${r}_raw.csv 2>/dev/null
    python tools/ncu_stall_summary.py $O/${TAG}_prof_$r.ncu-rep > $O/${TAG}_prof_${r}_stalls.txt 2>/dev/null
done
ls -la $O | tail -14
