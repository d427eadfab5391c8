#!/bin/bash
# Round 2, fourth GPU call (one B200): full GPU suite (incl. the reference-pinned distributed tests), compute-sanitizer, the
# reference's own device path (ngscuda) at C1 and 13.6 M dofs, slab-size sweep + ncu, bench line with netgen_check.
TAG=${1:-r2d}
O=gpurun_out; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -q ) > $O/${TAG}_pytest_gpu.log 2>&1; tail -8 $O/${TAG}_pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $O/${TAG}_sanitizer_smoke.log 2>&1; echo "sanitizer smoke rc=$?"; tail -3 $O/${TAG}_sanitizer_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_reorder.py -m gpu -q -x -k "rcm or reorder_equals or reorder_long" > $O/${TAG}_sanitizer_reorder.log 2>&1; echo "sanitizer reorder rc=$?"; tail -3 $O/${TAG}_sanitizer_reorder.log
( source oracle/_ref/ngs/env.sh; timeout 300 python oracle/ref_gpu_cg.py --maxh 0.03 --full > $O/${TAG}_refgpu_c1.json 2> $O/${TAG}_refgpu_c1.err ); cat $O/${TAG}_refgpu_c1.json; tail -4 $O/${TAG}_refgpu_c1.err
( source oracle/_ref/ngs/env.sh; timeout 600 python oracle/ref_gpu_cg.py --maxh 0.05 --nref 2 --iters 50 > $O/${TAG}_refgpu_14M.json 2> $O/${TAG}_refgpu_14M.err ); cat $O/${TAG}_refgpu_14M.json; tail -4 $O/${TAG}_refgpu_14M.err
timeout 400 python tools/sweep_slab.py --parts 8 --cps 48 64 96 110 --out $O/${TAG}_slab.jsonl > /dev/null 2> $O/${TAG}_slab.err; cut -c1-260 $O/${TAG}_slab.jsonl; tail -3 $O/${TAG}_slab.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sell_spmv -s 8 -c 1 -f -o $O/${TAG}_prof_sell_slab8 \
    python tools/sweep_slab.py --parts 8 --cps 96 > $O/${TAG}_ncu_slab.log 2>&1
python tools/ncu_summary.py $O/${TAG}_prof_sell_slab8.ncu-rep > $O/${TAG}_prof_sell_slab8.txt 2>/dev/null; head -12 $O/${TAG}_prof_sell_slab8.txt
python tools/ncu_stall_summary.py $O/${TAG}_prof_sell_slab8.ncu-rep > $O/${TAG}_prof_sell_slab8_stalls.txt 2>/dev/null
timeout 900 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; cat $O/${TAG}_bench_n1.json; tail -3 $O/${TAG}_bench_n1.err
ls -la $O | tail -8
