#!/bin/bash
# Round 2, third GPU call (one B200): tests, the reference's own device path (ngscuda) at C1 and 13.6 M dofs, option sweep of
# the SELL product on the 13.6 M-dof netgen system, ncu of the CG update kernel (a launch that does work).
TAG=${1:-r2c}
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_reorder.py tests/test_gpu_dropin.py -m gpu -q ) > $O/${TAG}_pytest_new.log 2>&1; tail -8 $O/${TAG}_pytest_new.log
( source oracle/_ref/ngs/env.sh; timeout 300 python oracle/ref_gpu_cg.py --maxh 0.03 --full > $O/${TAG}_refgpu_c1.json 2> $O/${TAG}_refgpu_c1.err ); cat $O/${TAG}_refgpu_c1.json; tail -12 $O/${TAG}_refgpu_c1.err
( source oracle/_ref/ngs/env.sh; timeout 600 python oracle/ref_gpu_cg.py --maxh 0.05 --nref 2 --iters 50 > $O/${TAG}_refgpu_14M.json 2> $O/${TAG}_refgpu_14M.err ); cat $O/${TAG}_refgpu_14M.json; tail -5 $O/${TAG}_refgpu_14M.err
timeout 600 python tools/netgen_scale.py --nref 2 --cache /dev/shm/ng2 --full --out $O/${TAG}_netgen_14M.json > /dev/null 2> $O/${TAG}_netgen_14M.err; cut -c1-3000 $O/${TAG}_netgen_14M.json; tail -3 $O/${TAG}_netgen_14M.err
timeout 600 python tools/sweep_netgen.py --cache /dev/shm/ng2 --sigma 1024 4096 16384 65536 --variant 0 1 2 --cps 0 48 --out $O/${TAG}_sweep_netgen14M.jsonl 2> $O/${TAG}_sweep.err | cut -c1-330; tail -3 $O/${TAG}_sweep.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cg_fused_kernel -s 2 -c 1 -f -o $O/${TAG}_prof_cg_fused \
    python bench.py --steps 10 --warmup 3 --no-full-solve --no-cpu-baseline > $O/${TAG}_ncu_cg_fused.log 2>&1
python tools/ncu_summary.py $O/${TAG}_prof_cg_fused.ncu-rep > $O/${TAG}_prof_cg_fused.txt 2>/dev/null; head -40 $O/${TAG}_prof_cg_fused.txt
python tools/ncu_stall_summary.py $O/${TAG}_prof_cg_fused.ncu-rep > $O/${TAG}_prof_cg_fused_stalls.txt 2>/dev/null
ls -la $O | tail -8
