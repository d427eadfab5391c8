#!/bin/bash
# First GPU call of the next round (one B200): re-validate HEAD, then the measurements DESIGN.md section 6 asks for.
# usage (under gpurun, about 6 minutes):  bash tools/gpu_next_round.sh [tag]
TAG=${1:-r2a}
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -q ) > $O/${TAG}_pytest_gpu.log 2>&1; tail -3 $O/${TAG}_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; cut -c1-300 $O/${TAG}_bench_n1.json
# the opt-in options at the bench size (cg_fold_u), on a 29.5 M-row complex and a 3.3 M-block-row 3x3 matrix (sell_c16_all)
timeout 240 python tools/ab_options.py --m 160 --mc 77 --mb 37 --out $O/${TAG}_ab_options.jsonl > /dev/null 2>&1; cat $O/${TAG}_ab_options.jsonl
# what one rank of the 2/4/8-GPU runs multiplies, swept over the grid size
timeout 240 python tools/sweep_slab.py --out $O/${TAG}_slab.jsonl > /dev/null 2>&1; cut -c1-220 $O/${TAG}_slab.jsonl
# ncu: the CG update kernel (5 streams in, 3 out, 5.4 TB/s) and the compressed complex loop
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cg_fused_kernel -s 4 -c 1 -f -o $O/${TAG}_prof_cg_fused \
    python bench.py --steps 10 --warmup 3 --no-full-solve --no-cpu-baseline > $O/${TAG}_ncu_cg_fused.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sell_spmv -s 3 -c 1 -f -o $O/${TAG}_prof_sell_complex_c16 \
    python tools/bench_configs.py c5 --scale 0.3 --spmv-only --opt sell_c16_all=1 > $O/${TAG}_ncu_complex.log 2>&1
for r in cg_fused sell_complex_c16; do
    ncu -i $O/${TAG}_prof_$r.ncu-rep --page raw --csv > $O/${TAG}_prof_${r}_raw.csv 2>/dev/null
    python tools/ncu_stall_summary.py $O/${TAG}_prof_$r.ncu-rep > $O/${TAG}_prof_${r}_stalls.txt 2>/dev/null
done
ls -la $O | tail -15
