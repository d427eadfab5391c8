#!/bin/bash
# single GPU: A/B of the streaming hints in the CG update kernel (slab size and bench size), whole-round grid at slab size,
# DRAM traffic of the product on the 108 M-dof netgen system (ncu)
TAG=${1:-r2n}
O=gpurun_out; mkdir -p $O
for o in cg_stream_hints=0 cg_stream_hints=1; do
  timeout 300 python tools/sweep_slab.py --parts 8 --cps 0 48 --opt $o --out $O/${TAG}_slab.jsonl > /dev/null 2>&1
done
cut -c1-300 $O/${TAG}_slab.jsonl
for o in cg_stream_hints=0 cg_stream_hints=1; do
  timeout 300 python bench.py --no-cpu-baseline --no-full-solve --no-netgen-check --opt $o > $O/${TAG}_bench_$o.json 2> $O/${TAG}_bench_$o.err
  python -c "import json;d=json.load(open('$O/${TAG}_bench_$o.json'));print('$o', d['value'], d['roofline']['avg_launch_ms'], d['roofline']['cg_update_kernels_ms_per_iteration'], d['e2e']['value'])"
done
source oracle/_ref/ngs/env.sh
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct --clock-control none \
    -k regex:sell_spmv -s 4 -c 2 --csv --log-file $O/${TAG}_ncu_netgen108M_dram.csv python tools/netgen_big.py --nref 3 --cpu-iters 0 --iters 3 > /dev/null 2> $O/${TAG}_ncu_netgen108M.err
tail -11 $O/${TAG}_ncu_netgen108M_dram.csv | cut -c1-330
