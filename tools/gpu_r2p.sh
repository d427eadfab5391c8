#!/bin/bash
# single GPU: product grids beyond 96 CTAs per SM (down to one 8-slice step per CTA) on the 108 M-dof netgen system and on the generator system
TAG=${1:-r2p}
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_workloads.py -m gpu -q -x ) > $O/${TAG}_pytest.log 2>&1; tail -3 $O/${TAG}_pytest.log
for c in 96 384 1536 3000; do
  timeout 300 python bench.py --no-cpu-baseline --no-full-solve --no-netgen-check --steps 50 --opt spmv_ctas_per_sm=$c > $O/${TAG}_bench_cps$c.json 2> $O/${TAG}_bench_cps$c.err
  python -c "import json;d=json.load(open('$O/${TAG}_bench_cps$c.json'));print('generator cps $c', d['value'], d['roofline']['avg_launch_ms'])"; tail -1 $O/${TAG}_bench_cps$c.err
done
source oracle/_ref/ngs/env.sh
python tools/netgen_big.py --nref 3 --cpu-iters 0 --iters 50 --cps-sweep 96 192 384 768 1536 3072 --out $O/${TAG}_netgen108M_grid_sweep.json 2>&1 | grep "cps\|spmv\|cg "
