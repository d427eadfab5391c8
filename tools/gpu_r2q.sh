#!/bin/bash
# single GPU: slot-order internal permutation -- tests, netgen 13.6 M / 108 M, DRAM traffic (ncu) at 13.6 M, bench sanity
TAG=${1:-r2q}
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dist.py ) > $O/${TAG}_pytest_gpu.log 2>&1; tail -5 $O/${TAG}_pytest_gpu.log
source oracle/_ref/ngs/env.sh
timeout 600 python tools/netgen_big.py --nref 2 --cpu-iters 0 --full --out $O/${TAG}_netgen_big_nref2.json 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('spmv_kernel_ms','spmv_frac_of_peak','spmv_call_ms_incl_gather','cg_it_per_s','cg_frac_of_peak','full_solve_steps','sell_padding')})"
timeout 900 python tools/netgen_big.py --nref 3 --cpu-iters 0 --full --out $O/${TAG}_netgen_big_nref3.json 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('spmv_kernel_ms','spmv_frac_of_peak','spmv_call_ms_incl_gather','cg_it_per_s','cg_frac_of_peak','full_solve_steps','sell_padding','create_device_matrix_s')})"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sell_spmv -s 4 -c 1 --csv \
    --log-file $O/${TAG}_ncu_netgen13p6M_dram.csv python tools/netgen_big.py --nref 2 --cpu-iters 0 --iters 3 > /dev/null 2>&1; tail -3 $O/${TAG}_ncu_netgen13p6M_dram.csv | cut -c150-330
timeout 300 python bench.py --no-cpu-baseline --no-full-solve --no-netgen-check > $O/${TAG}_bench.json 2>/dev/null; python -c "import json;d=json.load(open('$O/${TAG}_bench.json'));print('bench', d['value'], d['roofline']['avg_launch_ms'], d['e2e']['value'])"
