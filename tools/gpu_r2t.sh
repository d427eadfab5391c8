#!/bin/bash
# single GPU: GMRES with the one-reduction orthogonalisation -- tests, A/B on the generated and the netgen C5, ncu of the two new kernels
TAG=${1:-r2t}
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reorder.py tests/test_gpu_workloads.py -m gpu -q -x -k "gmres or GMRES" ) > $O/${TAG}_pytest.log 2>&1; tail -5 $O/${TAG}_pytest.log
for m in 0 1; do
  timeout 300 python tools/bench_configs.py c5 --opt gmres_orth=$m --out $O/${TAG}_c5gen_orth$m.jsonl 2>$O/${TAG}_c5gen_orth$m.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('generated C5 gmres_orth=$m', {k:d.get(k) for k in ('rows','iterations','it_per_s','gmres_frac_of_measured_peak','gmres_mgs_frac_of_measured_peak','kernel_ms_by_class')})"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gmres_dots_kernel -s 24 -c 1 -o $O/${TAG}_ncu_gmres_dots python tools/bench_configs.py c5 --scale 0.5 > $O/${TAG}_ncu1.log 2>&1; tail -2 $O/${TAG}_ncu1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gmres_project_kernel -s 30 -c 1 -o $O/${TAG}_ncu_gmres_project python tools/bench_configs.py c5 --scale 0.5 > $O/${TAG}_ncu2.log 2>&1; tail -2 $O/${TAG}_ncu2.log
source oracle/_ref/ngs/env.sh
( time timeout 400 python tools/netgen_configs.py c5 --cpu-iters 0 --out $O/${TAG}_c5.json ) 2> $O/${TAG}_c5.err | cut -c1-1800
tail -5 $O/${TAG}_c5.err
