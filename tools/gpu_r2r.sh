#!/bin/bash
# single GPU: persistent cooperative CG kernel -- tests, C1 through the adapter (netgen, maxh 0.03), 1.7 M-dof netgen, sanitizer
TAG=${1:-r2r}
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reorder.py tests/test_gpu_csrview.py tests/test_gpu_workloads.py tests/test_gpu_next.py -m gpu -q -x ) > $O/${TAG}_pytest.log 2>&1; tail -6 $O/${TAG}_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "persistent" > $O/${TAG}_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -2 $O/${TAG}_sanitizer.log
( source oracle/_ref/ngs/env.sh; timeout 600 python integration/run_ngsolve_dropin.py --maxh 0.03 > $O/${TAG}_dropin.json 2> $O/${TAG}_dropin.err ); python -c "import json;d=json.load(open('$O/${TAG}_dropin.json'));print({k:d[k] for k in ('ndof','cpu_steps','cpu_solve_s','dev_steps','dev_solve_s','dev_fused','rel_diff','fused_solve_s','pycg_solve_s')})"; tail -2 $O/${TAG}_dropin.err
for m in 1 0; do
python tools/netgen_scale.py --nref 1 --cache /dev/shm/ng1 --modes -1 --iters 300 --full --cpu-iters 0 --opt cg_persistent=$m 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1])['reorder=-1']; print('1.7M persistent=$m', round(d['cg_it_per_s'],1), d['cg_frac'], d['full_steps'], d['full_s'])"
done
python tools/bench_configs.py c1 --out $O/${TAG}_c1.jsonl 2>/dev/null | tail -2 | cut -c1-400
