#!/bin/bash
# 8-GPU call: (a) eight independent single-GPU replicas of one rank's slab problem, run at the same time -- what the box gives
# per GPU without any exchange; (b) bench N = 8 with a smaller product grid
TAG=${1:-r2j}
O=gpurun_out; mkdir -p $O
for g in 0 1 2 3 4 5 6 7; do
  CUDA_VISIBLE_DEVICES=$g timeout 300 python tools/sweep_slab.py --parts 8 --cps 96 --out $O/${TAG}_replica_$g.jsonl > /dev/null 2>&1 &
done
wait
cat $O/${TAG}_replica_*.jsonl | python -c "import sys,json; [print(round(d['spmv_ms'],4), round(d['cg_ms_per_iteration'],4)) for d in map(json.loads, sys.stdin)]"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 8 --steps 100 --warmup 3 --no-full-solve --opt spmv_ctas_per_sm=48 > $O/${TAG}_bench_n8_cps48.json 2> $O/${TAG}_bench_n8_cps48.err
python -c "import json;d=json.load(open('$O/${TAG}_bench_n8_cps48.json'));print('cps48', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['cg_update_kernels_ms_per_iteration'], d['e2e']['value'])"
