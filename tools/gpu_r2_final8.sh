#!/bin/bash
# Round 2, 8-GPU validation: many-GPU tests and the bench line at N = 8
TAG=${1:-r2y}
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -k "many_gpu or two_gpu_solve" ) > $O/${TAG}_pytest_many.log 2>&1; tail -3 $O/${TAG}_pytest_many.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 8 --steps 20 --warmup 5 > $O/${TAG}_bench_n8.json 2> $O/${TAG}_bench_n8.err
python -c "import json;d=json.load(open('$O/${TAG}_bench_n8.json'));print('n8', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['cg_update_kernels_ms_per_iteration'], d['e2e']['value'], d['gpu_launches'], d['config']['full_solve'], d['config']['partition'][:120])"; tail -1 $O/${TAG}_bench_n8.err
