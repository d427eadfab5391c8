#!/bin/bash
# 8-GPU call: bench N = 8 (adaptive product grid, fused push) and the 13.6 M-dof netgen system on 8 GPUs
TAG=${1:-r2m}
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 8 --steps 100 --warmup 3 > $O/${TAG}_bench_n8.json 2> $O/${TAG}_bench_n8.err
python -c "import json;d=json.load(open('$O/${TAG}_bench_n8.json'));print('n8', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['cg_update_kernels_ms_per_iteration'], d['e2e']['value'], d['gpu_launches'], d['config']['full_solve'])"; tail -1 $O/${TAG}_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/netgen_multi.py \
    --cache /dev/shm/ng2 --nref 2 --out $O/${TAG}_netgen_14M_n8.json 2> $O/${TAG}_netgen_14M_n8.err | cut -c1-900; tail -2 $O/${TAG}_netgen_14M_n8.err
