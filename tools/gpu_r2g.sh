#!/bin/bash
# 2-GPU call: bench at N = 2, fused product+push (default) against the separate push kernel
TAG=${1:-r2g}
O=gpurun_out; mkdir -p $O
for opt in "" "--opt dist_fused_push=0" ""; do
    name=$(echo "$opt" | tr -d ' -' | tr '=' '_'); name=${name:-default}
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        bench.py --gpus 2 --steps 100 --warmup 3 --no-full-solve $opt > $O/${TAG}_bench_n2_$name.json 2> $O/${TAG}_bench_n2_$name.err
    python -c "import json;d=json.load(open('$O/${TAG}_bench_n2_$name.json'));print('$name', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['cg_update_kernels_ms_per_iteration'], d['e2e']['value'], d['gpu_launches'])"; tail -1 $O/${TAG}_bench_n2_$name.err
done
