#!/bin/bash
# 8-GPU call: the many-GPU solve tests, bench at N = 8 with the fused product+push kernel (default) and with the separate push kernel
TAG=${1:-r2h}
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -k "many_gpu" ) > $O/${TAG}_pytest_many.log 2>&1; tail -4 $O/${TAG}_pytest_many.log
for opt in "" "--opt dist_fused_push=0"; do
    name=$(echo "$opt" | tr -d ' -' | tr '=' '_'); name=${name:-default}
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
        bench.py --gpus 8 --steps 100 --warmup 3 $opt > $O/${TAG}_bench_n8_$name.json 2> $O/${TAG}_bench_n8_$name.err
    python -c "import json;d=json.load(open('$O/${TAG}_bench_n8_$name.json'));print('$name', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['cg_update_kernels_ms_per_iteration'], d['e2e']['value'], d['gpu_launches'], d['config']['full_solve'])"; tail -1 $O/${TAG}_bench_n8_$name.err
done
