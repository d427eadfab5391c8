#!/bin/bash
# Round 2, 2-GPU call: distributed tests (incl. reference-pinned exchange and both push variants), bench at N = 2 with the
# fused product+push kernel (default), with the separate push kernel, and with the interface-first split.
TAG=${1:-r2e}
O=gpurun_out; mkdir -p $O
( time timeout 1200 python -m pytest tests/test_gpu_dist.py tests/test_gpu_dist_reference.py -m gpu -q ) > $O/${TAG}_pytest_dist.log 2>&1; tail -8 $O/${TAG}_pytest_dist.log
for opt in "" "--opt dist_fused_push=0" "--opt dist_overlap=1"; do
    name=$(echo "$opt" | tr -d ' -' | tr '=' '_'); name=${name:-default}
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        bench.py --gpus 2 --steps 100 --warmup 3 --no-full-solve $opt > $O/${TAG}_bench_n2_$name.json 2> $O/${TAG}_bench_n2_$name.err
    python -c "import json;d=json.load(open('$O/${TAG}_bench_n2_$name.json'));print('$name', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['cg_update_kernels_ms_per_iteration'], d['e2e']['value'], d['gpu_launches'])"; tail -2 $O/${TAG}_bench_n2_$name.err
done
ls -la $O | tail -6
