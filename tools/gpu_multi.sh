#!/bin/bash
# multi-GPU call: distributed tests + bench lines at N = 2 .. NG.  usage: bash tools/gpu_multi.sh <NG> [tag] [tests=1] [bench options]
# e.g. the A/B of the interface-first split:  bash tools/gpu_multi.sh 8 r2ov 0 "--opt dist_overlap=1"
NG=${1:-2}; TAG=${2:-r1}; TESTS=${3:-1}; OPTS=${4:-}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/${TAG}_gpus.txt
if [ "$TESTS" = 1 ]; then
    ( time timeout 900 python -m pytest tests/test_gpu_dist.py -q ) > $O/${TAG}_pytest_gpu_dist.log 2>&1; tail -4 $O/${TAG}_pytest_gpu_dist.log
fi
for N in 2 4 8; do
    [ $N -le $NG ] || continue
    [ "$TESTS" = 0 ] && [ $N -ne $NG ] && continue
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
        bench.py --gpus $N --steps 50 --warmup 3 $OPTS > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err
    cat $O/${TAG}_bench_n$N.json; tail -2 $O/${TAG}_bench_n$N.err
done
