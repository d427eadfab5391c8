#!/bin/bash
# multi-GPU call: CG on the netgen-numbered systems across NG GPUs (tools/netgen_multi.py)
NG=${1:-2}; TAG=${2:-r2k}
O=gpurun_out; mkdir -p $O
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 tools/netgen_multi.py "$@"; }
run --cache /dev/shm/ng1 --nref 1 --cpu-full --out $O/${TAG}_netgen_1p7M_n$NG.json 2> $O/${TAG}_netgen_1p7M_n$NG.err | cut -c1-1200; tail -3 $O/${TAG}_netgen_1p7M_n$NG.err
run --cache /dev/shm/ng2 --nref 2 --out $O/${TAG}_netgen_14M_n$NG.json 2> $O/${TAG}_netgen_14M_n$NG.err | cut -c1-1500; tail -3 $O/${TAG}_netgen_14M_n$NG.err
