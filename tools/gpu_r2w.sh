#!/bin/bash
# eight GPUs (gpurun --gpus 8): BASELINE config 5 (complex Helmholtz, 29.5 M dofs, Jacobi-GMRES) with the batched and the serial
# orthogonalisation -> profiles/r2_c5_gmres_n8.jsonl
mkdir -p gpurun_out
timeout 44 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    tools/bench_c5_multi.py --orth 1 0 > gpurun_out/r2w_c5_n8.jsonl 2> gpurun_out/r2w_c5_n8.err
cut -c1-400 gpurun_out/r2w_c5_n8.jsonl; tail -3 gpurun_out/r2w_c5_n8.err
