#!/bin/bash
# single GPU: per-kernel durations of the CG loop at slab size (what one rank of the 8-GPU run computes), ncu launch list
TAG=${1:-r2i}
O=gpurun_out; mkdir -p $O
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 120 --log-file $O/${TAG}_ncu_launches_slab8.csv \
    python tools/sweep_slab.py --parts 8 --cps 96 --steps 16 > $O/${TAG}_ncu_slab.log 2>&1
python tools/launch_summary.py $O/${TAG}_ncu_launches_slab8.csv 2>/dev/null | head -30
timeout 300 python tools/sweep_slab.py --parts 8 --cps 32 48 96 --out $O/${TAG}_slab.jsonl > /dev/null 2>&1; cut -c1-260 $O/${TAG}_slab.jsonl
NGSB_NO_CUDA_GRAPH=1 timeout 300 python tools/sweep_slab.py --parts 8 --cps 96 --out $O/${TAG}_slab_nograph.jsonl > /dev/null 2>&1; cut -c1-260 $O/${TAG}_slab_nograph.jsonl
