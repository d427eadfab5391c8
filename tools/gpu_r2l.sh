#!/bin/bash
# single GPU: BASELINE configs[2] on a real netgen mesh (maxh 0.05 + 3x Refine, ~1.1e8 dofs), after a 13.6 M-dof sanity run
TAG=${1:-r2l}
O=gpurun_out; mkdir -p $O
source oracle/_ref/ngs/env.sh
timeout 600 python tools/netgen_big.py --nref 2 --full --out $O/${TAG}_netgen_big_nref2.json 2> $O/${TAG}_netgen_big_nref2.err | cut -c1-1800; tail -2 $O/${TAG}_netgen_big_nref2.err
( while true; do echo "$(date +%s) $(nvidia-smi --query-gpu=memory.used --format=csv,noheader,nounits) $(free -g | awk '/Mem:/{print $3}')"; sleep 2; done ) > $O/${TAG}_mem.log 2>/dev/null &
MEMPID=$!
timeout 1700 python tools/netgen_big.py --nref 3 --full --cpu-iters 3 --out $O/${TAG}_netgen_big_nref3.json 2> $O/${TAG}_netgen_big_nref3.err | cut -c1-2000; tail -12 $O/${TAG}_netgen_big_nref3.err
kill $MEMPID
echo "peak GPU MiB: $(awk '{print $2}' $O/${TAG}_mem.log | sort -n | tail -1)  peak host GiB: $(awk '{print $3}' $O/${TAG}_mem.log | sort -n | tail -1)"
