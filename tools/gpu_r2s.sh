#!/bin/bash
# single GPU: BASELINE configs C4, C2, C5 on real netgen meshes at the named sizes (assembled by the reference build on the box)
TAG=${1:-r2s}
O=gpurun_out; mkdir -p $O
source oracle/_ref/ngs/env.sh
nproc; free -g | head -2
( time timeout 210 python tools/netgen_configs.py c4 --out $O/${TAG}_c4.json ) 2> $O/${TAG}_c4.err | cut -c1-1500
tail -4 $O/${TAG}_c4.err
( time timeout 240 python tools/netgen_configs.py c2 --out $O/${TAG}_c2.json ) 2> $O/${TAG}_c2.err | cut -c1-1500
tail -4 $O/${TAG}_c2.err
( time timeout 360 python tools/netgen_configs.py c5 --cpu-iters 2 --out $O/${TAG}_c5.json ) 2> $O/${TAG}_c5.err | cut -c1-1500
tail -6 $O/${TAG}_c5.err
