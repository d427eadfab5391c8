#!/bin/bash
# single GPU: the whole GPU suite (with the batched GMRES orthogonalisation and its distributed form, two/three ranks on one GPU)
TAG=${1:-r2u}
O=gpurun_out; mkdir -p $O
( time timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q -x -k "gmres" ) > $O/${TAG}_pytest_distgmres.log 2>&1; tail -4 $O/${TAG}_pytest_distgmres.log
( time timeout 600 python -m pytest tests -m gpu -q -x ) > $O/${TAG}_pytest_gpu.log 2>&1; tail -6 $O/${TAG}_pytest_gpu.log
