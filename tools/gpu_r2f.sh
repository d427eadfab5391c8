#!/bin/bash
# Round 2, single-GPU call: csr_keep tests + whole suite, bench line (memory of the 111 M-dof system with the CSR arrays released)
TAG=${1:-r2f}
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_csrview.py -m gpu -q -x ) > $O/${TAG}_pytest_csrview.log 2>&1; tail -12 $O/${TAG}_pytest_csrview.log
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/${TAG}_pytest_gpu.log 2>&1; tail -6 $O/${TAG}_pytest_gpu.log
( while true; do nvidia-smi --query-gpu=memory.used --format=csv,noheader,nounits; sleep 1; done ) > $O/${TAG}_mem.log 2>/dev/null &
MEMPID=$!
timeout 900 python bench.py --no-netgen-check > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; cut -c1-600 $O/${TAG}_bench_n1.json; tail -3 $O/${TAG}_bench_n1.err
kill $MEMPID
echo "peak GPU memory (MiB): $(sort -n $O/${TAG}_mem.log | tail -1)"
