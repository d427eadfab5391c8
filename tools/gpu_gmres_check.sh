#!/bin/bash
# GMRES / complex-kernel check: parity tests that touch GMRES, config 5 numbers, ncu capture of the complex SELL kernel
TAG=${1:-r1}; O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q -k "gmres or GMRES or helmholtz or complex" ) > $O/${TAG}_pytest_gmres.log 2>&1; tail -3 $O/${TAG}_pytest_gmres.log
timeout 600 python tools/bench_configs.py c5 c1 --out $O/${TAG}_configs_c5.jsonl > $O/${TAG}_configs_c5.log 2>&1; tail -2 $O/${TAG}_configs_c5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sell_spmv -s 3 -c 1 -f -o $O/${TAG}_prof_sell_complex \
    python tools/bench_configs.py c5 --scale 0.3 > $O/${TAG}_ncu_complex.log 2>&1
ncu -i $O/${TAG}_prof_sell_complex.ncu-rep --page raw --csv > $O/${TAG}_prof_sell_complex_raw.csv 2>/dev/null
ls -la $O | tail -5
