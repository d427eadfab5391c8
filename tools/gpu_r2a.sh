#!/bin/bash
# Round 2, first GPU call (one B200): new reorder tests + whole GPU suite, netgen-numbered systems at 1.7 M and 13.9 M dofs
# (natural numbering vs the library's Cuthill-McKee reordering), DRAM traffic of the SpMV in both, bench sanity.
TAG=${1:-r2a}
O=gpurun_out; mkdir -p $O
( nvidia-smi --query-gpu=name,memory.total --format=csv; nproc; free -g | head -2; df -h /dev/shm | tail -1 ) > $O/${TAG}_box.txt 2>&1; cat $O/${TAG}_box.txt
( time timeout 600 python -m pytest tests/test_gpu_reorder.py -m gpu -q -x ) > $O/${TAG}_pytest_reorder.log 2>&1; tail -15 $O/${TAG}_pytest_reorder.log
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/${TAG}_pytest_gpu.log 2>&1; tail -5 $O/${TAG}_pytest_gpu.log
timeout 600 python tools/netgen_scale.py --nref 1 --cache /dev/shm/ng1 --cpu-full --full --out $O/${TAG}_netgen_1p7M.json > /dev/null 2> $O/${TAG}_netgen_1p7M.err; cut -c1-1500 $O/${TAG}_netgen_1p7M.json; tail -3 $O/${TAG}_netgen_1p7M.err
timeout 900 python tools/netgen_scale.py --nref 2 --cache /dev/shm/ng2 --full --out $O/${TAG}_netgen_14M.json > /dev/null 2> $O/${TAG}_netgen_14M.err; cut -c1-2500 $O/${TAG}_netgen_14M.json; tail -3 $O/${TAG}_netgen_14M.err
for mode in 0 1; do
  timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:sell_spmv -s 3 -c 2 --csv \
      --log-file $O/${TAG}_ncu_netgen14M_reorder${mode}.csv python tools/netgen_scale.py --nref 2 --cache /dev/shm/ng2 --modes $mode --spmv-only --reps 3 > /dev/null 2>&1
  tail -8 $O/${TAG}_ncu_netgen14M_reorder${mode}.csv | cut -c1-400
done
timeout 300 python bench.py --no-cpu-baseline > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; cut -c1-400 $O/${TAG}_bench_n1.json; tail -3 $O/${TAG}_bench_n1.err
ls -la $O | tail -12
