#!/bin/bash
out=gpurun_out/r2_run21.log
: > $out
echo "== gpu tests all" >> $out
timeout 1800 python -m pytest tests -m gpu -x -q >> $out 2>&1; echo "rc=$?" >> $out
echo "== cfg2 launch list" >> $out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/r2_cg_cfg2_ncu_launches_staged.csv python bench.py --workload cfg2 --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2> gpurun_out/r2_cfg2_ncu.err; echo "rc=$?" >> $out
echo "== cfg2 staged spmm full" >> $out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:csr_spmm_tiled -s 40 -c 1 -o gpurun_out/r2_spmm_staged_full -f python bench.py --workload cfg2 --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_cfg2_ncu2.err; echo "rc=$?" >> $out
echo "== bench" >> $out
timeout 1800 python bench.py > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err; echo "rc=$?" >> $out
tail -c 400 gpurun_out/r2_bench_g.json >> $out
