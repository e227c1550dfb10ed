#!/bin/bash
out=gpurun_out/r2_run10.log
: > $out
echo "== reorth (qsplit default)" >> $out
timeout 600 python scripts/bench_reorth.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== reorth (qsplit 1)" >> $out
COLA_REORTH_QSPLIT=1 timeout 600 python scripts/bench_reorth.py 2>&1 | sed 's/; update.*//' >> $out; echo "rc=$?" >> $out
echo "== gpu tests (lanczos family)" >> $out
timeout 900 python -m pytest tests -m gpu -x -q -k "lanczos or reorth or slq or eig or cfg4 or cfg5 or hutch or unary" >> $out 2>&1; echo "rc=$?" >> $out
echo "== ncu launch list, cfg2 bench" >> $out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r2_cg_cfg2_ncu_launches.csv python bench.py --workload cfg2 --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2_ncu_bench.json 2> gpurun_out/r2_ncu_bench.err; echo "rc=$?" >> $out
echo "== ncu full: spmm + sweeps" >> $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csr_spmm_pipe|sweep_kernel" -s 40 -c 6 -o gpurun_out/r2_cg_cfg2_full python bench.py --workload cfg2 --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_ncu_full.err; echo "rc=$?" >> $out
