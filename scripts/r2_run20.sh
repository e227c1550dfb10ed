#!/bin/bash
out=gpurun_out/r2_run20.log
: > $out
echo "== gpu tests all" >> $out
timeout 1800 python -m pytest tests -m gpu -x -q >> $out 2>&1; echo "rc=$?" >> $out
echo "== cfg4 launch list (m=40, 64 probes)" >> $out
LANCZOS_M=40 PROBE_CHUNK=64 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r2_slq_cfg4_ncu_launches.csv python scripts/bench_extra.py cfg4 --probes 64 > /dev/null 2> gpurun_out/r2_cfg4_ncu.err; echo "rc=$?" >> $out
echo "== mode_tc ncu metrics" >> $out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread --clock-control none -k regex:mode_tc_kernel -s 30 -c 8 --csv --log-file gpurun_out/r2_mode_tc_ncu_metrics.csv python scripts/check_mode_tc.py > /dev/null 2>&1; echo "rc=$?" >> $out
echo "== bench" >> $out
timeout 1800 python bench.py > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err; echo "rc=$?" >> $out
echo "== bench reference arm" >> $out
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "rc=$?" >> $out
tail -c 600 gpurun_out/r2_bench_ref.json >> $out
