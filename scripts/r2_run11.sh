#!/bin/bash
out=gpurun_out/r2_run11.log
: > $out
for c in 1 2 3; do
echo "== reorth dots CTAs/SM <= $c" >> $out
COLA_REORTH_DOTS_CTAS=$c timeout 600 python scripts/bench_reorth.py 2>&1 | sed 's/; update.*//' >> $out; echo "rc=$?" >> $out
done
echo "== cfg5 (bench_extra)" >> $out
timeout 600 python scripts/bench_extra.py cfg5 >> $out 2>&1; echo "rc=$?" >> $out
echo "== cfg5 launch list (m=32)" >> $out
LANCZOS_M=32 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r2_cfg5_ncu_launches.csv python scripts/bench_extra.py cfg5 > /dev/null 2> gpurun_out/r2_cfg5_ncu.err; echo "rc=$?" >> $out
echo "== cfg4 launch list (m=40, 64 probes)" >> $out
LANCZOS_M=40 PROBE_CHUNK=64 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r2_slq_cfg4_ncu_launches.csv python scripts/bench_extra.py cfg4 --probes 64 > /dev/null 2> gpurun_out/r2_cfg4_ncu.err; echo "rc=$?" >> $out
echo "== cfg3 launch list" >> $out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/r2_cfg3_ncu_launches.csv python scripts/bench_extra.py cfg3 > /dev/null 2> gpurun_out/r2_cfg3_ncu.err; echo "rc=$?" >> $out
