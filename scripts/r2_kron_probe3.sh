#!/bin/bash
out=gpurun_out/r2_kron_probe3.log
: > $out
echo "== v3 hints on: diag" >> $out
DIAG_REPS=6 timeout 400 python scripts/diag_kron3.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== v3 hints on: timing" >> $out
timeout 200 python scripts/time_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== v3 hints off: timing" >> $out
COLA_KRON_DBG=32 timeout 200 python scripts/time_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== v3 hints on, no pairs: timing" >> $out
COLA_KRON_NO_PAIRS=1 timeout 200 python scripts/time_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== v3 hints on, cpc 2: timing" >> $out
COLA_KRON_CPC=2 timeout 200 python scripts/time_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== v3 prof (hints on)" >> $out
COLA_KRON_PROF=1 timeout 200 python scripts/run_kron_tc_once.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== v3 prof (hints on, no stores)" >> $out
COLA_KRON_DBG=4 COLA_KRON_PROF=1 timeout 200 python scripts/run_kron_tc_once.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== v3 ncu hints on" >> $out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:kron_fused3 -s 2 -c 2 --csv --log-file gpurun_out/r2_kron3_ncu_metrics_hints.csv python scripts/run_kron_tc_once.py >> $out 2>&1; echo "rc=$?" >> $out
