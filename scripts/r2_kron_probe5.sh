#!/bin/bash
out=gpurun_out/r2_kron_probe5.log
: > $out
echo "== v4 diag" >> $out
DIAG_REPS=8 timeout 400 python scripts/diag_kron3.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== v4 check" >> $out
timeout 300 python scripts/check_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== v4 tests" >> $out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_next.py -m gpu -x -q -k "kron or tensor_core or full_size" >> $out 2>&1; echo "rc=$?" >> $out
for cpc in 1 2; do
echo "== v4 timing cpc=$cpc" >> $out
COLA_KRON_CPC=$cpc timeout 200 python scripts/time_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
done
echo "== v4 timing no pairs" >> $out
COLA_KRON_NO_PAIRS=1 timeout 200 python scripts/time_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
for dbg in 1 2 4 7; do
echo "== v4 timing dbg=$dbg" >> $out
COLA_KRON_DBG=$dbg timeout 200 python scripts/time_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
done
echo "== v4 prof" >> $out
COLA_KRON_PROF=1 timeout 200 python scripts/run_kron_tc_once.py 2>&1 | tail -4 >> $out
echo "== v4 ncu" >> $out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:kron_fused4 -s 2 -c 2 --csv --log-file gpurun_out/r2_kron4_ncu_metrics.csv python scripts/run_kron_tc_once.py >> $out 2>&1; echo "rc=$?" >> $out
