#!/bin/bash
out=gpurun_out/r2_kron_probe2.log
: > $out
echo "== v3 diag (slot held through tcgen05.st)" >> $out
timeout 400 python scripts/diag_kron3.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== v3 diag cpc 2" >> $out
COLA_KRON_CPC=2 timeout 400 python scripts/diag_kron3.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== v2 diag" >> $out
COLA_KRON_GEN=2 DIAG_REPS=6 timeout 400 python scripts/diag_kron3.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== v3 timing" >> $out
timeout 200 python scripts/time_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== v3 ncu" >> $out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum --clock-control none -k regex:kron_fused3 -s 2 -c 2 --csv --log-file gpurun_out/r2_kron3_ncu_metrics.csv python scripts/run_kron_tc_once.py >> $out 2>&1; echo "rc=$?" >> $out
