#!/bin/bash
# fused Kronecker kernels: correctness, then timing per generation and with the bring-up knobs
out=gpurun_out/r2_kron_probe.log
: > $out
echo "== v3 check" >> $out
timeout 300 python scripts/check_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
for cpc in 1 2 4; do
  echo "== v3 cpc=$cpc" >> $out
  COLA_KRON_CPC=$cpc timeout 200 python scripts/time_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
done
echo "== v3 cpc=1 no pairs" >> $out
COLA_KRON_NO_PAIRS=1 timeout 200 python scripts/time_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
for dbg in 1 2 4 8 3 7 15; do
  echo "== v3 cpc=1 dbg=$dbg" >> $out
  COLA_KRON_DBG=$dbg timeout 200 python scripts/time_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
done
echo "== v2 cpc=4" >> $out
COLA_KRON_GEN=2 timeout 200 python scripts/time_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
