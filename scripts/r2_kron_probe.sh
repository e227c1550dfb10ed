#!/bin/bash
# v2 fused Kronecker kernel: correctness, then timing against v1 and with the bring-up knobs
out=gpurun_out/r2_kron_probe.log
: > $out
echo "== v2 check" >> $out
timeout 300 python scripts/check_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== v1 check/timing" >> $out
COLA_KRON_V1=1 timeout 300 python scripts/time_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
for cpc in 1 2 4; do
  echo "== v2 cpc=$cpc" >> $out
  COLA_KRON_CPC=$cpc timeout 200 python scripts/time_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
done
for dbg in 1 2 4 8 3 7 15; do
  echo "== v2 cpc=1 dbg=$dbg" >> $out
  COLA_KRON_CPC=1 COLA_KRON_DBG=$dbg timeout 200 python scripts/time_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
done
