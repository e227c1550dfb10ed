#!/bin/bash
out=gpurun_out/r2_run13.log
: > $out
echo "== cfg3 default" >> $out
timeout 300 python scripts/bench_extra.py cfg3 >> $out 2>&1
timeout 200 python scripts/time_kron_tc.py >> $out 2>&1
echo "== cfg3 self-reset counters (no memset per launch)" >> $out
COLA_KRON_DBG=64 timeout 300 python scripts/bench_extra.py cfg3 >> $out 2>&1
COLA_KRON_DBG=64 timeout 200 python scripts/time_kron_tc.py >> $out 2>&1
echo "== gpu tests all" >> $out
timeout 1500 python -m pytest tests -m gpu -x -q >> $out 2>&1; echo "rc=$?" >> $out
