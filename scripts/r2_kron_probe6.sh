#!/bin/bash
out=gpurun_out/r2_kron_probe6.log
: > $out
echo "== check" >> $out
timeout 300 python scripts/check_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== diag" >> $out
DIAG_REPS=6 timeout 400 python scripts/diag_kron3.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== timing default" >> $out
timeout 200 python scripts/time_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== gpu tests" >> $out
timeout 1500 python -m pytest tests -m gpu -x -q >> $out 2>&1; echo "rc=$?" >> $out
