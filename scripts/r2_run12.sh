#!/bin/bash
out=gpurun_out/r2_run12.log
: > $out
echo "== cfg3 coop" >> $out
timeout 300 python scripts/bench_extra.py cfg3 >> $out 2>&1; echo "rc=$?" >> $out
timeout 200 python scripts/time_kron_tc.py >> $out 2>&1
echo "== cfg3 no coop" >> $out
COLA_KRON_NO_COOP=1 timeout 300 python scripts/bench_extra.py cfg3 >> $out 2>&1; echo "rc=$?" >> $out
COLA_KRON_NO_COOP=1 timeout 200 python scripts/time_kron_tc.py >> $out 2>&1
echo "== reorth dots (caps 2 / 3)" >> $out
timeout 600 python scripts/bench_reorth.py 2>&1 | sed 's/; update.*//' >> $out; echo "rc=$?" >> $out
echo "== cfg5" >> $out
timeout 600 python scripts/bench_extra.py cfg5 >> $out 2>&1; echo "rc=$?" >> $out
echo "== gpu tests: cg workspace + lanczos family" >> $out
timeout 900 python -m pytest tests -m gpu -x -q -k "workspace or lanczos or reorth or slq" >> $out 2>&1; echo "rc=$?" >> $out
