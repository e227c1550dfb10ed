#!/bin/bash
out=gpurun_out/r2_run14.log
: > $out
echo "== mode_tc check" >> $out
timeout 600 python scripts/check_mode_tc.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== cfg4 (bench_extra, 128 probes)" >> $out
timeout 900 python scripts/bench_extra.py cfg4 >> $out 2>&1; echo "rc=$?" >> $out
echo "== gpu tests (kron / slq / lanczos / matmat)" >> $out
timeout 1200 python -m pytest tests -m gpu -x -q -k "kron or slq or lanczos or matmat or cfg4 or mode_contract or unary or hutch" >> $out 2>&1; echo "rc=$?" >> $out
