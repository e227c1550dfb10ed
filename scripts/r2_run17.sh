#!/bin/bash
out=gpurun_out/r2_run17.log
: > $out
echo "== check kron fused" >> $out
timeout 300 python scripts/check_kron_tc.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== mode_tc check" >> $out
timeout 600 python scripts/check_mode_tc.py 2>&1 | tail -13 >> $out; echo "rc=$?" >> $out
echo "== gpu tests (kron / tensor core / mode / blockdiag / matmat)" >> $out
timeout 1200 python -m pytest tests -m gpu -x -q -k "kron or tensor_core or mode_contract or full_size or cfg3 or cfg4 or matmat or blockdiag" >> $out 2>&1; echo "rc=$?" >> $out
echo "== next rows" >> $out
timeout 900 python scripts/bench_next_rows.py 2> gpurun_out/r2_next_rows.err | head -2 >> $out
