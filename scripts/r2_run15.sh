#!/bin/bash
out=gpurun_out/r2_run15.log
: > $out
echo "== mode_tc check" >> $out
timeout 600 python scripts/check_mode_tc.py 2>&1 | tail -5 >> $out; echo "rc=$?" >> $out
echo "== cfg3 matmat" >> $out
timeout 200 python scripts/time_kron_tc.py >> $out 2>&1
echo "== gpu tests (kron / tensor core / mode)" >> $out
timeout 1200 python -m pytest tests -m gpu -x -q -k "kron or tensor_core or mode_contract or full_size or cfg3 or cfg4" >> $out 2>&1; echo "rc=$?" >> $out
echo "== bench" >> $out
timeout 1800 python bench.py > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; echo "rc=$?" >> $out
tail -3 gpurun_out/r2_bench_d.err >> $out
