#!/bin/bash
out=gpurun_out/r2_run7.log
: > $out
echo "== backward gpu tests" >> $out
timeout 600 python -m pytest tests/test_backward.py -m gpu -x -q >> $out 2>&1; echo "rc=$?" >> $out
echo "== bench" >> $out
timeout 1500 python bench.py > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; echo "rc=$?" >> $out
tail -3 gpurun_out/r2_bench_c.err >> $out
