#!/bin/bash
out=gpurun_out/r2_run18.log
: > $out
echo "== diag cfg3 cg" >> $out
timeout 300 python scripts/diag_cfg3_cg.py 2>&1 | grep -v Warning | tail -4 >> $out
echo "== gpu tests all" >> $out
timeout 1800 python -m pytest tests -m gpu -x -q >> $out 2>&1; echo "rc=$?" >> $out
echo "== smoke" >> $out
timeout 600 python __graft_entry__.py --smoke >> $out 2>&1; echo "rc=$?" >> $out
echo "== bench" >> $out
timeout 1800 python bench.py > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; echo "rc=$?" >> $out
tail -3 gpurun_out/r2_bench_e.err >> $out
