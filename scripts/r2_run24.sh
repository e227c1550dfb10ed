#!/bin/bash
out=gpurun_out/r2_run24.log
: > $out
echo "== smoke" >> $out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $out 2>&1; echo "rc=$?" >> $out
echo "== gpu tests all" >> $out
timeout 1800 python -m pytest tests -m gpu -x -q >> $out 2>&1; echo "rc=$?" >> $out
echo "== next rows" >> $out
timeout 800 python scripts/bench_next_rows.py > gpurun_out/r2_next_rows_late.jsonl 2> gpurun_out/r2_next_rows_late.err; echo "rc=$?" >> $out
echo "== bench" >> $out
timeout 1800 python bench.py > gpurun_out/r2_bench_j.json 2> gpurun_out/r2_bench_j.err; echo "rc=$?" >> $out
echo "== bench reference arm" >> $out
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref2.json 2> gpurun_out/r2_bench_ref2.err; echo "rc=$?" >> $out
tail -c 500 gpurun_out/r2_bench_ref2.json >> $out
