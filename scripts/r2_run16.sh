#!/bin/bash
out=gpurun_out/r2_run16.log
: > $out
echo "== gpu tests all" >> $out
timeout 1800 python -m pytest tests -m gpu -x -q >> $out 2>&1; echo "rc=$?" >> $out
echo "== next rows bench (KronSum etc.)" >> $out
timeout 900 python scripts/bench_next_rows.py > gpurun_out/r2_next_rows.jsonl 2> gpurun_out/r2_next_rows.err; echo "rc=$?" >> $out
cat gpurun_out/r2_next_rows.jsonl >> $out
