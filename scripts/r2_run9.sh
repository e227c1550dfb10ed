#!/bin/bash
out=gpurun_out/r2_run9.log
: > $out
echo "== spmm (1-D tiles, reverted)" >> $out
timeout 300 python scripts/bench_spmm.py >> $out 2>&1; echo "rc=$?" >> $out
for mb in 34 45 68; do
echo "== spmv strips $mb MB" >> $out
COLA_SPMV_BLOCK_MB=$mb timeout 600 python scripts/bench_spmv.py 2>&1 | grep "k=" >> $out; echo "rc=$?" >> $out
done
echo "== bench" >> $out
timeout 1800 python bench.py > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; echo "rc=$?" >> $out
tail -3 gpurun_out/r2_bench_c.err >> $out
