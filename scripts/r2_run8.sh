#!/bin/bash
out=gpurun_out/r2_run8.log
: > $out
echo "== spmm strips (default)" >> $out
timeout 300 python scripts/bench_spmm.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== spmm no strips" >> $out
COLA_CSR_NO_STRIPS=1 timeout 300 python scripts/bench_spmm.py >> $out 2>&1; echo "rc=$?" >> $out
for rpg in 4 16; do
echo "== spmm strips rpg=$rpg" >> $out
COLA_CSR_RPG=$rpg timeout 300 python scripts/bench_spmm.py >> $out 2>&1; echo "rc=$?" >> $out
done
echo "== spmm ragged grid 1000 (strips + tail tiles)" >> $out
GRID=1000 timeout 300 python scripts/bench_spmm.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== backward + strips + csr gpu tests" >> $out
timeout 900 python -m pytest tests/test_backward.py tests/test_gpu_parity_next.py tests/test_gpu_parity.py tests/test_gpu_fullscale.py -m gpu -x -q -k "backward or param_grad or strips or matmat or cg or cfg2 or cfg5" >> $out 2>&1; echo "rc=$?" >> $out
echo "== spmv strips (default 32 MB)" >> $out
timeout 600 python scripts/bench_spmv.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== spmv strips 16 MB" >> $out
COLA_SPMV_BLOCK_MB=16 timeout 600 python scripts/bench_spmv.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== spmv no strips" >> $out
COLA_SPMV_BLOCK_MB=0 timeout 600 python scripts/bench_spmv.py >> $out 2>&1; echo "rc=$?" >> $out
echo "== reorth" >> $out
timeout 600 python scripts/bench_reorth.py 2>&1 | grep "b=1 " >> $out; echo "rc=$?" >> $out
COLA_REORTH_NO_FOLD=1 timeout 600 python scripts/bench_reorth.py 2>&1 | grep "b=1 " >> $out; echo "rc=$?" >> $out
