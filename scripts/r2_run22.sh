#!/bin/bash
out=gpurun_out/r2_run22.log
: > $out
for u in 1 3 4 5; do
  COLA_LIB=scripts/_variants/libcola_u$u.so timeout 200 python scripts/prof_spmm_tiled.py 2>&1 | grep staged | sed "s/^/unroll $u: /" >> $out
done
timeout 200 python scripts/prof_spmm_tiled.py 2>&1 | grep staged | sed "s/^/unroll 2 (shipped): /" >> $out
