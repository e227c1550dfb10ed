#!/bin/bash
out=gpurun_out/r2_run19.log
: > $out
for mode in clocks noclocks clocks noclocks; do
  if [ $mode = noclocks ]; then export COLA_BENCH_NO_CLOCKS=1; else unset COLA_BENCH_NO_CLOCKS; fi
  timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_tmp.json 2> /dev/null
  python - >> $out <<PY
import json
d=json.load(open("gpurun_out/r2_bench_tmp.json"))
s=d["secondary"]
print("$mode", "value %.1f e2e %.1f cfg3 %.0f it/s (%.4f ms) slq %.2f s cfg5 %.3f s" % (d["value"], d["e2e"]["value"], s["cfg3"]["iters_per_s"], s["cfg3"]["ms_per_iter"], s["slq"]["seconds"], s["cfg5"]["seconds"]))
PY
done
