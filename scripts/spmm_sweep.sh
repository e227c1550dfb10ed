for cfg in "COLA_CSR_PIPE=0" "COLA_CSR_PIPE=1" "COLA_CSR_BATCH=5" "COLA_CSR_BATCH=5 COLA_CSR_MINB=3" "COLA_CSR_BATCH=4 COLA_CSR_MINB=3" "COLA_CSR_RPG=8" "COLA_CSR_NO_PREFETCH=1" "COLA_CSR_RPG=2"; do
  echo "== $cfg"; env $cfg python scripts/bench_spmm.py 2>&1 | grep -E "^spmm|^max err|Error|error"
done
