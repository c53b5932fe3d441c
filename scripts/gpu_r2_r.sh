cd "$(dirname "$0")/.."
for ct in 1 2; do
  echo "== FIB_PERSIST_CTAS=$ct"
  for k in 4v br; do FIB_PERSIST_CTAS=$ct timeout 300 python scripts/persist_probe.py $k 300 2>&1 | grep "one fib_step call for all"; done
done
FIB_PERSIST_CTAS=2 timeout 600 python -m pytest tests/test_gpu_persist.py -q -x 2>&1 | tail -3
FIB_PERSIST_CTAS=2 FIB_PERSIST_TIMELINE=1 timeout 120 python scripts/persist_probe.py 4v 6 2>&1 | grep -A1 timeline | tail -2 | cut -c1-330
