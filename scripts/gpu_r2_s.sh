cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_persist.py -q -x --timeout 600 2>&1 | tail -3
for k in 4v br; do timeout 300 python scripts/persist_probe.py $k 300 2>&1 | grep "one fib_step call"; done
