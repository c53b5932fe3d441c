# round 2, call M: lazy Laplacian in the persistent kernel's edge rows; probe ring in page-locked memory
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_persist.py tests/test_gpu_ops_observers.py tests/test_c_example.py -q -x --timeout 600 2>&1 | tail -8
for k in 4v br; do timeout 300 python scripts/persist_probe.py $k 300 2>&1 | grep -v "^$\|elapsed"; done
FIB_PERSIST_TIMELINE=1 timeout 120 python scripts/persist_probe.py 4v 6 2>&1 | grep timeline | tail -3
FIB_PERSIST_TIMELINE=1 timeout 120 python scripts/persist_probe.py br 6 2>&1 | grep timeline | tail -3
