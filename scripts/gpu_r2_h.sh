# round 2, call H: persistent kernel with the enforced-field shared tile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_persist.py -q --timeout 300 2>&1 | tail -8
echo "== timeline 4v / BR (ns)"
FIB_PERSIST_TIMELINE=1 python scripts/persist_probe.py 4v 3 2>&1 | tail -3
FIB_PERSIST_TIMELINE=1 python scripts/persist_probe.py br 3 2>&1 | tail -3
echo "== rates"
python scripts/persist_probe.py 4v 300; python scripts/persist_probe.py br 300
FIB_PERSIST=0 python scripts/persist_probe.py 4v 300; FIB_PERSIST=0 python scripts/persist_probe.py br 300
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_smallgrids.py tests/test_c_example.py -q --timeout 600 2>&1 | tail -5
