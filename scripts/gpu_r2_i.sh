# round 2, call I: persistent kernel: late first poll; cooperative vs plain launch gap
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_persist.py -q --timeout 300 2>&1 | tail -3
echo "== timeline 4v / BR (ns)"
FIB_PERSIST_TIMELINE=1 python scripts/persist_probe.py 4v 3 2>&1 | tail -2
FIB_PERSIST_TIMELINE=1 python scripts/persist_probe.py br 3 2>&1 | tail -2
echo "== rates (cooperative launch)"
python scripts/persist_probe.py 4v 300; python scripts/persist_probe.py br 300
echo "== rates (plain launch, experiment)"
FIB_PERSIST_COOP=0 timeout 120 python scripts/persist_probe.py 4v 300; FIB_PERSIST_COOP=0 timeout 120 python scripts/persist_probe.py br 300
