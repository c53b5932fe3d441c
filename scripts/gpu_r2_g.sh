# round 2, call G: persistent kernel timeline (where do 3 us per step go?), then perf + tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== persistent kernel timeline, 4v 512^2 (ns per phase, middle tile)"
FIB_PERSIST_TIMELINE=1 python scripts/persist_probe.py 4v 6 2>&1 | tail -8
echo "== BR"
FIB_PERSIST_TIMELINE=1 python scripts/persist_probe.py br 4 2>&1 | tail -5
echo "== persistent kernel rates"
python scripts/persist_probe.py 4v 200; python scripts/persist_probe.py br 200
FIB_PERSIST=0 python scripts/persist_probe.py 4v 200; FIB_PERSIST=0 python scripts/persist_probe.py br 200
timeout 900 python -m pytest tests/test_gpu_persist.py tests/test_gpu_wide_flavours.py -q --timeout 600 2>&1 | tail -5
for k in court_ultra court br; do python tests/quick_perf.py $k 4096 6 2>&1 | tail -1; done
python tests/quick_perf.py 4v 4096 10 --phase | tail -1
FIB_STEPS_PER_LAUNCH=1 python tests/quick_perf.py 4v 4096 10 --phase | tail -1
