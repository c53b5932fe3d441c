# round 2, call L: deferred persistent iterations + in-kernel probe ring; fused-kernel variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_persist.py tests/test_gpu_ops_observers.py tests/test_c_example.py -q -x --timeout 600 2>&1 | tail -8
for k in 4v br; do timeout 300 python scripts/persist_probe.py $k 300 2>&1 | grep -v "^$"; done
echo "== fused 4v variants (8192^2, 100 steps)"
for v in "" fuse4 fuse6 fuseu2 fusepf2; do
  echo "-- variant: ${v:-default}"
  FIB_B200_LIB=${v:+build/variants/lib_$v.so} timeout 300 python tests/quick_perf.py 4v 8192 100 2>&1 | tail -1
done
