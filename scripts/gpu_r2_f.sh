# round 2, call F: packed == scalar with the opaque multiply, persistent kernel with concurrent polling,
# LUT multi-step diagnosis, full tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== diag: packed (default) vs scalar build, TWO time steps"
( export FIB_PERSIST=0 FIB_DEBUG_SUBSTEPS=2
  FIB_SMALL_CELLS=0 python tests/diag_packed.py dump gpurun_out/diag_packed.npz
  FIB_SMALL_CELLS=0 FIB_B200_LIB=$PWD/build/variants/lib_scalarall.so python tests/diag_packed.py dump gpurun_out/diag_scalar.npz
  python tests/diag_packed.py cmp gpurun_out/diag_packed.npz gpurun_out/diag_scalar.npz | grep -v kernel; echo "(end of differences)" )
echo "== persistent kernel"
timeout 300 python scripts/persist_probe.py 4v 100; timeout 300 python scripts/persist_probe.py br 100
FIB_PERSIST=0 python scripts/persist_probe.py 4v 100; FIB_PERSIST=0 python scripts/persist_probe.py br 100
timeout 900 python -m pytest tests/test_gpu_persist.py -q --timeout 300 2>&1 | tail -6
python scripts/persist_probe.py 4v 20 > gpurun_out/pp.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_persist_launches.csv python scripts/persist_probe.py 4v 20 > gpurun_out/ncu_pp.log 2>&1
grep persist gpurun_out/r2_persist_launches.csv | tail -3 | awk -F'","' '{print $5, $NF}'
echo "== LUT flavours vs the compiled reference header, step by step"
python tests/diag_lut.py; FIB_SMALL_CELLS=100000000 python tests/diag_lut.py
echo "== full GPU tests"
timeout 2400 python -m pytest tests -m gpu -q -rf --timeout 1200 > gpurun_out/r2f_tests.log 2>&1; echo "pytest rc=$?"
grep -E "^E  .*(Error|assert)|^FAILED|passed|failed" gpurun_out/r2f_tests.log | cut -c1-220 | head -40
for k in br br_exact br_skip court court_ultra court_lut; do python tests/quick_perf.py $k 4096 6 2>&1 | tail -1; done
python tests/quick_perf.py 4v 8192 10 | tail -1
