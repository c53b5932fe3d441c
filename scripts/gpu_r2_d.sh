# round 2, call D: packed-vs-scalar diagnosis, persistent kernel after the fence fix, LUT smem A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== diag: packed (default) vs scalar build, ONE time step, step kernels"
export FIB_PERSIST=0 FIB_DEBUG_SUBSTEPS=1
FIB_SMALL_CELLS=0 python tests/diag_packed.py dump gpurun_out/diag_packed.npz
FIB_SMALL_CELLS=0 FIB_B200_LIB=$PWD/build/variants/lib_scalarall.so python tests/diag_packed.py dump gpurun_out/diag_scalar.npz
python tests/diag_packed.py dump gpurun_out/diag_small.npz
python tests/diag_packed.py cmp gpurun_out/diag_packed.npz gpurun_out/diag_scalar.npz
echo "-- small (VEC=1) vs scalar wide"
python tests/diag_packed.py cmp gpurun_out/diag_small.npz gpurun_out/diag_scalar.npz | grep -v kernel
unset FIB_PERSIST FIB_DEBUG_SUBSTEPS
echo "== persistent kernel"
python scripts/persist_probe.py 4v 100; python scripts/persist_probe.py br 100
FIB_PERSIST=0 python scripts/persist_probe.py 4v 100; FIB_PERSIST=0 python scripts/persist_probe.py br 100
timeout 600 python -m pytest tests/test_gpu_persist.py -q --timeout 300 2>&1 | tail -4
echo "== LUT in shared memory A/B, all-state MINB 6, 4v scalar"
for sz in 2048 4096; do python tests/quick_perf.py court_lut $sz 6 | tail -1; for v in lutsmem lutsmem4; do FIB_B200_LIB=$PWD/build/variants/lib_$v.so python tests/quick_perf.py court_lut $sz 6 | tail -1; done; done
FIB_B200_LIB=$PWD/build/variants/lib_all2_6.so python tests/quick_perf.py court_ultra 4096 6 | tail -1
python tests/quick_perf.py 4v 8192 10 | tail -1; FIB_B200_LIB=$PWD/build/variants/lib_v4scalar.so python tests/quick_perf.py 4v 8192 10 | tail -1
FIB_STEPS_PER_LAUNCH=1 python tests/quick_perf.py 4v 8192 10 | tail -1; FIB_STEPS_PER_LAUNCH=1 FIB_B200_LIB=$PWD/build/variants/lib_v4scalar.so python tests/quick_perf.py 4v 8192 10 | tail -1
timeout 600 python -m pytest tests/test_gpu_wide_flavours.py -q -k "lut_wide" 2>&1 | tail -3
