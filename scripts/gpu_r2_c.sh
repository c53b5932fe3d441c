# round 2, call C: persistent kernel after the polling fix, tuned defaults, bench smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== persistent kernel tests"
timeout 900 python -m pytest tests/test_gpu_persist.py -q --timeout 300 2>&1 | tail -8
echo "== 512^2 configs: persistent vs one launch per step"
python scripts/persist_probe.py 4v 100; python scripts/persist_probe.py br 100
FIB_PERSIST=0 python scripts/persist_probe.py 4v 100; FIB_PERSIST=0 python scripts/persist_probe.py br 100
python scripts/persist_probe.py 4v 20 > gpurun_out/pp.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_persist_launches.csv python scripts/persist_probe.py 4v 20 > gpurun_out/ncu_pp.log 2>&1
grep -c persist gpurun_out/r2_persist_launches.csv; grep persist gpurun_out/r2_persist_launches.csv | tail -4 | cut -c1-300
echo "== tuned defaults (4096^2)"
for k in br br_exact br_skip court court_ultra court_lut; do python tests/quick_perf.py $k 4096 6 2>&1 | tail -1; done
echo "== full GPU tests"
timeout 2400 python -m pytest tests -m gpu -q -rf --timeout 1200 > gpurun_out/r2c_tests.log 2>&1; echo "pytest rc=$?"
grep -E "^E  .*(Error|assert)|^FAILED|passed|failed" gpurun_out/r2c_tests.log | head -30
echo "== bench smoke (8192^2)"
python bench.py --size 8192 --steps 5 --warmup 3 > gpurun_out/r2c_bench8192.json 2> gpurun_out/r2c_bench8192.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c_bench8192.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench8192.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'dram_frac',d['roofline'].get('dram_frac'),'reps',d['timed_regions']['repeats'])
for s in d.get('suite',[]): print('  %-60s %8.2f  frac %.3f  %s'%(s['case'],s.get('value',0),s.get('frac',0),s.get('kernel','')[:60]))
print(d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('numpy_restatement',{}).get('value'))
PY
