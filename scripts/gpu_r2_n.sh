# round 2, call N: pipelined upload (stepping behind the copies)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipelined_upload.py -q -x --timeout 600 2>&1 | tail -15
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-suite > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err; echo "bench N=1 rc=$?"; tail -3 gpurun_out/r2n_bench_n1.err
python - <<'PY'
import json
for f in ('gpurun_out/r2n_bench_n1.json',):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'value %.1f e2e %.1f (%.3f s) ms/step %.2f frac %.3f reps %d launches %d' % (d['value'], d['e2e']['value'], d['e2e']['seconds'], d['ms_per_step'], d['roofline']['frac'], d['timed_regions']['repeats'], d['gpu_launches']), d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
