# round 2, call K (2 GPUs): multi-GPU tests inside pytest, bench at N=2 (e2e with async uploads)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_multi.py -q --timeout 1000 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err; echo "bench N=2 rc=$?"; tail -3 gpurun_out/r2k_bench_n2.err
python bench.py --gpus 1 --steps 20 --warmup 5 --no-suite > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; echo "bench N=1 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2k_bench_n1.json','gpurun_out/r2k_bench_n2.json'):
    try:
        d=json.load(open(f))
        print(f, 'value %.1f e2e %.1f (%.3f s) ms/step %.2f frac %.3f reps %d launches %d' % (d['value'], d['e2e']['value'], d['e2e']['seconds'], d['ms_per_step'], d['roofline']['frac'], d['timed_regions']['repeats'], d['gpu_launches']), d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
