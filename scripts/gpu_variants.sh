cd "$(dirname "$0")/.."
for k in 4v br br_exact br_skip court_ultra; do python tests/quick_perf.py $k 512 200 | tail -1; done
python tests/quick_perf.py br 1024 50 | tail -1; python tests/quick_perf.py br 1536 50 | tail -1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --no-cpu 2>/dev/null | tail -1 > gpurun_out/bench_final_n1.json; python -c "import json;d=json.load(open('gpurun_out/bench_final_n1.json'));print(d['value'],d['e2e'],d['roofline']['frac'],d['clocks'])"
