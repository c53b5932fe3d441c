# GPU round A: tests + bench.  Output under gpurun_out/ (keep it small: 64 MiB cap).
set -x
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python bench.py --size 8192 --steps 5 --warmup 3 > gpurun_out/bench_8192.json 2> gpurun_out/bench_8192.err; tail -3 gpurun_out/bench_8192.err; cat gpurun_out/bench_8192.json
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
python bench.py --impl reference --steps 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
