cd "$(dirname "$0")/.."
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader
nvidia-smi topo -m | head -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/dist_parity.py 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --size 8192 --steps 5 --warmup 3 2>gpurun_out/bench_n${N}_8192.err | tail -1 | tee gpurun_out/bench_n${N}_8192.json; tail -5 gpurun_out/bench_n${N}_8192.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N 2>gpurun_out/bench_n${N}.err | tail -1 | tee gpurun_out/bench_n${N}.json; tail -5 gpurun_out/bench_n${N}.err
python bench.py --no-cpu 2>gpurun_out/bench_n1.err | tail -1 | tee gpurun_out/bench_n1.json
