cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/x_n8.err | tail -1 > gpurun_out/x_n8.json
python -c "import json;d=json.load(open('gpurun_out/x_n8.json'));print('N=8: value %.0f e2e %.0f (%.3f s)'%(d['value'],d['e2e']['value'],d['e2e']['seconds']))"
grep elapsed gpurun_out/x_n8.err | tr '\n' ' '
