cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tests/dist_parity.py 2>&1 | grep -E "ranks|OK|MISMATCH|Error|error" | head
for N in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/scale_n$N.err | tail -1 > gpurun_out/scale_n$N.json
  python -c "import json;d=json.load(open('gpurun_out/scale_n$N.json'));print($N,d['value'],d['e2e']['value'],d['roofline']['frac'],d['clocks'])"
done
python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/scale_n1.err | tail -1 > gpurun_out/scale_n1.json
python -c "import json;d=json.load(open('gpurun_out/scale_n1.json'));print(1,d['value'],d['e2e']['value'],d['roofline']['frac'],d['clocks'])"
