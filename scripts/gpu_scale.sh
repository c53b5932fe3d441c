# 8-GPU box: NCCL bit-identity at 8 ranks, then the bench at N = 8, 4, 2, 1 exactly as the driver launches it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tests/dist_parity.py 2>&1 | grep -E "ranks|OK|MISMATCH|Error|error" | head
for N in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/scale_n$N.err > gpurun_out/scale_n$N.out
  grep '^{' gpurun_out/scale_n$N.out | tail -1 > gpurun_out/scale_n$N.json
  echo "N=$N stdout lines: $(wc -l < gpurun_out/scale_n$N.out)"
  python -c "import json;d=json.load(open('gpurun_out/scale_n$N.json'));print($N,d['value'],d['e2e']['value'],d['e2e']['seconds'],d['roofline']['frac'],d['clocks'])"
done
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/scale_n1.err | tail -1 > gpurun_out/scale_n1.json
python -c "import json;d=json.load(open('gpurun_out/scale_n1.json'));print(1,d['value'],d['e2e']['value'],d['e2e']['seconds'],d['roofline']['frac'],d['clocks'])"
