# 2 GPUs: NCCL bit-identity incl. the pipelined upload behind NCCL shards, then the N=2 bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/dist_parity.py > gpurun_out/dist2.log 2>&1; echo "dist_parity rc=$?"
grep -E "ranks|OK|MISMATCH|Error|error|window" gpurun_out/dist2.log | head -20; tail -5 gpurun_out/dist2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/v_n2.err > gpurun_out/v_n2.json; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/v_n2.json'));print(2,d['value'],d['e2e']['value'],d['e2e']['seconds'])"; tail -3 gpurun_out/v_n2.err
