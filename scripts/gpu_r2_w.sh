cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/w_$label.err | tail -1 > gpurun_out/w_$label.json
  python -c "import json;d=json.load(open('gpurun_out/w_$label.json'));print('$label: value %.0f e2e %.0f (%.3f s)'%(d['value'],d['e2e']['value'],d['e2e']['seconds']))"
}
run nopipe FIB_PIPELINE_NCCL=0
run default FIB_X=1
run blocks512 FIB_PIPELINE_BLOCK_ROWS=512
run blocks2048 FIB_PIPELINE_BLOCK_ROWS=2048
