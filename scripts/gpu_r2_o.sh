cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pipelined (default)"; timeout 300 python scripts/pipeline_probe.py 32768 10
echo "== block rows 2048"; FIB_PIPELINE_BLOCK_ROWS=2048 timeout 300 python scripts/pipeline_probe.py 32768 10
echo "== block rows 4096"; FIB_PIPELINE_BLOCK_ROWS=4096 timeout 300 python scripts/pipeline_probe.py 32768 10
echo "== off"; FIB_PIPELINE_MIN_CELLS=99999999999 timeout 300 python scripts/pipeline_probe.py 32768 10
