set -x
python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -x -k "round_robin" 2>&1 | tail -4 > gpurun_out/r2_t7.log
MAICOS_B200_TRACE_HOST=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 --configs C5 --no-e2e > gpurun_out/r2_c5_8.json 2> gpurun_out/r2_c5_8.err
cat gpurun_out/r2_t7.log
