set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2_t6.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2_ref8.json 2> gpurun_out/r2_ref8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench2g.json 2> gpurun_out/r2_bench2g.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 --configs C5 > gpurun_out/r2_bench4g.json 2> gpurun_out/r2_bench4g.err
cat gpurun_out/r2_t6.log; tail -c 400 gpurun_out/r2_bench8.err; tail -c 300 gpurun_out/r2_ref8.err
