# Round-2 measurement pass on the 8-GPU box
set -x
python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -x -k "round_robin" 2>&1 | tail -3 > gpurun_out/r02_multidevice_test.log
for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_${n}gpu.json 2> gpurun_out/r02_bench_${n}gpu.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus 8 --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_8gpu.json 2> gpurun_out/r02_bench_reference_8gpu.err
python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/r02_bench_1gpu_on8box.json 2> /dev/null
python tools/r2_devices_demo.py > gpurun_out/r02_use_devices.json 2> gpurun_out/r02_use_devices.err
cat gpurun_out/r02_multidevice_test.log; tail -c 300 gpurun_out/r02_bench_8gpu.err; tail -c 300 gpurun_out/r02_use_devices.err
