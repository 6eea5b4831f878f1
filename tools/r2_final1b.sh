# refresh of the single-GPU bench lines and the launch list with the final build (deferred batches)
set -x
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02_bench_steps2.json 2> /dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1
python bench.py --precision fp64 --no-configs --no-cpu-baseline > gpurun_out/r02_bench_fp64.json 2> gpurun_out/r02_bench_fp64.err
tail -c 200 gpurun_out/r02_bench.err
