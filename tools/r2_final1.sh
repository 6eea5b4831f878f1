# Round-2 measurement pass on one B200 (run through gpurun; outputs land in gpurun_out/ and are copied to profiles/)
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_gpu_tests.log
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
python bench.py --precision fp64 --no-configs --no-cpu-baseline > gpurun_out/r02_bench_fp64.json 2> gpurun_out/r02_bench_fp64.err
python tools/r2_npt_probe.py 33333 256 > gpurun_out/r02_npt.log 2>&1
python tools/r2_unbinned.py 33333 64 > gpurun_out/r02_unbinned.json 2> /dev/null
python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02_bench_steps2.json 2> /dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sf_contract_tc -c 1 -f -o gpurun_out/r02_tc python bench.py --steps 1 --warmup 1 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sf_tables -c 1 -f -o gpurun_out/r02_tables python bench.py --steps 1 --warmup 1 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_tables.log 2>&1
cat gpurun_out/r02_gpu_tests.log; tail -2 gpurun_out/r02_smoke.log; cat gpurun_out/r02_npt.log; tail -c 300 gpurun_out/r02_bench.err
