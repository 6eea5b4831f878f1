# final ncu evidence for profiles/: launch list of the bench command + one full capture per contraction kernel
set -x
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r3_plain.json 2> gpurun_out/r3_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r3_ll.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sf_contract_tc -s 2 -c 1 -o gpurun_out/r3_tc -f \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r3_tc_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sf_tables -s 2 -c 1 -o gpurun_out/r3_tables -f \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r3_tables_ncu.log 2>&1
ls -la gpurun_out/r3_*
