# one full ncu capture of the tc contraction kernel (after the plain run above exited 0)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sf_contract_tc -s 2 -c 1 -o gpurun_out/${1:-tc} -f \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --precision i8x5 > gpurun_out/${1:-tc}_ncu.log 2>&1
tail -2 gpurun_out/${1:-tc}_ncu.log
