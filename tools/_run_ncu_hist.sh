# one full ncu capture of the weighted-histogram kernel at the config-4 size (after a plain run exited 0)
tag=${1:-hs}
kern=${2:-k_hist_sorted}
timeout 200 python tools/bench_profile.py > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${kern} -s 3 -c 1 -o gpurun_out/${tag} -f \
  python tools/bench_profile.py > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
