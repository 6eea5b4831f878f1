set -x
python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_sfactor.py -x -q -m gpu -k "unbinned or cells" 2>&1 | tail -15
python tools/r2_unbinned.py 33333 64 > gpurun_out/r02_unbinned.json 2> gpurun_out/r02_unbinned.err; cat gpurun_out/r02_unbinned.json; tail -5 gpurun_out/r02_unbinned.err
