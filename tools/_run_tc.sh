tag=${1:-x}
timeout 300 python -m pytest tests/test_gpu_sfactor.py -x -q -m gpu -k "split_precision or tensor" 2>&1 | tail -3
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --precision i8x5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['kernel_ms_per_launch'], d['ms_per_step'], d.get('other_precision_mode'))"
MAICOS_B200_LIB=maicos_b200/_lib/libmaicos_b200_trace.so timeout 200 python tools/tc_trace.py 2>&1 | grep -v Warn | head -${2:-12}
