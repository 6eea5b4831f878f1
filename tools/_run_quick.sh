tag=${1:-q}
timeout 600 python -m pytest tests/test_gpu_sfactor.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1]); print('value %.4e e2e %.4e' % (d['value'], d['e2e']['value']), d['roofline']['kernel_ms_per_launch'], d['ms_per_step'], d['other_precision_mode']['max_rel_diff_binned_S_vs_headline_mode'])"
