tag=${1:-all}
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value %.4e e2e %.4e ms/step %.3f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches']))
print('roofline', {k:d['roofline'][k] for k in ('kernel','achieved','peak','frac','kernel_ms_per_launch','kernel_share_of_step','fp64_equivalent_tflops')})
print('other', d['other_precision_mode'])
print('cpu', d.get('cpu_baseline')); print('parity', d.get('parity')); print('clocks', d.get('clocks'))
PY
