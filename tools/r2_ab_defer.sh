for i in 1 2; do
for d in 1 0; do
MAICOS_B200_DEFER=$d python bench.py --no-cpu-baseline --no-configs > gpurun_out/defer_$d.json 2> /dev/null
python -c "
import json
d=json.loads(open('gpurun_out/defer_$d.json').read().strip().splitlines()[-1])
print('defer=$d value %.4e (%.3f ms)  e2e %.4e (%.3f ms)  pageable %.4e' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['pageable_trajectory']['value']), d['clocks']['sm_mhz'])
"
done
done
