set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --configs C4 > gpurun_out/r02_c4_collection.json 2> gpurun_out/r02_c4_collection.err
tail -3 gpurun_out/r02_c4_collection.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_c4_collection.json').read().strip().splitlines()[-1])
c = d.get('configs', {})
for k, v in c.items():
    if isinstance(v, dict):
        print(k, {a: b for a, b in v.items() if a in ('frames_per_s', 'steady_state_frames_per_s', 'speedup_vs_solo_runs', 'h2d_copies_over_the_two_timed_passes', 'three_solo_runs_seconds', 'e2e_seconds', 'error')})
PY
