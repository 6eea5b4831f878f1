# usage: _run_variants.sh tag suffix1 suffix2 ...   (bench in i8x5 with each library variant)
tag=$1; shift
for v in "$@"; do
  MAICOS_B200_LIB=maicos_b200/_lib/libmaicos_b200$v.so timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --precision i8x5 > gpurun_out/${tag}$v.json 2> gpurun_out/${tag}$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/${tag}$v.json').read().strip().splitlines()[-1]); print('$v', '%.4e' % d['value'], d['roofline']['kernel_ms_per_launch'], d['ms_per_step'], d['other_precision_mode']['max_rel_diff_binned_S_vs_headline_mode'])"
done
