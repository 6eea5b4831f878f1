set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python tools/r2_c4_probe.py 2>&1 | grep -v "^pageable" | tail -8
