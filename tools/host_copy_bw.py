"""Host staging copy (mb200_host_copy) bandwidth into pinned memory vs thread count, 12 MB frames (1M atoms)."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from maicos_b200 import _cabi
n = 1_000_000
src = np.random.default_rng(0).random((8, n, 3), dtype=np.float32)
dst = _cabi.pinned_empty((8, n, 3), np.float32, slot="bw")
for nt in (1, 2, 4, 6, 8, 12, 16):
    for f in range(8):
        _cabi.host_copy(dst[f], src[f], nt)
    t0 = time.perf_counter()
    for rep in range(5):
        for f in range(8):
            _cabi.host_copy(dst[f], src[f], nt)
    dt = (time.perf_counter() - t0) / 40
    print(f"threads {nt:2d}: {dt * 1e3:.3f} ms per 12 MB frame = {12e6 / dt / 1e9:.1f} GB/s")
t0 = time.perf_counter()
for rep in range(5):
    for f in range(8):
        np.copyto(dst[f], src[f])
print(f"np.copyto: {(time.perf_counter() - t0) / 40 * 1e3:.3f} ms")
