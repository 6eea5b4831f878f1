import ctypes, sys
sys.path.insert(0, ".")
import numpy as np
import maicos_b200 as mb
from maicos_b200 import _cabi
mb.set_precision("i8x5")
u = mb.synthetic.spce_universe(33333, 20261019, n_frames=2)
mb.Saxs(u.atoms, qmax=2.0, dq=0.1).run()
lib = _cabi.load()
buf = np.zeros(512 * 32, dtype=np.uint64)
lib.mb200_debug_tc_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.mb200_debug_tc_trace(buf.ctypes.data, buf.size) == 0
tr = buf.reshape(512, 32).astype(np.int64)
T = tr[20:120]
loop_top, after_b, after_a, after_fence, issued = T[:,28], T[:,25], T[:,24], T[:,27], T[:,26]
print("per K-step means: period %.0f | top->after b_full wait %.0f | b->after a_full wait %.0f | fence %.0f | fence->issued+commit+syncwarp %.0f | issued->next loop top %.0f" % (
    np.diff(after_a).mean(), (after_b-loop_top).mean(), (after_a-after_b).mean(), (after_fence-after_a).mean(), (issued-after_fence).mean(), (loop_top[1:]-issued[:-1]).mean()))
for t in range(5): print((T[t,[28,25,24,27,26]]-T[0,28]).tolist())
