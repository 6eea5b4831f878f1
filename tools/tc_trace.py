"""Developer tool: per-K-step timeline of k_sf_contract_tc (CTA 0, first item) from a TRACE=1 build.

    make -C maicos_b200/csrc TRACE=1 && MAICOS_B200_LIB=maicos_b200/_lib/libmaicos_b200_trace.so python tools/tc_trace.py
"""
import ctypes, sys
sys.path.insert(0, ".")
import numpy as np
import maicos_b200 as mb
from maicos_b200 import _cabi

n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 33333
mb.set_precision("i8x5")
u = mb.synthetic.spce_universe(n_mol, 20261019, n_frames=2)
mb.Saxs(u.atoms, qmax=2.0, dq=0.1).run()
lib = _cabi.load()
buf = np.zeros(512 * 32, dtype=np.uint64)
lib.mb200_debug_tc_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.mb200_debug_tc_trace(buf.ctypes.data, buf.size) == 0
tr = buf.reshape(512, 32).astype(np.int64)
t0 = tr[0, 24]
T = tr[8:200]
gen_seed, gen_free, gen_done = T[:, 0:24:3], T[:, 1:24:3], T[:, 2:24:3]
a_rdy, b_rdy, issued, loaded = T[:, 24], T[:, 25], T[:, 26], T[:, 27]
print("period (a_full seen, cycles): mean %.0f" % np.diff(a_rdy).mean())
print("generator: seed wait -> free wait  mean %.0f ; free wait -> done (generation) mean %.0f ; done -> next seed ready %.0f" % (
    (gen_free - gen_seed).mean(), (gen_done - gen_free).mean(), (gen_seed[1:] - gen_done[:-1]).mean()))
print("spread of gen_done across warps per step: mean %.0f" % (gen_done.max(1) - gen_done.min(1)).mean())
print("last gen_done -> control sees a_full: mean %.0f" % (a_rdy - gen_done.max(1)).mean())
print("a_full+b_full seen -> issued+commit %.0f" % (issued - a_rdy).mean())
if loaded.any(): print("commit -> MMAs of the step complete (serialised build): mean %.0f min %.0f" % ((loaded - issued).mean(), (loaded - issued).min()))
# s_free(t) observed by generators of step t+4: earliest gen_free[t+4] is an upper bound on MMA(t) completion
lat = gen_free[4:].min(1) - issued[:-4]
print("commit(t) -> first generator past s_free wait of step t+4: mean %.0f min %.0f" % (lat.mean(), lat.min()))
for t in range(20, 32):
    print(t, "gen_seed", (gen_seed[t] - a_rdy[20]).tolist(), "\n   gen_free", (gen_free[t] - a_rdy[20]).tolist(), "\n   gen_done", (gen_done[t] - a_rdy[20]).tolist(),
          "\n   ctl ready,issued", int(a_rdy[t] - a_rdy[20]), int(issued[t] - a_rdy[20]))
print("MMA warp per K-step (relative to step 20): operands seen, commit issued | first generator past the s_free wait of step t + 4")
for t in range(20, 32):
    print(t, int(T[t, 24] - a_rdy[20]), int(T[t, 26] - a_rdy[20]), "|", int(gen_free[t + 4].min() - a_rdy[20]))
