"""Config 5 under torchrun: Saxs on ONE 1M-atom frame (qmax = 3/A), q-vector blocks sharded over the ranks
(parallel.init(mode="qblocks")), against the single-GPU result of rank 0.  Prints wall / device time per rank count."""
import sys, time, warnings
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
warnings.simplefilter("ignore")
import maicos_b200 as mb
from maicos_b200 import parallel

u = mb.synthetic.spce_universe(333_333, 20261022, n_frames=1)
mb.Saxs(u.atoms, qmax=3.0).run()                      # warm-up, every rank alone on its GPU
t0 = time.perf_counter(); ref = mb.Saxs(u.atoms, qmax=3.0).run(); t_one = time.perf_counter() - t0
dev_one = ref.timings["device_call_s"]
parallel.init(mode="qblocks")
mb.Saxs(u.atoms, qmax=3.0).run()                      # warm-up of the sharded plan blocks + NCCL
parallel.barrier()
t0 = time.perf_counter(); got = mb.Saxs(u.atoms, qmax=3.0).run(); parallel.barrier(); t_sh = time.perf_counter() - t0
if parallel.rank() == 0:
    a, b = ref.results.scattering_intensities, got.results.scattering_intensities
    err = float(np.nanmax(np.abs(a - b) / np.abs(a)))
    print(f"C5 one frame, 999 999 atoms, qmax 3: 1 GPU wall {t_one:.4f} s (device call {dev_one:.4f} s); "
          f"{parallel.world()} GPUs q-blocks wall {t_sh:.4f} s (device call {got.timings['device_call_s']:.4f} s); "
          f"speed-up {t_one / t_sh:.2f}x; max rel diff I(q) {err:.2e}; counts equal {bool(np.array_equal(ref.sums.bincount, got.sums.bincount))}")
parallel.shutdown()
