"""Worker of tests/test_distributed_cpu.py: frame-sharded toy analysis under torchrun + gloo."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))

from maicos_b200 import parallel  # noqa: E402
from test_host_logic import Series, small_universe  # noqa: E402


class Checkpointed(Series):
    """Records every ``_conclude`` (concfreq checkpoints + the final one)."""

    def _conclude(self):
        super()._conclude()
        self.__dict__.setdefault("log", []).append((int(self._index), float(self.means.observable),
                                                    float(self.sems.observable), np.array(self.means.vec)))


def main(out):
    parallel.init(mode="frames", backend="gloo")
    u = small_universe(n_frames=23)
    cp = Checkpointed(u.atoms, concfreq=4).run()
    ana = Series(u.atoms).run(start=1, step=2)
    few = Series(u.atoms).run(start=3, stop=5)  # two frames: with three ranks one of them has nothing to analyse
    # q-block style reduction of a per-frame partial sum
    buf = np.full(7, float(parallel.rank() + 1))
    parallel.allreduce_sum(buf)
    if parallel.rank() == 0:
        np.savez(out, mean=ana.means.observable, sem=ana.sems.observable, total=ana.sums.observable,
                 vec_mean=ana.means.vec, vec_sem=ana.sems.vec, count=ana.sums.count, frames=ana.frames,
                 times=ana.times, timeseries=ana.timeseries, index=ana._index, result=ana.results.mean,
                 corrtime=ana.corrtime, buf=buf, world=parallel.world(), few_mean=few.means.observable,
                 few_sem=few.sems.observable, few_count=few.sums.count, few_index=few._index, few_vec=few.means.vec,
                 cp_index=[c[0] for c in cp.log], cp_mean=[c[1] for c in cp.log], cp_sem=[c[2] for c in cp.log],
                 cp_vec=np.stack([c[3] for c in cp.log]))
    parallel.barrier()
    parallel.shutdown()


if __name__ == "__main__":
    main(sys.argv[1])
