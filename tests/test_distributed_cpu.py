"""N > 1 path on CPU: world_size 2 over gloo -- frame sharding + the single merging allreduce."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from maicos_b200 import parallel
from test_host_logic import Series, small_universe

HERE = Path(__file__).resolve().parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_merge_welford_single_process_is_identity():
    means = {"a": np.array([1.0, 2.0]), "b": 3.0}
    sems = {"a": np.array([0.1, 0.2]), "b": 0.3}
    sums = {"a": np.array([5.0, 10.0]), "b": 15.0}
    n, m, s, t, ex = parallel.merge_welford(5, means, sems, sums, {"ts": np.arange(5.0)})
    assert n == 5
    assert_allclose(m["a"], means["a"]) and None
    assert_allclose(s["a"], sems["a"])
    assert_allclose(s["b"], 0.3)
    assert_allclose(t["a"], sums["a"])
    assert_array_equal(ex["ts"], np.arange(5.0))


@pytest.mark.parametrize("world", [2, 3])
def test_ranks_frame_sharding_matches_serial(tmp_path, world):
    out = tmp_path / "dist.npz"
    env = dict(os.environ, OMP_NUM_THREADS="1")
    env.pop("WORLD_SIZE", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), str(HERE / "_dist_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    d = np.load(out)
    assert int(d["world"]) == world
    # two frames on `world` ranks: with three ranks one has no frame and joins the merge with n = 0
    few = Series(small_universe(n_frames=23).atoms).run(start=3, stop=5)
    assert int(d["few_index"]) == few.n_frames == 2
    assert_allclose(d["few_mean"], few.means.observable, rtol=1e-13)
    assert_allclose(d["few_sem"], few.sems.observable, rtol=1e-10)
    assert_allclose(d["few_vec"], few.means.vec, rtol=1e-13)
    assert int(d["few_count"]) == few.sums.count
    serial = Series(small_universe(n_frames=23).atoms).run(start=1, step=2)
    assert int(d["index"]) == serial.n_frames == 11
    assert_array_equal(d["frames"], serial.frames)
    assert_allclose(d["times"], serial.times)
    assert_allclose(d["timeseries"], serial.timeseries, rtol=0, atol=0)
    assert_allclose(d["mean"], serial.means.observable, rtol=1e-13)
    assert_allclose(d["sem"], serial.sems.observable, rtol=1e-10)
    assert_allclose(d["total"], serial.sums.observable, rtol=1e-13)
    assert_allclose(d["vec_mean"], serial.means.vec, rtol=1e-13)
    assert_allclose(d["vec_sem"], serial.sems.vec, rtol=1e-10)
    assert int(d["count"]) == serial.sums.count  # integer observables stay exact
    assert_allclose(d["result"], serial.results.mean, rtol=1e-13)
    assert_allclose(d["corrtime"], serial.corrtime)
    assert_array_equal(d["buf"], np.full(7, world * (world + 1) / 2.0))
    # concfreq under frame sharding: partial results are merged and concluded every (concfreq // world) local frames,
    # i.e. at the global frame counts a serial run with concfreq = (concfreq // world) * world concludes at
    from _dist_worker import Checkpointed

    every = max(1, 4 // world) * world
    ser = Checkpointed(small_universe(n_frames=23).atoms, concfreq=every).run()
    assert [c[0] for c in ser.log] == list(d["cp_index"]) and len(ser.log) == 23 // every + 1
    assert_allclose(d["cp_mean"], [c[1] for c in ser.log], rtol=1e-13)
    assert_allclose(d["cp_sem"], [c[2] for c in ser.log], rtol=1e-9)
    assert_allclose(d["cp_vec"], np.stack([c[3] for c in ser.log]), rtol=1e-13)
