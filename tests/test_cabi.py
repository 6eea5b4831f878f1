"""The C-ABI library: loads, exports every symbol the header declares, fails loudly without a GPU."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from maicos_b200 import _cabi

ROOT = Path(__file__).resolve().parents[1]


def header_symbols():
    text = (ROOT / "include" / "maicos_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exists_and_loads():
    assert _cabi.library_path().exists(), "build it: python -c 'import __graft_entry__ as g; g.build()'"
    lib = _cabi.load()
    assert lib.mb200_abi_version() >= 1


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(str(_cabi.library_path()))
    names = header_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/maicos_b200.h but not exported"
        assert name in _cabi.SIGNATURES, f"{name} has no ctypes signature in maicos_b200/_cabi.py"
    assert sorted(_cabi.SIGNATURES) == names


def test_maxn_host_helper():
    """mb200_sfactor_maxn is pure host code: _cmath.pyx:93-96 including the float32 cast."""
    from maicos_b200.lib.math import sfactor_maxn
    from oracle import kernel

    rng = np.random.default_rng(5)
    for _ in range(500):
        dims = rng.uniform(5, 400, 3)
        qmax = rng.uniform(0.05, 6)
        assert tuple(sfactor_maxn(dims, qmax)) == tuple(kernel.port_maxn(dims, qmax))


def test_no_cpu_fallback_without_device():
    """No GPU -> every compute entry point raises; there is nothing to fall back to."""
    if _cabi.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    from maicos_b200.lib.math import structure_factor
    from maicos_b200.lib._hist import profile_hist

    with pytest.raises(_cabi.MB200Error, match="no CUDA device"):
        structure_factor(np.zeros((4, 3)), np.array([10.0, 10, 10]), 0, 2, 0, np.pi, np.ones(4))
    with pytest.raises(_cabi.MB200Error, match="no CUDA device"):
        profile_hist(0, np.zeros((1, 4, 3), np.float32), 2, None, np.linspace(0, 1, 5)[None])
    # NULL context is rejected by the library itself as well
    rc = _cabi.load().mb200_ctx_synchronize(None)
    assert rc == -2 and b"no CUDA context" in _cabi.load().mb200_last_error()


def test_reference_error_conventions():
    """Probed on the compiled reference: dtype mismatch -> ValueError, bad shapes -> AssertionError."""
    from maicos_b200.lib.math import compute_structure_factor, structure_factor

    assert compute_structure_factor is structure_factor  # CHANGELOG.rst:38-40
    with pytest.raises(ValueError, match="Buffer dtype mismatch, expected 'double' but got 'float'"):
        structure_factor(np.zeros((4, 3), np.float32), np.array([10.0, 10, 10]), 0, 2, 0, np.pi, np.ones(4))
    with pytest.raises(AssertionError):
        structure_factor(np.zeros((4, 2)), np.array([10.0, 10, 10]), 0, 2, 0, np.pi, np.ones(4))
    with pytest.raises(AssertionError):
        structure_factor(np.zeros((4, 3)), np.array([10.0, 10, 10]), 0, 2, 0, np.pi, np.ones(5))


def test_library_sass_carries_the_claimed_instructions():
    """The built library targets sm_100a and its hot kernels contain the instructions DESIGN.md names: tcgen05 int8 MMAs
    (UTCIMMA) with one commit (UTCBAR), TMEM stores / loads (STTM / LDTM) and bulk copies (UBLKCP) in the split-precision
    contraction, FP64 DMMA in the FP64 contraction, native shared-memory atomics in the histogram kernels."""
    import shutil
    import subprocess
    from pathlib import Path

    import maicos_b200

    lib = Path(maicos_b200.__file__).parent / "_lib" / "libmaicos_b200.so"
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not lib.exists() or not Path(cuobjdump).exists():
        pytest.skip("library or cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", str(lib)], capture_output=True, text=True, timeout=600).stdout
    assert "sm_100a" in out or "EF_CUDA_SM100" in out
    kernels = {}
    name = None
    for line in out.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            kernels[name] = []
        elif name is not None:
            kernels[name].append(line)

    def body(fragment):
        hits = [k for k in kernels if fragment in k]
        assert hits, f"no kernel named *{fragment}* in the library"
        return "\n".join("\n".join(kernels[k]) for k in hits)

    tc = body("k_sf_contract_tc")
    for op in ("UTCIMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS"):
        assert op in tc, f"{op} missing from k_sf_contract_tc"
    assert "STTM" in body("k_sf_contract_tcILb1"), "the TMEM-operand variant must store its A operand with tcgen05.st"
    assert "DMMA" in body("13k_sf_contractE"), "FP64 contraction without DMMA"
    assert "UBLKCP" in body("13k_sf_contractE"), "FP64 contraction without bulk-copy staging"
    assert "ATOMS" in body("k_hist_sorted") and "ATOMS" in body("6k_histILi0")
    assert "UBLKPF" in body("k_hist_sorted"), "the counting-sort histogram prefetches its next tile into L2 with a bulk prefetch"


@pytest.mark.parametrize("L,qmax,n_bins,window", [
    (24.87, 6.0, 60, (0.0, 0.0, np.pi)),          # shipped water, default q range
    (99.97, 2.0, 20, (0.0, 0.0, np.pi)),          # BASELINE config 2
    (215.37, 2.0, 20, (0.0, 0.0, np.pi)),         # 1M atoms, two l tiles
    (60.0, 3.0, 30, (0.5, 0.3, 1.2)),             # qmin and theta window
    (50.0, 3.0, 0, (0.0, 0.0, np.pi)),            # raw-cube entry order
])
def test_q_plan_host_side(L, qmax, n_bins, window, monkeypatch):
    """``mb200_plan_describe`` (host only): the plan accepts exactly the q vectors the reference kernel accepts
    (_cmath.pyx:93-115 through the oracle's port), and the multi-threaded construction equals the serial one."""
    import ctypes

    from oracle import kernel

    lib = _cabi.load()
    dims = np.double(np.float32([L, L * 1.01, L * 0.99]))
    qmin, tmin, tmax = window

    def describe():
        out = np.zeros(16, dtype=np.int64)
        _cabi.check(lib.mb200_plan_describe(_cabi.ptr(dims), qmin, qmax, tmin, tmax, n_bins, _cabi.ptr(out)))
        return out.tolist()

    monkeypatch.setenv("MAICOS_B200_HOST_THREADS", "1")
    serial = describe()
    monkeypatch.setenv("MAICOS_B200_HOST_THREADS", "5")
    threaded = describe()
    assert serial == threaded
    q, _ = kernel.port_structure_factor(np.zeros((1, 3)), dims, qmin, qmax, tmin, tmax, np.ones(1))
    assert serial[0] == np.count_nonzero(q)
    assert serial[1:4] == [int(m) for m in kernel.port_maxn(dims, qmax)]
    n_rt, n_lt, lt, used = serial[5:9]
    assert serial[4] == 1 and n_lt * lt >= serial[3] and lt % 8 == 0 and lt <= 48
    assert 0 < used <= n_rt * n_lt
    if n_lt > 1:  # quads sorted by their highest populated l tile: blocks without any valid q are never launched
        assert used < n_rt * n_lt
