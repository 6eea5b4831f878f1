"""Generate the XTC fixture of tests/golden/ (run in the build container only; /root/reference is not on the GPU box).

``xtc_cntwater_2frames.xtc``: the first two frames of the reference's ``tests/data/cntwater/cntwater.xtc``, byte for byte.
``xtc_cntwater_frame0.npz``: the coordinates of the same system's ``cntwater.gro`` (written by another program from the same
first frame, three decimals in nm) -- an independent statement of what frame 0 decodes to -- plus box and times.

    python tests/golden/make_golden_xtc.py
"""
import struct
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from maicos_b200.mdshim import read_gro  # noqa: E402  (file parsing only)

OUT = Path(__file__).resolve().parent
REF = Path("/root/reference/tests/data/cntwater")

data = (REF / "cntwater.xtc").read_bytes()
off, cuts = 0, []
for _ in range(2):  # frame header: magic, natoms, step, time, box[9], natoms, precision, minint[3], maxint[3], smallidx, nbytes
    magic, natoms = struct.unpack_from(">ii", data, off)
    assert magic == 1995 and natoms > 9
    (nbytes,) = struct.unpack_from(">i", data, off + 56 + 32)
    off += 56 + 36 + (nbytes + 3) // 4 * 4
    cuts.append(off)
(OUT / "xtc_cntwater_2frames.xtc").write_bytes(data[:cuts[-1]])
gro = read_gro(REF / "cntwater.gro")
np.savez_compressed(OUT / "xtc_cntwater_frame0.npz", gro_positions=gro["positions"], dimensions=gro["dimensions"],
                    times=np.array([0.0, 0.5]), n_frames_in_file=51)
print("wrote", cuts[-1], "bytes of XTC and", gro["positions"].shape, "reference coordinates")

# a second, geometric check of the decoder on a file that is too large to commit: every O-H bond of every frame of the
# reference's examples/water_nvt.xtc (512 SPC/E molecules, 201 frames) has its model length
from maicos_b200.mdshim import read_xtc  # noqa: E402

w = read_xtc("/root/reference/examples/water_nvt.xtc")
L = w["dimensions"][:, None, :3]
for h in (1, 2):
    v = w["positions"][:, h::3] - w["positions"][:, 0::3]
    v -= np.round(v / L) * L
    d = np.linalg.norm(v, axis=2)
    assert 0.98 < d.min() and d.max() < 1.02, (d.min(), d.max())
print("water_nvt.xtc:", w["positions"].shape, "all O-H bonds within", float(d.min()), float(d.max()))
