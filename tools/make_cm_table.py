"""Regenerate ``maicos_b200/share/cromer_mann.json`` from the reference's data file.

The Cromer-Mann coefficients are published data (D. T. Cromer & J. B. Mann, Acta Cryst. A 24
(1968) 321; International Tables for Crystallography Vol. C ch. 6.1).  MAICoS ships them as
``src/maicos/share/scatteringfactors.dat`` (read by ``lib/tables.py:31-43``).  This script
parses that file where it lies and writes the values as JSON so that the table travels to
machines without ``/root/reference``; run it only when the reference's table changes.
"""
import json
import sys
from pathlib import Path

src = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/maicos/share/scatteringfactors.dat")
dst = Path(__file__).resolve().parents[1] / "maicos_b200" / "share" / "cromer_mann.json"
table = {}
for line in src.read_text().splitlines():
    if not line or line[0] == "#":
        continue
    p = line.split()
    table[p[0]] = {"a": [float(x) for x in p[1:5]], "b": [float(x) for x in p[5:9]], "c": float(p[9])}
dst.write_text(json.dumps({"source": "Cromer & Mann, Acta Cryst. A24 (1968) 321", "elements": table}, indent=0))
print(f"{len(table)} elements -> {dst}")
