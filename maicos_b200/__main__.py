"""Command line front end: ``python -m maicos_b200 <module> -s <topology> [-f <trajectory>] ...``.

The reference builds its CLI with mdacli from the docstrings (``/root/reference/src/maicos/__main__.py:16-25``;
usage ``examples/usage-bash.sh``: ``maicos densityplanar -s slit_flow.tpr -f slit_flow.trr -atomgroup 'type OW HW'
-concfreq 10``).  When mdacli and MDAnalysis are importable the same generator is used on this
package's classes; otherwise a small argparse front end with the same flag spelling (one
``-<parameter>`` flag per constructor argument, ``-s / -f / -b / -e / -dt / -v`` for the universe and the
frame range) loads GROMACS ``.gro`` / ``.trr`` files through the in-memory shim.
"""

from __future__ import annotations

import argparse
import inspect
import sys

from . import __version__, modules
from .core import AnalysisBase


def _module_classes() -> dict:
    return {name.lower(): getattr(modules, name) for name in modules.__all__}


def _convert(default, text: str):
    if isinstance(default, bool):
        return text.lower() in ("1", "true", "yes", "on")
    if isinstance(default, (int, float)):
        try:
            return int(text)
        except ValueError:
            return float(text)
    if default is None:
        if text.lower() == "none":
            return None
        try:
            return float(text)
        except ValueError:
            return text
    return text


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="maicos_b200", description="MAICoS hot-path analyses on B200 (CUDA, no CPU fallback).")
    ap.add_argument("--version", action="version", version=f"maicos_b200 {__version__}")
    sub = ap.add_subparsers(dest="module", required=True)
    for name, cls in _module_classes().items():
        sp = sub.add_parser(name, help=(cls.__doc__ or "").strip().splitlines()[0])
        sp.add_argument("-s", dest="topology", required=True, help="topology / coordinate file (.gro)")
        sp.add_argument("-f", dest="trajectory", default=None, help="trajectory file (.trr)")
        sp.add_argument("-b", dest="begin", type=int, default=None, help="first frame (index)")
        sp.add_argument("-e", dest="end", type=int, default=None, help="last frame (index, exclusive)")
        sp.add_argument("-dt", dest="step", type=int, default=None, help="frame step")
        sp.add_argument("-v", dest="verbose", action="store_true")
        for pname, par in inspect.signature(cls.__init__).parameters.items():
            if pname == "self":
                continue
            if pname in ("atomgroup", "refgroup"):
                sp.add_argument(f"-{pname}", default="all" if pname == "atomgroup" else None,
                                help="selection string, e.g. 'type OW HW'")
            else:
                sp.add_argument(f"-{pname}", default=None, help=f"default: {par.default!r}")
    return ap


def main(argv=None) -> int:
    try:  # the reference's own generator when its dependencies exist
        import MDAnalysis  # noqa: F401
        from mdacli import cli

        cli(name="MAICoS-B200", module_list=["maicos_b200"], base_class=AnalysisBase, version=__version__,
            description=__doc__, ignore_warnings=True)
        return 0
    except ImportError:
        pass
    args = build_parser().parse_args(argv)
    cls = _module_classes()[args.module]
    from .mdshim import universe_from_gro

    u = universe_from_gro(args.topology, args.trajectory)
    kwargs = {}
    for pname, par in inspect.signature(cls.__init__).parameters.items():
        if pname == "self":
            continue
        value = getattr(args, pname)
        if pname == "atomgroup":
            kwargs[pname] = u.select_atoms(value)
        elif pname == "refgroup":
            if value is not None:
                kwargs[pname] = u.select_atoms(value)
        elif value is not None:
            kwargs[pname] = _convert(par.default, value)
    analysis = cls(**kwargs).run(start=args.begin, stop=args.end, step=args.step, verbose=args.verbose)
    if analysis.module_has_save:
        analysis.save()
    return 0


if __name__ == "__main__":
    sys.exit(main())
