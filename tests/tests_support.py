"""Shared helpers of the test suite."""
import glob
import os

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def golden_files(kind: str, extra_dirs=()):
    """Golden fixtures of one problem kind ("kino", "toy"): the committed oracle-generated files AND, when present,
    the CasADi dumps `casadi_<kind>_*.npz` written by tools/dump_casadi_golden.py in the reference environment --
    both are compared with the CUDA path / the oracle by the same tests."""
    out = []
    for d in (GOLD, *extra_dirs):
        out += sorted(glob.glob(os.path.join(d, f"{kind}_*.npz"))) + sorted(glob.glob(os.path.join(d, f"casadi_{kind}_*.npz")))
    return out
