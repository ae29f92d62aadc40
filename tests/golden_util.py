"""Shared helpers for the golden-fixture tests (test infrastructure)."""
import glob
import os

import numpy as np

import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixtures():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLD, "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    g = {k: z[k] for k in z.files}
    cfg = O.Config.from_name(str(g["config"]))
    return cfg, str(g["script"]), g


def tags_in_order(script):
    return [ln.split()[1] for ln in script.strip().splitlines() if ln.split() and ln.split()[0] == "D"]


def first_dump_after_ic(script):
    seen = False
    for ln in script.strip().splitlines():
        t = ln.split()
        if not t:
            continue
        if t[0] == "I":
            seen = True
        if t[0] == "D" and seen:
            return t[1]
    return None


def ic_from_golden(cfg, script, g):
    """solver.initialize() evaluates exp() with the reference's libm; take the interior of the first
    dump after the I op from the fixture instead so that comparisons can be bit-exact."""
    tag = first_dump_after_ic(script)

    def ic(tree):
        P = tree.size
        d = g[tag + "/data"].reshape((cfg.nvar, P) + (cfg.psize,) * cfg.rank)
        return d[(slice(None), slice(None)) + O.interior_slices(cfg)]

    return ic


def rel_err(a, b):
    """field-max-normalised error: max|a-b| / max|b| per field, worst field."""
    worst = 0.0
    for f in range(a.shape[0]):
        den = max(np.abs(b[f]).max(), 1e-300)
        worst = max(worst, np.abs(a[f] - b[f]).max() / den)
    return worst
