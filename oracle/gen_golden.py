"""TEST INFRASTRUCTURE: generate tests/golden/*.npz by running the UNMODIFIED reference
(oracle/_ref/ref_dump_<cfg>, built from /root/reference by oracle/Makefile) over fixed scripts.

Run in the build container only (needs /root/reference):  python oracle/gen_golden.py
The fixtures are committed; the GPU box and the test-suite never need the reference.

Script language: see oracle/ref_dump.cpp.  Ids used by R/K ops: id = morton << 6 | level.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from refdump_io import load  # noqa: E402
from oracle import Config, face_halo_mask  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

# (fixture name, config, script, {tag: "full" | "interior-stats"})
CASES = [
    # single periodic root: a patch that is its own neighbor in every direction (N4)
    ("root2d", "r2_s8_h1_d7_euler", """
P
X
D probe
I
X
D t0
S 3
D t3
""", {}),
    # level 1: +/- neighbors are the same patch; then one child refined: finer across the wrap
    ("wrap2d", "r2_s8_h1_d7_adv", """
A
X
P
X
D l1
R 1
X
P
X
D l1r
I
X
D t0
S 4
D t4
""", {}),
    # multi-level 2D Euler with hash refine + coarsen (balancing ripple + veto exercised)
    ("amr2d_euler", "r2_s8_h1_d7_euler", """
A
X
A
X
H 1 300 0 1 4
X
H 2 300 300 1 5
X
H 3 250 350 1 6
X
P
X
D probe
I
X
D t0
S 5
D t5
H 4 200 400 1 6
X
D regrid
S 3
D t8
""", {}),
    # the C1 configuration: 10x10 patches, halo width 2, advection
    ("c1_adv_h2", "r2_s10_h2_d7_adv", """
A
X
A
X
H 11 350 0 1 5
X
H 12 300 300 1 5
X
P
X
D probe
I
X
D t0
S 10
D t10
H 13 300 300 1 5
X
D regrid
S 5
D t15
""", {}),
    # BASELINE.md known-answer config KA-2D (16x16 Euler, levels 3-5, 280 patches)
    ("ka2d", "r2_s16_h1_d7_euler", """
A
X
A
X
A
X
B 0.25 99 7 0.5 0.5
X
B 0.25 99 7 0.5 0.5
X
I
X
D t0
S 10
D t10
S 190
D t200
""", {"t200": "stats"}),
    # 3D Euler, 4^3 patches, three levels
    ("amr3d_euler", "r3_s4_h1_d5_euler", """
A
X
H 21 400 0 1 3
X
H 22 150 300 1 4
X
P
X
D probe
I
X
D t0
S 4
D t4
H 23 150 400 1 4
X
D regrid
S 2
D t6
""", {}),
    # 3D Euler at the reference's literal 8^3 patch shape (C3)
    ("c3_euler", "r3_s8_h1_d5_euler", """
A
X
R 1 1835009
X
P
X
D probe
I
X
D t0
S 3
D t3
""", {}),
    # 8^3 patches on three levels with a regrid (refine + coarsen) between two runs of steps:
    # coarse/fine faces in all directions for the 3D plane-marching kernel and the 3D plan kernel
    ("c3_amr", "r3_s8_h1_d5_euler", """
A
X
H 41 400 0 1 2
X
I
X
D t0
S 4
D t4
H 42 300 300 1 3
X
D regrid
S 3
D t7
""", {"t4": "stats", "regrid": "stats", "t7": "stats"}),
    # 3D advection, halo 2
    ("adv3d_h2", "r3_s4_h2_d5_adv", """
A
X
H 31 400 0 1 3
X
H 32 200 300 1 4
X
P
X
D probe
I
X
D t0
S 6
D t6
""", {}),
    ("adv3d", "r3_s8_h1_d5_adv", """
A
X
R 1 1835009
X
I
X
D t0
S 5
D t5
""", {}),
]


def main():
    os.makedirs(GOLD, exist_ok=True)
    total = 0
    only = set(sys.argv[1:])                 # optional: regenerate only the named fixtures
    for name, cfg, script, opts in CASES:
        if only and name not in only:
            continue
        exe = os.path.join(HERE, "_ref", "ref_dump_" + cfg)
        with tempfile.TemporaryDirectory() as td:
            sp, op = os.path.join(td, "s.txt"), os.path.join(td, "o.bin")
            open(sp, "w").write(script)
            subprocess.check_call([exe, sp, op])
            d = load(op)
        out = {"script": np.array(script), "config": np.array(cfg)}
        for k, v in d.items():
            tag = k.split("/")[0]
            if opts.get(tag) == "stats" and k.endswith("/data"):
                # too large to commit in full: keep per-field sums / max over interior+face cells
                m = face_halo_mask(Config.from_name(cfg)).ravel()
                out[tag + "/sum"] = v[..., m].sum(axis=(1, 2))
                out[tag + "/max"] = v[..., m].max(axis=(1, 2))
                out[tag + "/first_patch"] = v[:, 0, :]
                continue
            out[k] = v
        path = os.path.join(GOLD, name + ".npz")
        np.savez_compressed(path, **out)
        sz = os.path.getsize(path)
        total += sz
        tags = sorted({k.split("/")[0] for k in d if "/" in k})
        print("%-14s %-22s %7.1f KB  patches=%s" % (
            name, cfg, sz / 1024.0, [int(d[t + "/ids"].shape[0]) for t in tags]))
    print("total %.1f KB" % (total / 1024.0))


if __name__ == "__main__":
    main()
