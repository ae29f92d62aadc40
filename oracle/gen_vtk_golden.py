"""TEST INFRASTRUCTURE: golden VTK files written by the UNMODIFIED reference's
include/ndtree/vtk_print.hpp (through oracle/_ref/ref_dump_<cfg>, op V) for small probe-filled
trees -> tests/golden/vtk/<name>.npz (file bytes + leaf ids + padded patch data).
tests/test_vtk_print.py feeds the same ids and data to include/ndtree/vtk_print.hpp of this repo
(host only) and requires identical bytes.   Run in the build container:  python oracle/gen_vtk_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from refdump_io import load  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "vtk")
CASES = [
    # 2D Euler, 8x8 patches: root -> 4 -> child 0 refined (7 leaves, levels 1-2), index-probe data
    ("vtk_r2_s8_h1_d7_euler", "r2_s8_h1_d7_euler", "A\nX\nR 1\nX\nP\nX\nV probe\nD probe\n"),
    # 3D Euler, 4^3 patches: root -> 8 -> child 0 refined (15 leaves)
    ("vtk_r3_s4_h1_d5_euler", "r3_s4_h1_d5_euler", "A\nX\nR 1\nX\nP\nX\nV probe\nD probe\n"),
    # 2D advection with halo 2 (the C1 patch shape)
    ("vtk_r2_s10_h2_d7_adv", "r2_s10_h2_d7_adv", "A\nX\nP\nX\nV probe\nD probe\n"),
]

os.makedirs(OUT, exist_ok=True)
for name, cfg, script in CASES:
    with tempfile.TemporaryDirectory() as td:
        sp, op = os.path.join(td, "s.txt"), os.path.join(td, "o.bin")
        open(sp, "w").write(script)
        subprocess.run([os.path.join(HERE, "_ref", "ref_dump_" + cfg), sp, op], check=True, cwd=td,
                       stdout=subprocess.DEVNULL)
        d = load(op)
        vtk = np.frombuffer(open(os.path.join(td, "vtk_output", "probe.vtk"), "rb").read(), dtype=np.uint8)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), config=cfg, vtk=vtk, ids=d["probe/ids"],
                            data=d["probe/data"])
        print(name, len(vtk), "bytes,", len(d["probe/ids"]), "leaves, data", d["probe/data"].shape)
