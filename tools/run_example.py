"""BASELINE config 1: the reference's example_registration pipeline (box filter, 0.25 m voxel grid,
KD-tree build, KNN k=10, covariances, normals, GICP/LM/Geman-McClure with robust-scale annealing and
the default 1000-point random sampling) through the C++ facade, on the bundled scan pair as committed
under tests/golden/ (the voxelised clouds: the raw .ply files live in the reference repository, which
the GPU box does not have).  Prints the example's own per-stage table.  usage: run_example.py [loops]"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_cpp_facade import compile_cpp, write_ply  # noqa: E402

loops = sys.argv[1] if len(sys.argv) > 1 else "100"
exe = compile_cpp("examples/example_registration.cpp", "example_registration")
d = np.load(os.path.join(ROOT, "tests", "golden", "bundled_pair.npz"))
with tempfile.TemporaryDirectory() as tmp:
    src, tgt, gt = (os.path.join(tmp, n) for n in ("source.ply", "target.ply", "T.txt"))
    write_ply(src, d["source_ds"])
    write_ply(tgt, d["target_ds"])
    np.savetxt(gt, d["T_target_source"])
    sys.exit(subprocess.run([exe, src, tgt, loops, gt]).returncode)
