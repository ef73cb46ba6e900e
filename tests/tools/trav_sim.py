"""Design experiment behind DESIGN.md section 4 ("what was measured and rejected", wide BVHs): dumps the cfg4 scene and the
oracle's LBVH, builds tests/tools/trav_sim.cpp and prints node visits / box tests / leaf tests per ray of the candidate
traversal schemes (binary near-first, + entry distance on the stack, 4- and 8-wide collapses, sorted / unsorted).
CPU only; test infrastructure (it uses the oracle's tree).   python tests/tools/trav_sim.py [n_rays] [sah | p16]
(the second argument swaps the LBVH for a sweep-SAH or a PLOC hierarchy over the same spheres: tree-quality experiment)"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import oracle as O  # noqa: E402
import vk_renderer_b200.scenes as scenes  # noqa: E402
from helpers import apply_scene  # noqa: E402

n_rays = sys.argv[1] if len(sys.argv) > 1 else "200000"
d = tempfile.mkdtemp()
sc = scenes.grid_spheres()
o = apply_scene(O, sc, fast=True).build_bvh()
sc.spheres.astype(np.float32).tofile(os.path.join(d, "spheres.bin"))
np.ascontiguousarray(o.bvh_nodes()).tofile(os.path.join(d, "nodes.bin"))
exe = os.path.join(d, "trav_sim")
subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "tools", "trav_sim.cpp")], check=True)
subprocess.run([exe, os.path.join(d, "spheres.bin"), os.path.join(d, "nodes.bin"), n_rays] + sys.argv[2:], check=True)
