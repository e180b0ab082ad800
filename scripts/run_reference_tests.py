#!/usr/bin/env python
"""Run the REFERENCE's own test files against oracle/_ref (the unmodified reference sources compiled against the
OpenMesh stand-in, `make -C oracle ref`), to pin the oracle the goldens come from. Build container only
(needs /root/reference); test infrastructure, never on the product path.

    OMP_NUM_THREADS=4 python scripts/run_reference_tests.py [> profiles/r02_reference_tests.txt]

The reference tests import two packages this image lacks; small stand-ins are generated into a temporary directory:
  trimesh   creation.icosphere(n) / creation.icosahedron(): the same construction (icosahedron table, midpoint
            subdivision, projection to the unit sphere) from trimem_b200.meshgen, which SURVEY App. D matched to
            trimesh's vertex order through the hard-coded trajectory of tests/test_sampling.py
  autograd  grad(f): 8th-order central differences (the tests compare gradients at 1e-10; the step error is ~1e-13)
Files that need h5py / meshio / the mc_app command line (test_output.py, test_integration.py) are not run.
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
FILES = ["test_properties.py", "test_energy.py", "test_gradient.py", "test_neighbours.py", "test_flips.py", "test_sampling.py"]

TRIMESH = '''
import sys
sys.path.insert(0, %r)
from types import SimpleNamespace
from trimem_b200 import meshgen

class creation:
    @staticmethod
    def icosphere(n=3):
        v, f = meshgen.icosphere(n)
        return SimpleNamespace(vertices=v, faces=f)
    @staticmethod
    def icosahedron():
        v, f = meshgen.icosphere(0)
        return SimpleNamespace(vertices=v, faces=f)
''' % ROOT

AUTOGRAD = '''
import numpy as _np
C = (4.0 / 5.0, -1.0 / 5.0, 4.0 / 105.0, -1.0 / 280.0)

def grad(f):
    def df(x):
        x = _np.array(x, dtype=float)
        g = _np.zeros_like(x)
        h = 1.0e-2
        for idx in _np.ndindex(x.shape):
            acc = 0.0
            for k, c in enumerate(C, start=1):
                xp = x.copy(); xp[idx] += k * h
                xm = x.copy(); xm[idx] -= k * h
                acc += c * (f(xp) - f(xm))
            g[idx] = acc / h
        return g
    return df
'''


def main():
    if not os.path.isdir(os.path.join(REF, "tests")):
        sys.exit("no /root/reference here: this script runs in the build container only")
    if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "trimem")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"])
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "autograd"))
        with open(os.path.join(tmp, "trimesh.py"), "w") as fh:
            fh.write(TRIMESH)
        with open(os.path.join(tmp, "autograd", "__init__.py"), "w") as fh:
            fh.write(AUTOGRAD)
        with open(os.path.join(tmp, "autograd", "numpy.py"), "w") as fh:
            fh.write("from numpy import *\nimport numpy as _n\nlinalg = _n.linalg\n")
        # the reference's Python package (trimem.mc) next to the oracle build of trimem.core
        pkg = os.path.join(tmp, "pkg", "trimem")
        os.makedirs(pkg)
        for name in os.listdir(os.path.join(ROOT, "oracle", "_ref", "trimem")):
            os.symlink(os.path.join(ROOT, "oracle", "_ref", "trimem", name), os.path.join(pkg, name))
        os.symlink(os.path.join(REF, "src", "trimem", "mc"), os.path.join(pkg, "mc"))
        env = dict(os.environ)
        env["PYTHONPATH"] = os.pathsep.join([os.path.join(tmp, "pkg"), tmp, os.path.join(REF, "tests")])
        env.setdefault("OMP_NUM_THREADS", "4")
        cmd = [sys.executable, "-m", "pytest", "-q" if "-v" not in sys.argv else "-v", "-rs", "-p", "no:cacheprovider", "--rootdir", tmp] + \
              [os.path.join(REF, "tests", f) for f in FILES]
        print("# reference tests against oracle/_ref:", " ".join(FILES), flush=True)
        rc = subprocess.call(cmd, env=env, cwd=tmp)
    sys.exit(rc)


if __name__ == "__main__":
    main()
