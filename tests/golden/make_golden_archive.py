#!/usr/bin/env python
"""Golden fixture for the checkpoint wire format, produced by the REFERENCE itself:

    source oracle/_ref/ngs/env.sh && python tests/golden/make_golden_archive.py

  archive_wire.npz   for a real, a complex and a Mat<3,3> sparse matrix of a small netgen mesh: the CSR arrays and the bytes
                     SparseMatrix<TM>::DoArchive writes into ngcore's BinaryOutArchive (linalg/sparsematrix_impl.hpp:443-452,
                     reached through tests/golden/ref_helpers.cpp)
"""
import os
import subprocess
import sys
import sysconfig
import tempfile

import ngsolve
from ngsolve import *          # noqa: F401,F403
from netgen.csg import unit_cube

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
ngsolve.ngsglobals.msg_level = 0
pfx = os.path.join(ROOT, "oracle", "_ref", "ngs")
tmp = tempfile.mkdtemp()
so = os.path.join(tmp, "ref_helpers" + sysconfig.get_config_var("EXT_SUFFIX"))
subprocess.check_call([os.path.join(pfx, "bin", "ngscxx"), "-shared", os.path.join(HERE, "ref_helpers.cpp"), "-L" + os.path.join(pfx, "lib"),
                       "-lngla", "-lngstd", "-lngbla", "-L" + os.path.join(pfx, "lib", "python3.12", "site-packages", "netgen"), "-lngcore", "-o", so])
sys.path.insert(0, tmp)
import ref_helpers             # noqa: E402

mesh = Mesh(unit_cube.GenerateMesh(maxh=0.5))
store = {}


def add(name, mat):
    val, col, rowptr = mat.CSR()
    fn = os.path.join(tmp, name + ".bin")
    ref_helpers.archive(mat, fn)
    store[name + "_rowptr"] = np.array(rowptr, dtype=np.uint64)
    store[name + "_col"] = np.array(col, dtype=np.int32)
    store[name + "_val"] = np.array(val)
    store[name + "_bytes"] = np.frombuffer(open(fn, "rb").read(), dtype=np.uint8)
    print(name, type(mat).__name__, mat.height, mat.nze, len(store[name + "_bytes"]))


fes = H1(mesh, order=2)
u, v = fes.TnT()
add("real", BilinearForm(grad(u) * grad(v) * dx + u * v * dx).Assemble().mat)
fc = H1(mesh, order=2, complex=True)
u, v = fc.TnT()
add("complex", BilinearForm(grad(u) * grad(v) * dx + (1 + 2j) * u * v * dx).Assemble().mat)
fv = H1(mesh, order=1, dim=3)
u, v = fv.TnT()
add("block3", BilinearForm(InnerProduct(grad(u), grad(v)) * dx + InnerProduct(u, v) * dx + u[0] * v[1] * dx).Assemble().mat)
np.savez_compressed(os.path.join(HERE, "archive_wire.npz"), **store)
