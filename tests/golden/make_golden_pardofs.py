#!/usr/bin/env python
"""Golden fixture for the distributed index maps, produced by the REFERENCE's own ParallelDofs class:

    make -C oracle ref_pardofs && source oracle/_ref/ngs/env.sh && python tests/golden/make_golden_pardofs.py

oracle/_ref/ref_pardofs runs linalg/paralleldofs.cpp:20-108 (constructor: exchangedofs, ismasterdof, global_ndof) and
linalg/paralleldofs.hpp:213-334 (ReduceDofData / ScatterDofData / AllReduceDofData, the exchange behind the parallel Jacobi
diagonal, linalg/jacobi.cpp:60-61) on N ranks inside one process (threads as ranks, oracle/ref_pardofs/harness.cpp).
Two partitions:
  grid4   4 ranks, 2 x 2 element blocks of a 9 x 9 x 3 dof lattice: faces shared by 2 ranks, the centre line by all 4
  rand3   3 ranks, 240 dofs, every dof on a random non-empty subset of the ranks
Stored per rank r: dp_first / dp (the dist_procs table handed in), data, gid (global dof of every local dof), and the
reference outputs ex_first / ex_dofs / master / allred / scat, plus global_ndof.
"""
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_pardofs")


def run_reference(ranks):
    """ranks: list of (dp_first, dp, data) -> list of dicts of reference outputs"""
    np_ = len(ranks)
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fi, "wb") as f:
            f.write(struct.pack("i", np_))
            for first, dp, data in ranks:
                n = len(first) - 1
                f.write(struct.pack("i", n))
                f.write(np.asarray(first, dtype=np.int32).tobytes())
                f.write(np.asarray(dp, dtype=np.int32).tobytes())
                f.write(np.asarray(data, dtype=np.float64).tobytes())
        subprocess.check_call([BIN, fi, fo])
        raw = open(fo, "rb").read()
    out, off = [], 0
    for first, dp, data in ranks:
        n = len(first) - 1
        g = struct.unpack_from("q", raw, off)[0]; off += 8
        ex_first = np.frombuffer(raw, dtype=np.int32, count=np_ + 1, offset=off).copy(); off += 4 * (np_ + 1)
        ex_dofs = np.frombuffer(raw, dtype=np.int32, count=int(ex_first[-1]), offset=off).copy(); off += 4 * int(ex_first[-1])
        master = np.frombuffer(raw, dtype=np.uint8, count=n, offset=off).copy(); off += n
        allred = np.frombuffer(raw, dtype=np.float64, count=n, offset=off).copy(); off += 8 * n
        scat = np.frombuffer(raw, dtype=np.float64, count=n, offset=off).copy(); off += 8 * n
        out.append(dict(global_ndof=g, ex_first=ex_first, ex_dofs=ex_dofs, master=master, allred=allred, scat=scat))
    assert off == len(raw)
    return out


def tables(owners_of, nranks, rng):
    """owners_of: list over global dofs of the sorted rank lists -> per rank (gid, dp_first, dp, data)"""
    res = []
    for r in range(nranks):
        gid = np.array([g for g, ow in enumerate(owners_of) if r in ow], dtype=np.int64)      # local numbering ascending in the global id
        first = [0]
        dp = []
        for g in gid:
            others = [p for p in owners_of[g] if p != r]
            dp.extend(others)
            first.append(len(dp))
        res.append((gid, np.array(first, dtype=np.int32), np.array(dp, dtype=np.int32), rng.random(len(gid)) + 0.5))
    return res


rng = np.random.default_rng(2025)
store = {}
# ---- grid4
NX, NY, NZ = 9, 9, 3
owners = []
for z in range(NZ):
    for y in range(NY):
        for x in range(NX):
            ow = sorted({2 * b + a for a in range(2) for b in range(2) if 4 * a <= x <= 4 * a + 4 and 4 * b <= y <= 4 * b + 4})
            owners.append(ow)
# ---- rand3
owners_r = []
for g in range(240):
    k = rng.choice([1, 1, 1, 2, 2, 3])
    owners_r.append(sorted(rng.choice(3, size=k, replace=False).tolist()))
for name, ow, nr in (("grid4", owners, 4), ("rand3", owners_r, 3)):
    t = tables(ow, nr, rng)
    ref = run_reference([(first, dp, data) for (_, first, dp, data) in t])
    store[name + "_nranks"] = nr
    store[name + "_nglobal"] = len(ow)
    for r in range(nr):
        gid, first, dp, data = t[r]
        pre = "%s_r%d_" % (name, r)
        store[pre + "gid"], store[pre + "dp_first"], store[pre + "dp"], store[pre + "data"] = gid, first, dp, data
        for k, v in ref[r].items():
            store[pre + k] = v
        assert ref[r]["global_ndof"] == len(ow), (ref[r]["global_ndof"], len(ow))
        # the reference's all-reduced values are the sums over the sharers
        tot = np.zeros(len(ow))
        for q in range(nr):
            np.add.at(tot, t[q][0], t[q][3])
        assert np.allclose(ref[r]["allred"], tot[gid], rtol=1e-15)
    print(name, "ranks", nr, "local ndofs", [len(t[r][0]) for r in range(nr)], "exchange", [int(ref[r]["ex_first"][-1]) for r in range(nr)])
np.savez_compressed(os.path.join(HERE, "pardofs_reference.npz"), **store)
