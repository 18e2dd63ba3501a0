"""decomposePar-style partitioning (csrc/decompose.cpp, meshgen hex sub-blocks): sub-meshes,
processor-interface pairing, and N-rank oracle == 1-rank oracle."""
import numpy as np
import pytest

from firefoam_dev_b200 import meshgen as mg
from oracle import oracle as orc
from helpers import random_ldu


def hex_xyz(nx, ny, nz):
    c = np.arange(nx * ny * nz)
    return np.stack([c % nx, (c // nx) % ny, c // (nx * ny)], 1).astype(float) + 0.5


def test_direct_blocks_equal_generic_decompose():
    dims, procs = (12, 10, 8), (2, 2, 2)
    s = mg.hex_block(*dims)
    c2p = mg.partition_hierarchical(hex_xyz(*dims), procs)
    assert np.array_equal(c2p, mg.partition_simple(hex_xyz(*dims), procs))
    subs2 = mg.decompose(s, c2p, 8)
    for r in range(8):
        a, b = mg.hex_block(*dims, *procs, r), subs2[r]
        assert np.array_equal(a.addr.lowerAddr, b.addr.lowerAddr)
        assert np.array_equal(a.addr.upperAddr, b.addr.upperAddr)
        assert np.array_equal(a.upper, b.upper)
        np.testing.assert_allclose(a.diag, b.diag, rtol=1e-14)
        np.testing.assert_allclose(a.source, b.source, rtol=1e-11, atol=1e-18)
        assert [i.neighbProcNo for i in a.addr.interfaces] == [i.neighbProcNo for i in b.addr.interfaces]
        for k in range(len(a.bou)):
            assert np.array_equal(a.addr.interfaces[k].faceCells, b.addr.interfaces[k].faceCells)
            assert np.array_equal(a.bou[k], b.bou[k])


def test_uneven_split_counts():
    sz = [mg.hex_sizes(7, 5, 3, 3, 2, 1, r)[0] for r in range(6)]
    assert sum(sz) == 7 * 5 * 3 and sz[0] == 3 * 3 * 3 and sz[2] == 2 * 3 * 3


@pytest.mark.parametrize("pre", ["diagonal", "none"])
def test_multirank_oracle_matches_single_rank(pre):
    s = mg.hex_block(12, 10, 8)
    psi1 = np.zeros(s.addr.nCells)
    p1 = orc.pcg_solve(s, psi1, pre, 1e-8, 0.0, 5000)
    for procs in ((2, 1, 1), (2, 2, 1), (2, 2, 2)):
        R = procs[0] * procs[1] * procs[2]
        subs = [mg.hex_block(12, 10, 8, *procs, r) for r in range(R)]
        psis = [np.zeros(x.addr.nCells) for x in subs]
        pR = orc.pcg_solve(subs, psis, pre, 1e-8, 0.0, 5000)
        assert pR.nIterations == p1.nIterations
        c2p = mg.partition_hierarchical(hex_xyz(12, 10, 8), procs)
        full = np.empty_like(psi1)
        for r in range(R):
            full[np.nonzero(c2p == r)[0]] = psis[r]
        np.testing.assert_allclose(full, psi1, rtol=1e-6, atol=1e-9)


def test_rcb_on_irregular_graph_amul():
    s = random_ldu(500, 6.0, seed=11)
    xyz = np.random.default_rng(0).uniform(size=(500, 3))
    c2p = mg.partition_rcb(xyz, 5)
    assert np.bincount(c2p).tolist() == [100] * 5
    subs = mg.decompose(s, c2p, 5)
    x = np.random.default_rng(1).standard_normal(500)
    ys = orc.amul(subs, [x[sub.cells] for sub in subs])
    y = np.empty(500)
    for sub, yy in zip(subs, ys):
        y[sub.cells] = yy
    np.testing.assert_allclose(y, orc.amul(s, x)[0], rtol=1e-13, atol=1e-14)
    # both sides of every processor patch list the same global faces in the same order
    L = {}
    for p, sub in enumerate(subs):
        for k, itf in enumerate(sub.addr.interfaces):
            L[(p, itf.neighbProcNo)] = sub.bou[k]
    for (p, q), b in L.items():
        assert np.array_equal(b, L[(q, p)])


def test_decomposed_submeshes_carry_the_laplacian_inputs():
    """decompose() hands every rank the fvm::laplacian inputs of its sub-mesh (bench.py --workload poly at
    N > 1): assembling them with the oracle reproduces the rank's share of the global matrix -- upper
    bit for bit, diag to rounding (the cut faces enter through the processor patches' internalCoeffs)."""
    full = mg.bcc_poly(6, 5, 7)
    subs = mg.decompose(full, mg.partition_rcb(full.xyz, 4), 4)
    assert sum(s.addr.nCells for s in subs) == full.addr.nCells
    for s in subs:
        a = s.addr
        up, dg = orc.laplacian_assemble(a.lowerAddr, a.upperAddr, a.nCells, s.gamma_f, s.magSf, s.deltaCoeffs,
                                        s.sign, s.diag0)
        assert np.array_equal(up, s.upper)
        np.testing.assert_allclose(dg, s.diag, rtol=1e-13)
        assert len(s.bou) == len(a.interfaces) > 0
