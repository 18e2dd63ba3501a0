"""The device row structure (csrc/plan.cpp) validated on the CPU: the numpy emulation of the
kernels' row loops over the sliced-ELL arrays must reproduce the oracle's Amul and DIC
bit-for-bit (natural / level order) or to rounding (multicolour order)."""
import numpy as np
import pytest

from firefoam_dev_b200 import meshgen as mg
from firefoam_dev_b200.cases import StecklerHydrostatic
from firefoam_dev_b200.ldu import LduAddressing
from oracle import oracle as orc
from helpers import PlanView, random_ldu

NAT, MC, LEV = 0, 1, 2


def systems():
    yield "hex", mg.hex_block(7, 5, 6)
    yield "random", random_ldu(300, 5.0, seed=3)
    yield "random-sparse", random_ldu(257, 1.2, seed=5)   # empty rows, ragged


@pytest.mark.parametrize("name,s", list(systems()))
def test_structure_invariants(name, s):
    a = s.addr
    for ordering in (NAT, MC, LEV):
        P = PlanView(ordering, a)
        assert P.nTotal.sum() == 2 * a.nFaces
        m = P.faceOf >= 0
        # every face appears exactly twice, once in a lower group and once in an upper group
        assert np.array_equal(np.bincount(P.faceOf[m], minlength=a.nFaces), np.full(a.nFaces, 2))
        assert P.colourStart[0] == 0 and P.colourStart[-1] == a.nCells
        if P.perm.size:
            assert np.array_equal(np.sort(P.perm), np.arange(a.nCells))
            assert np.array_equal(P.iperm[P.perm], np.arange(a.nCells))
        for r in range(a.nCells):
            cols = [P.col[P.entry(r, j)] for j in range(P.nTotal[r])]
            assert all(c < r for c in cols[:P.nLower[r]]) and all(c > r for c in cols[P.nLower[r]:])
            fl = [P.faceOf[P.entry(r, j)] for j in range(P.nLower[r])]
            fu = [P.faceOf[P.entry(r, j)] for j in range(P.nLower[r], P.nTotal[r])]
            assert fl == sorted(fl) and fu == sorted(fu)
        # rows of one colour never neighbour each other
        if ordering != NAT:
            colour = np.searchsorted(P.colourStart, np.arange(a.nCells), side="right") - 1
            rl = P.iperm[a.lowerAddr]
            ru = P.iperm[a.upperAddr]
            assert np.all(colour[rl] != colour[ru])
            if ordering == LEV:   # elimination order preserved: owner before neighbour
                assert np.all(rl < ru)


@pytest.mark.parametrize("name,s", list(systems()))
def test_spmv_emulation_bit_exact_natural(name, s):
    P = PlanView(NAT, s.addr)
    x = np.random.default_rng(1).standard_normal(s.addr.nCells)
    y = P.spmv(s.diag, P.values(s.upper), x)
    assert np.array_equal(y, orc.amul(s, x)[0])


@pytest.mark.parametrize("name,s", list(systems()) + [("hex-larger", mg.hex_block(40, 9, 8))])
def test_symmetric_single_read_layout(name, s):
    """SymPlan: every face stored once; same row-sum order as OpenFOAM -> bit-identical Amul."""
    x = np.random.default_rng(6).standard_normal(s.addr.nCells)
    ref = orc.amul(s, x)[0]
    assert not PlanView(LEV, s.addr).symValid          # level orders keep the full-row ELL
    for ordering, tile in ((NAT, 0), (MC, 0), (MC, 32), (MC, 64)):
        P = PlanView(ordering, s.addr, tileRows=tile)
        assert P.symValid
        if tile and s.addr.nCells > 2 * tile:
            assert P.nTiles == -(-s.addr.nCells // tile)
        m = P.sym["uFace"] >= 0
        assert np.array_equal(np.bincount(P.sym["uFace"][m], minlength=s.addr.nFaces),
                              np.ones(s.addr.nFaces, dtype=np.int64))
        y = P.to_natural(P.spmv_sym(P.to_internal(s.diag), s.upper, P.to_internal(x)))
        if ordering == NAT:
            assert np.array_equal(y, ref)
        else:
            np.testing.assert_allclose(y, ref, rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("name,s", list(systems()))
def test_spmv_emulation_permuted(name, s):
    x = np.random.default_rng(2).standard_normal(s.addr.nCells)
    ref = orc.amul(s, x)[0]
    for ordering in (MC, LEV):
        P = PlanView(ordering, s.addr)
        y = P.to_natural(P.spmv(P.to_internal(s.diag), P.values(s.upper), P.to_internal(x)))
        np.testing.assert_allclose(y, ref, rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("name,s", list(systems()) + [("poly", mg.bcc_poly(5, 4, 6)), ("hex-larger", mg.hex_block(40, 9, 8))])
def test_16bit_column_offsets(name, s):
    """colBase[slice entry] + col16[entry] reproduces the 32-bit column wherever the slice entry is marked
    as fitting; wide entries (-1) keep the 32-bit column; padding lanes are never marked."""
    for ordering, ren in ((NAT, 0), (NAT, 1), (MC, 0), (MC, 1), (LEV, 0)):
        P = PlanView(ordering, s.addr, renumber=ren)
        if P.colBase.size == 0:
            continue
        assert P.colBase.size == P.nEntries // 32 and P.col16.size == P.nEntries
        for r in range(s.addr.nCells):
            for j in range(P.nTotal[r]):
                e = P.entry(r, j)
                b = P.colBase[e >> 5]
                if b >= 0:
                    assert b + int(P.col16[e]) == P.col[e]
        # on these small meshes every entry fits
        used = np.zeros(P.nEntries // 32, dtype=bool)
        for r in range(s.addr.nCells):
            for j in range(P.nTotal[r]):
                used[P.entry(r, j) >> 5] = True
        assert np.all(P.colBase[used] >= 0) and np.all(P.colBase[~used] == -1)


def test_rcm_renumbering_decision():
    """Auto mode renumbers a cache-hostile numbering (block-shuffled polyhedra) and leaves banded
    ones (lexicographic hex) alone; Levels (DIC-exact) is never renumbered."""
    poly = mg.bcc_poly(12, 12, 14)
    P = PlanView(NAT, poly.addr, renumber=-1)
    assert P.renumbered and P.spanUsed < 0.5 * P.spanNatural
    assert np.array_equal(np.sort(P.perm), np.arange(poly.addr.nCells))
    assert P.srValid and not P.symValid         # renumbered: the face-ordered single-read layout (SrPlan)
    assert PlanView(MC, poly.addr, renumber=-1).renumbered
    assert not PlanView(LEV, poly.addr, renumber=1).renumbered
    hexm = mg.hex_block(24, 20, 16)
    P = PlanView(NAT, hexm.addr, renumber=-1)
    assert not P.renumbered and P.perm.size == 0 and P.symValid


@pytest.mark.parametrize("name,s", list(systems()) + [("poly", mg.bcc_poly(5, 4, 6))])
def test_renumbered_natural_plan_is_bit_exact(name, s):
    """Rows in RCM order, but every row still sums its faces in ascending natural face order: Amul
    through the renumbered plan is bit-identical to the oracle's face loop."""
    a = s.addr
    P = PlanView(NAT, a, renumber=1)
    assert P.renumbered
    assert np.array_equal(np.sort(P.perm), np.arange(a.nCells))
    assert np.array_equal(P.iperm[P.perm], np.arange(a.nCells))
    for r in range(a.nCells):
        f = [P.faceOf[P.entry(r, j)] for j in range(P.nTotal[r])]
        assert f == sorted(f) and P.nLower[r] == 0
    x = np.random.default_rng(11).standard_normal(a.nCells)
    y = P.to_natural(P.spmv(P.to_internal(s.diag), P.values(s.upper), P.to_internal(x)))
    assert np.array_equal(y, orc.amul(s, x)[0])
    # single-read face-ordered layout (k_spmv_sr): every coefficient stored once, row sums still bit-exact
    assert P.srValid
    own = P.sr["ownFace"]
    assert np.array_equal(np.sort(own[own >= 0]), np.arange(a.nFaces))     # each face exactly once
    y = P.to_natural(P.spmv_sr(P.to_internal(s.diag), s.upper, P.to_internal(x)))
    assert np.array_equal(y, orc.amul(s, x)[0])
    # multicolour on top of the RCM base order: still a proper colouring
    Q = PlanView(MC, a, renumber=1)
    colour = np.searchsorted(Q.colourStart, np.arange(a.nCells), side="right") - 1
    assert np.all(colour[Q.iperm[a.lowerAddr]] != colour[Q.iperm[a.upperAddr]])


@pytest.mark.parametrize("name,s", list(systems()) + [("poly", mg.bcc_poly(5, 4, 6)), ("hex-larger", mg.hex_block(12, 9, 8))])
def test_tiled_multicolour_order(name, s):
    """Rows ordered (tile, colour, base position): same colouring, same IC0 preconditioner as the
    colour-major order (elimination order = colour order) -- only the storage order differs."""
    a = s.addr
    P0 = PlanView(MC, a)
    r = np.random.default_rng(8).standard_normal(a.nCells)
    v0 = P0.values(s.upper)
    rD0 = P0.dic_calc_rd(P0.to_internal(s.diag), v0)
    w0 = P0.to_natural(P0.dic_precondition(rD0, v0, P0.to_internal(r)))
    for tile in (32, 64, 128):
        P = PlanView(MC, a, tileRows=tile)
        C = P.nColours
        assert C == P0.nColours
        if a.nCells <= 2 * tile:
            assert P.nTiles == 1
            continue
        assert P.segStart.size == P.nTiles * C + 1 and P.segStart[0] == 0 and P.segStart[-1] == a.nCells
        assert np.array_equal(np.sort(P.perm), np.arange(a.nCells))
        # a tile holds `tile` consecutive base-order cells; inside a segment rows ascend in base order
        for t in range(P.nTiles):
            rows = np.arange(P.segStart[t * C], P.segStart[(t + 1) * C])
            assert np.array_equal(np.sort(P.perm[rows]), np.arange(t * tile, min(a.nCells, (t + 1) * tile)))
            for k in range(C):
                seg = P.perm[P.segStart[t * C + k]:P.segStart[t * C + k + 1]]
                assert np.all(np.diff(seg) > 0) and np.all(P.rowColour[P.segStart[t * C + k]:P.segStart[t * C + k + 1]] == k)
        # same colouring as the untiled plan
        col0 = np.empty(a.nCells, dtype=np.int64)
        col0[P0.perm] = np.searchsorted(P0.colourStart, np.arange(a.nCells), side="right") - 1
        assert np.array_equal(P.rowColour, col0[P.perm])
        # full-row ELL groups by COLOUR (elimination order), each group in ascending face order
        for row in range(a.nCells):
            cols = [P.col[P.entry(row, j)] for j in range(P.nTotal[row])]
            assert all(P.rowColour[c] < P.rowColour[row] for c in cols[:P.nLower[row]])
            assert all(P.rowColour[c] > P.rowColour[row] for c in cols[P.nLower[row]:])
        # identical preconditioner (bit for bit: same per-row operation order)
        v = P.values(s.upper)
        rD = P.dic_calc_rd(P.to_internal(s.diag), v)
        w = P.to_natural(P.dic_precondition(rD, v, P.to_internal(r)))
        assert np.array_equal(P.to_natural(rD), P0.to_natural(rD0)) and np.array_equal(w, w0)


@pytest.mark.parametrize("name,s", list(systems()))
def test_level_schedule_is_exact_dic(name, s):
    """DIC-exact: level-major order + OpenFOAM's per-row operation order == DICPreconditioner."""
    r = np.random.default_rng(4).standard_normal(s.addr.nCells)
    rD_ref, w_ref = orc.dic(s, r)
    P = PlanView(LEV, s.addr)
    val = P.values(s.upper)
    rD = P.dic_calc_rd(P.to_internal(s.diag), val)
    w = P.dic_precondition(rD, val, P.to_internal(r))
    assert np.array_equal(P.to_natural(rD), rD_ref)
    assert np.array_equal(P.to_natural(w), w_ref)


def test_multicolour_is_a_symmetric_ic0():
    """DIC-class (multicolour): M^-1 must be symmetric positive definite."""
    s = mg.hex_block(5, 4, 3)
    P = PlanView(MC, s.addr)
    assert P.nColours == 2     # structured hex is bipartite: red-black
    val = P.values(s.upper)
    rD = P.dic_calc_rd(P.to_internal(s.diag), val)
    n = s.addr.nCells
    M = np.stack([P.dic_precondition(rD, val, e) for e in np.eye(n)], 1)
    np.testing.assert_allclose(M, M.T, rtol=1e-12, atol=1e-14)
    assert np.linalg.eigvalsh(0.5 * (M + M.T)).min() > 0


def test_steckler_colourings():
    a = StecklerHydrostatic().addr
    assert PlanView(MC, a).nColours == 2
    assert PlanView(LEV, a).nColours == 30 + 15 + 20 - 2   # i+j+k hyperplanes


def test_interface_csr():
    subs = [mg.hex_block(6, 6, 4, 2, 2, 1, r) for r in range(4)]
    s = subs[0]
    for ordering in (NAT, MC):
        P = PlanView(ordering, s.addr)
        nSlots = sum(i.faceCells.size for i in s.addr.interfaces)
        assert P.patchStart[-1] == nSlots and P.slotRow.size == nSlots
        fc = np.concatenate([i.faceCells for i in s.addr.interfaces])
        rows = P.iperm[fc] if P.iperm.size else fc
        assert np.array_equal(P.slotRow, rows)
        assert np.array_equal(np.sort(P.bSlot), np.arange(nSlots))
        assert np.all(np.diff(P.bRow) > 0)
        for b in range(P.bRow.size):
            sl = P.bSlot[P.bStart[b]:P.bStart[b + 1]]
            assert np.all(P.slotRow[sl] == P.bRow[b]) and np.all(np.diff(sl) > 0)
        # the corner cell touches two processor patches
        assert (np.diff(P.bStart) == 2).any()


def test_rejects_bad_addressing():
    with pytest.raises(ValueError):
        PlanView(NAT, LduAddressing(3, [1, 0], [2, 1]))      # not upper-triangular order
    with pytest.raises(ValueError):
        PlanView(NAT, LduAddressing(3, [1], [1]))            # l == u
    with pytest.raises(ValueError):
        PlanView(NAT, LduAddressing(3, [0], [3]))            # out of range


def test_empty_and_single_cell():
    P = PlanView(NAT, LduAddressing(0, [], []))
    assert P.nEntries == 0
    P = PlanView(LEV, LduAddressing(1, [], []))
    assert P.nColours == 1 and P.nTotal[0] == 0
