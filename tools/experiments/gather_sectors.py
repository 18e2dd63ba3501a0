import sys, numpy as np
ROOT = __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/tests')
from firefoam_dev_b200 import meshgen as mg
from helpers import PlanView

def warp_sectors(pv, cols_of_row):
    """mean distinct 32-B sectors per warp gather request: request j of a warp = j-th entry of its 32 rows"""
    tot = req = 0
    for s in range(0, pv.N, 32):
        rows = range(s, min(pv.N, s + 32))
        maxn = max(len(cols_of_row[r]) for r in rows)
        for j in range(maxn):
            sec = {cols_of_row[r][j] >> 2 for r in rows if j < len(cols_of_row[r])}
            tot += len(sec); req += 1
    return tot / req

def cols(pv, resort):
    out = []
    for r in range(pv.N):
        nL, nT = int(pv.nLower[r]), int(pv.nTotal[r])
        c = [int(pv.col[pv.entry(r, j)]) for j in range(nT)]
        if resort:
            c = sorted(c[:nL]) + sorted(c[nL:])
        out.append(c)
    return out

s = mg.bcc_poly(*[int(a) for a in sys.argv[1:4]])
for ordering, name in ((0, "natural-order plan (Amul)"), (1, "multicolour plan (sweeps)")):
    for ren in (0, 1):
        pv = PlanView(ordering, s.addr, renumber=ren)
        a = warp_sectors(pv, cols(pv, False)); b = warp_sectors(pv, cols(pv, True))
        print(f"{name:28s} renumber={ren} colours={pv.nColours}: sectors/request face-order {a:5.2f}   column-sorted {b:5.2f}")
