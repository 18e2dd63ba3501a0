import sys, time, numpy as np, scipy.sparse as sp
from collections import deque
from block_multicolour_bricks import dic_factor, pcg_iters
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from firefoam_dev_b200 import meshgen as mg

def to_csr(s):
    a = s.addr; N = a.nCells
    return sp.coo_matrix((np.concatenate([s.diag, s.upper, s.upper]),
                   (np.concatenate([np.arange(N), a.lowerAddr, a.upperAddr]),
                    np.concatenate([np.arange(N), a.upperAddr, a.lowerAddr]))), shape=(N, N)).tocsr()

def aggregate(A, bs=32, mode="bfs"):
    """blocks of exactly bs cells (last one smaller): greedy growth from the lowest unassigned cell.
    mode bfs: plain breadth-first; mode conn: next cell = unassigned neighbour with most links into the block"""
    N = A.shape[0]; ip, ix = A.indptr, A.indices
    blk = -np.ones(N, dtype=np.int64); order = []
    nb = 0; cur = 0; seed = 0
    q = deque(); links = {}
    while len(order) < N:
        if mode == "bfs":
            if not q:
                while blk[seed] >= 0: seed += 1
                q.append(seed); blk[seed] = -2
            c = q.popleft()
        else:
            if not links:
                while blk[seed] >= 0: seed += 1
                links[seed] = 0
            c = max(links, key=lambda k: (links[k], -k)); del links[c]
        blk[c] = nb; order.append(c); cur += 1
        for k in range(ip[c], ip[c+1]):
            j = ix[k]
            if j == c: continue
            if mode == "bfs":
                if blk[j] == -1: blk[j] = -2; q.append(j)
            else:
                if blk[j] == -1: links[j] = links.get(j, 0) + 1
        if cur == bs:
            nb += 1; cur = 0
            if mode == "bfs":
                for j in q: blk[j] = -1
                q.clear()
            else: links.clear()
    return blk, (nb + (1 if cur else 0))

def block_colour_perm(A, blk, nb):
    N = A.shape[0]; ip, ix = A.indptr, A.indices
    adj = [set() for _ in range(nb)]
    for i in range(N):
        bi = blk[i]
        for k in range(ip[i], ip[i+1]):
            bj = blk[ix[k]]
            if bj != bi: adj[bi].add(bj)
    col = -np.ones(nb, dtype=np.int64)
    for b in range(nb):
        used = {col[x] for x in adj[b] if col[x] >= 0}
        c = 0
        while c in used: c += 1
        col[b] = c
    i = np.arange(N)
    return np.lexsort((i, blk, col[blk])), int(col.max()+1)

def run(name, A, b, perm, extra=""):
    t0=time.time(); Ap = A[perm][:, perm].tocsr(); it = pcg_iters(Ap, b[perm])
    print(f"{name:44s} iterations {it:5d} {extra} ({time.time()-t0:.0f}s)", flush=True)

what = sys.argv[1]
if what == "hex":
    dims = tuple(int(a) for a in sys.argv[2:5]); s = mg.hex_block(*dims)
else:
    dims = tuple(int(a) for a in sys.argv[2:5]); s = mg.bcc_poly(*dims)
A = to_csr(s); N = A.shape[0]; b = s.source
# base order: RCM for poly (what the plan does), natural for hex
if what == "poly":
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    base = reverse_cuthill_mckee(A, symmetric_mode=True)[::-1].copy()
    A = A[base][:, base].tocsr(); b = b[base]
run(f"{what}{dims} base order (DIC-exact class)", A, b, np.arange(N))
# greedy multicolour of cells (what DIC-class uses today)
ip, ix = A.indptr, A.indices
col = -np.ones(N, dtype=np.int64)
for i in range(N):
    used = {col[j] for j in ix[ip[i]:ip[i+1]] if col[j] >= 0}
    c = 0
    while c in used: c += 1
    col[i] = c
run("cell multicolour (today)", A, b, np.lexsort((np.arange(N), col)), f"colours {col.max()+1}")
for bs in (8, 32, 64):
    for mode in ("bfs", "conn"):
        blk, nb = aggregate(A, bs, mode)
        perm, nc = block_colour_perm(A, blk, nb)
        run(f"blocks of {bs} ({mode}), block-multicolour", A, b, perm, f"colours {nc}")
