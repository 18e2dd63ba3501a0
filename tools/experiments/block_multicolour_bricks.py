import sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from firefoam_dev_b200 import meshgen as mg

def dic_factor(A):
    """DIC diagonal for the ordering of A (csr, symmetric): Dt_i = D_i - sum_{j<i} a_ij^2/Dt_j"""
    L = sp.tril(A, -1).tocsr(); D = A.diagonal().copy(); Dt = D.copy()
    ip, ix, v = L.indptr, L.indices, L.data
    for i in range(A.shape[0]):
        s = 0.0
        for k in range(ip[i], ip[i+1]): s += v[k]*v[k]/Dt[ix[k]]
        Dt[i] = D[i]-s
    return L, Dt

def pcg_iters(A, b, tol=1e-6, maxit=5000):
    L, Dt = dic_factor(A)
    Lo = (L + sp.diags(Dt)).tocsr(); Up = Lo.T.tocsr()
    x = np.zeros_like(b); r = b.copy(); nf = 2*np.abs(b).sum()
    p = None; rho_old = 1.0
    for it in range(1, maxit+1):
        y = spl.spsolve_triangular(Lo, r, lower=True)
        z = spl.spsolve_triangular(Up, Dt*y, lower=False)
        rho = z@r
        p = z if p is None else z + (rho/rho_old)*p
        w = A@p; alpha = rho/(w@p)
        x += alpha*p; r -= alpha*w; rho_old = rho
        if np.abs(r).sum()/nf < tol: return it
    return maxit

def orderings(nx, ny, nz):
    N = nx*ny*nz
    i = np.arange(N); ix = i % nx; iy = (i//nx) % ny; iz = i//(nx*ny)
    out = {"natural": np.arange(N)}
    rb = (ix+iy+iz) & 1
    out["red-black"] = np.lexsort((i, rb))
    for bx,by,bz in [(2,1,1),(4,1,1),(8,1,1),(16,1,1),(2,2,2),(4,4,1),(4,4,4),(8,8,1),(8,8,8)]:
        cx, cy, cz = ix//bx, iy//by, iz//bz
        col = (cx+cy+cz) & 1                      # blocks coloured red-black
        blk = cx + (nx//bx+1)*(cy + (ny//by+1)*cz)
        out[f"block {bx}x{by}x{bz} (2 colours)"] = np.lexsort((i, blk, col))   # colour-major, block by block, natural inside
    return out

if __name__ == "__main__":
    dims = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (48, 48, 48)
    s = mg.hex_block(*dims)
    a = s.addr; N = a.nCells
    A = sp.coo_matrix((np.concatenate([s.diag, s.upper, s.upper]),
                       (np.concatenate([np.arange(N), a.lowerAddr, a.upperAddr]),
                        np.concatenate([np.arange(N), a.upperAddr, a.lowerAddr]))), shape=(N, N)).tocsr()
    for name, perm in orderings(*dims).items():
        t0 = time.time()
        Ap = A[perm][:, perm].tocsr()
        it = pcg_iters(Ap, s.source[perm])
        print(f"{dims} {name:32s} iterations {it:5d}   ({time.time()-t0:.1f}s)", flush=True)
