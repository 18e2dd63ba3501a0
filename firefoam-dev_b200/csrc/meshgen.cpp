// meshgen.cpp -- synthetic workloads of BASELINE.json as OpenFOAM-style LDU data.
//
// Host-only C++ (libb200mesh.so, no CUDA): the harness equivalent of blockMesh + decomposePar +
// the field set-up of solver/pEqn.H:3-4 (rhorAUf) for synthetic cases.  It produces exactly what
// fvMatrix<scalar>::solveSegregated hands to lduMatrix::solver: lowerAddr /
// upperAddr in upper-triangular order, diag WITH boundary internalCoeffs, upper,
// processor-interface faceCells + interfaceBouCoeffs, totalSource -- plus the face fields
// (gamma_f, magSf, deltaCoeffs) that fvm::laplacian consumes, so the assembly kernel can be
// checked against the same matrix.
//
// Workload generators are NOT the oracle: they only make inputs.  The CPU restatement of the
// reference algorithm lives in oracle/.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// uniform in (-1, 1), counter-based
inline double uni(uint64_t seed, uint64_t counter) {
    uint64_t r = splitmix64(seed ^ (counter * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull));
    return ((double)(r >> 11) + 0.5) * (2.0 / 9007199254740992.0) - 1.0;
}

struct Range { int lo, hi; };
// decomposePar `simple`/`hierarchical` on an evenly sorted coordinate: n groups of floor(N/n)
// cells, the first N mod n groups get one extra.
inline Range split(int N, int n, int g) {
    int base = N / n, extra = N % n;
    int lo = g * base + std::min(g, extra);
    int hi = lo + base + (g < extra ? 1 : 0);
    return {lo, hi};
}

struct HexSpec {
    int NX, NY, NZ, PX, PY, PZ, rank;
    uint64_t seed;
    double h, gamma0, psiOverDt;  // psiOverDt = psi/dt (ddt coefficient per unit volume)
};

struct HexBlock {
    HexSpec s;
    Range rx, ry, rz;
    int gx, gy, gz, lnx, lny, lnz;
    int32_t N;
    int32_t F;
    int nIf;
    int nbrRank[6];
    int nbrDir[6];    // 0:-z 1:-y 2:-x 3:+x 4:+y 5:+z
    int nbrFaces[6];
};

const double kPi = 3.14159265358979323846;

inline int64_t gid(const HexSpec& s, int i, int j, int k) {
    return (int64_t)i + (int64_t)s.NX * ((int64_t)j + (int64_t)s.NY * k);
}
inline double gamma_cell(const HexSpec& s, int i, int j, int k) {
    const double x = (i + 0.5) * s.h, y = (j + 0.5) * s.h, z = (k + 0.5) * s.h;
    const double xi = uni(s.seed, (uint64_t)gid(s, i, j, k));
    return s.gamma0 * (1.0 + 0.9 * std::sin(2 * kPi * x) * std::sin(2 * kPi * y) * std::sin(2 * kPi * z)) *
           std::pow(10.0, 0.5 * xi);
}
inline double xstar_cell(const HexSpec& s, int i, int j, int k) {
    const double x = (i + 0.5) * s.h, y = (j + 0.5) * s.h, z = (k + 0.5) * s.h;
    const double xi = uni(s.seed ^ 0xABCDEF1234567ull, (uint64_t)gid(s, i, j, k));
    return std::sin(kPi * x) * std::cos(2 * kPi * y) * std::sin(3 * kPi * z) + 0.5 * x + 0.01 * xi;
}

bool make_block(const HexSpec& s, HexBlock& b) {
    if (s.NX < 1 || s.NY < 1 || s.NZ < 1 || s.PX < 1 || s.PY < 1 || s.PZ < 1) return false;
    if (s.PX > s.NX || s.PY > s.NY || s.PZ > s.NZ) return false;
    if (s.rank < 0 || s.rank >= s.PX * s.PY * s.PZ) return false;
    b.s = s;
    b.gx = s.rank % s.PX;
    b.gy = (s.rank / s.PX) % s.PY;
    b.gz = s.rank / (s.PX * s.PY);
    b.rx = split(s.NX, s.PX, b.gx);
    b.ry = split(s.NY, s.PY, b.gy);
    b.rz = split(s.NZ, s.PZ, b.gz);
    b.lnx = b.rx.hi - b.rx.lo;
    b.lny = b.ry.hi - b.ry.lo;
    b.lnz = b.rz.hi - b.rz.lo;
    int64_t N = (int64_t)b.lnx * b.lny * b.lnz;
    int64_t F = (int64_t)(b.lnx - 1) * b.lny * b.lnz + (int64_t)b.lnx * (b.lny - 1) * b.lnz +
                (int64_t)b.lnx * b.lny * (b.lnz - 1);
    if (N > 0x7fffffff || F > 0x7fffffff) return false;
    b.N = (int32_t)N;
    b.F = (int32_t)F;
    b.nIf = 0;
    auto add = [&](bool present, int dir, int nbr, int nf) {
        if (!present) return;
        b.nbrRank[b.nIf] = nbr;
        b.nbrDir[b.nIf] = dir;
        b.nbrFaces[b.nIf] = nf;
        b.nIf++;
    };
    // ascending neighbour rank == processor patch order of decomposePar
    add(b.gz > 0, 0, s.rank - s.PX * s.PY, b.lnx * b.lny);
    add(b.gy > 0, 1, s.rank - s.PX, b.lnx * b.lnz);
    add(b.gx > 0, 2, s.rank - 1, b.lny * b.lnz);
    add(b.gx < s.PX - 1, 3, s.rank + 1, b.lny * b.lnz);
    add(b.gy < s.PY - 1, 4, s.rank + s.PX, b.lnx * b.lnz);
    add(b.gz < s.PZ - 1, 5, s.rank + s.PX * s.PY, b.lnx * b.lny);
    return true;
}

}  // namespace

extern "C" {

// sizes[0]=nCells, [1]=nFaces, [2]=nIfaces, [3..8]=faces per interface, [9..14]=neighbour ranks
int b200mesh_hex_sizes(int NX, int NY, int NZ, int PX, int PY, int PZ, int rank, int64_t* sizes) {
    HexSpec s{NX, NY, NZ, PX, PY, PZ, rank, 0, 1.0, 1.0, 0.0};
    HexBlock b;
    if (!make_block(s, b)) return 1;
    sizes[0] = b.N;
    sizes[1] = b.F;
    sizes[2] = b.nIf;
    for (int k = 0; k < 6; ++k) {
        sizes[3 + k] = k < b.nIf ? b.nbrFaces[k] : 0;
        sizes[9 + k] = k < b.nIf ? b.nbrRank[k] : -1;
    }
    return 0;
}

// Fills the p_rgh-shaped system of SURVEY.md 8d config 3/4 for one rank's sub-block of the global
// NX x NY x NZ box (cube cells of edge h):
//   gamma_c = gamma0 (1 + 0.9 sin2pix sin2piy sin2piz) 10^(0.5 xi_c), gamma_f = (gamma_o+gamma_n)/2
//   A = diag(psi V/dt) - laplacian(gamma)  [SPD M-matrix form of `- fvm::laplacian`, pEqn.H:32]
//   y = max patch: fixed value 0 (deltaCoeff 2/h); every other physical patch: zero flux
//   source = A x*, x* smooth + 1 % noise
// Outputs (caller-allocated):
//   lower, upper_addr [F]; gamma_f, magSf, delta [F]  (inputs of b200_assemble_laplacian, sign=-1)
//   diag0 [N]  = ddt diagonal + boundary internalCoeffs of ALL patches (incl. processor patches)
//   diag [N], upper [F] = the assembled matrix (what the solver receives)
//   source [N], xstar [N]
//   faceCells[k] [nFaces_k], bouCoeffs[k] [nFaces_k] for k < nIfaces (pointers in arrays of 6)
int b200mesh_hex_fill(int NX, int NY, int NZ, int PX, int PY, int PZ, int rank, uint64_t seed, double h,
                      double gamma0, double psiOverDt, int32_t* lower, int32_t* upper_addr,
                      double* gamma_f, double* magSf, double* delta, double* diag0, double* diag,
                      double* upper, double* source, double* xstar, int32_t* const* faceCells,
                      double* const* bouCoeffs) {
    HexSpec s{NX, NY, NZ, PX, PY, PZ, rank, seed, h, gamma0, psiOverDt};
    HexBlock b;
    if (!make_block(s, b)) return 1;
    const int lnx = b.lnx, lny = b.lny, lnz = b.lnz;
    const double S = h * h, dC = 1.0 / h, V = h * h * h;
    auto lid = [&](int i, int j, int k) { return (int32_t)(i + lnx * (j + lny * k)); };

    // face offsets: count owner faces per cell in order -> prefix.  Faces of cell (i,j,k) in
    // ascending neighbour order: +x, +y, +z (those inside the block).
    std::vector<int64_t> fstart((size_t)lnz + 1, 0);  // per k-plane offsets (all rows alike in count)
    for (int k = 0; k < lnz; ++k) {
        int64_t perPlane = (int64_t)(lnx - 1) * lny + (int64_t)lnx * (lny - 1) + (k < lnz - 1 ? (int64_t)lnx * lny : 0);
        fstart[k + 1] = fstart[k] + perPlane;
    }
#pragma omp parallel for schedule(static)
    for (int k = 0; k < lnz; ++k) {
        int64_t f = fstart[k];
        const int K = b.rz.lo + k;
        for (int j = 0; j < lny; ++j) {
            const int J = b.ry.lo + j;
            for (int i = 0; i < lnx; ++i) {
                const int I = b.rx.lo + i;
                const int32_t c = lid(i, j, k);
                const double gc = gamma_cell(s, I, J, K);
                const double xc = xstar_cell(s, I, J, K);
                double d0 = psiOverDt * V;
                if (J == NY - 1) d0 += gc * S * (2.0 * dC);  // fixed-value top patch, gamma_b = gamma_c
                xstar[c] = xc;
                diag0[c] = d0;
                auto face = [&](int32_t n, double gn) {
                    lower[f] = c;
                    upper_addr[f] = n;
                    gamma_f[f] = 0.5 * (gc + gn);
                    magSf[f] = S;
                    delta[f] = dC;
                    upper[f] = -1.0 * (dC * (gamma_f[f] * S));
                    ++f;
                };
                if (i < lnx - 1) face(lid(i + 1, j, k), gamma_cell(s, I + 1, J, K));
                if (j < lny - 1) face(lid(i, j + 1, k), gamma_cell(s, I, J + 1, K));
                if (k < lnz - 1) face(lid(i, j, k + 1), gamma_cell(s, I, J, K + 1));
            }
        }
    }
    // interfaces
    for (int p = 0; p < b.nIf; ++p) {
        const int dir = b.nbrDir[p];
        int32_t* fc = faceCells[p];
        double* bc = bouCoeffs[p];
        int64_t n = 0;
        auto emit = [&](int i, int j, int k, int di, int dj, int dk) {
            const int I = b.rx.lo + i, J = b.ry.lo + j, K = b.rz.lo + k;
            const double gf = 0.5 * (gamma_cell(s, I, J, K) + gamma_cell(s, I + di, J + dj, K + dk));
            const double coeff = dC * (gf * S);
            fc[n] = lid(i, j, k);
            bc[n] = coeff;            // interfaceBouCoeffs (Amul subtracts bou*psi_nbr)
            diag0[lid(i, j, k)] += coeff;  // interfaceIntCoeffs folded into diag
            ++n;
        };
        if (dir == 0 || dir == 5) {
            const int k = dir == 0 ? 0 : lnz - 1, dk = dir == 0 ? -1 : 1;
            for (int j = 0; j < lny; ++j) for (int i = 0; i < lnx; ++i) emit(i, j, k, 0, 0, dk);
        } else if (dir == 1 || dir == 4) {
            const int j = dir == 1 ? 0 : lny - 1, dj = dir == 1 ? -1 : 1;
            for (int k = 0; k < lnz; ++k) for (int i = 0; i < lnx; ++i) emit(i, j, k, 0, dj, 0);
        } else {
            const int i = dir == 2 ? 0 : lnx - 1, di = dir == 2 ? -1 : 1;
            for (int k = 0; k < lnz; ++k) for (int j = 0; j < lny; ++j) emit(i, j, k, di, 0, 0);
        }
    }
    // diag = diag0 - sum(upper) over the faces of each cell, in face order (negSumDiag of the
    // negated matrix); source = A x* including the processor-coupled neighbours
    for (int32_t c = 0; c < b.N; ++c) diag[c] = 0.0;
    for (int64_t f = 0; f < b.F; ++f) {
        diag[lower[f]] -= upper[f];
        diag[upper_addr[f]] -= upper[f];
    }
    for (int32_t c = 0; c < b.N; ++c) diag[c] += diag0[c];
    for (int32_t c = 0; c < b.N; ++c) source[c] = diag[c] * xstar[c];
    for (int64_t f = 0; f < b.F; ++f) {
        source[upper_addr[f]] += upper[f] * xstar[lower[f]];
        source[lower[f]] += upper[f] * xstar[upper_addr[f]];
    }
    for (int p = 0; p < b.nIf; ++p) {
        const int dir = b.nbrDir[p];
        const int di = dir == 2 ? -1 : dir == 3 ? 1 : 0, dj = dir == 1 ? -1 : dir == 4 ? 1 : 0,
                  dk = dir == 0 ? -1 : dir == 5 ? 1 : 0;
        for (int n = 0; n < b.nbrFaces[p]; ++n) {
            const int32_t c = faceCells[p][n];
            const int i = c % lnx, j = (c / lnx) % lny, k = c / (lnx * lny);
            const double xn = xstar_cell(s, b.rx.lo + i + di, b.ry.lo + j + dj, b.rz.lo + k + dk);
            source[c] -= bouCoeffs[p][n] * xn;
        }
    }
    return 0;
}


// ---------------------------------------------------------------------------------------------
// BCC-lattice Voronoi mesh (truncated octahedra, 14 faces per interior cell): SURVEY.md 8d config 5.
// Sub-lattice A at (i,j,k)a, B at (i+1/2,j+1/2,k+1/2)a; 8 hexagonal faces to the other sub-lattice
// (area 3*sqrt(3)/16 a^2, distance sqrt(3)/2 a), 6 square faces to the same sub-lattice (area a^2/8,
// distance a); cell volume a^3/2.  Cells are numbered in Morton order of the doubled integer
// coordinates, then shuffled (seeded) inside blocks of `shuffleBlock` cells, so lowerAddr segment
// lengths are irregular (0..14) and strides non-constant.  Faces are emitted in upper-triangular
// order.  Same gamma / boundary / manufactured-solution recipe as the hex workload.
namespace {
inline uint64_t part1by2(uint64_t x) {
    x &= 0x1fffff;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
struct Bcc {
    int nx, ny, nz;
    int64_t N;
    std::vector<int32_t> siteOfCell, cellOfSite;   // site = 2*(i + nx*(j + ny*k)) + sub
    int neighbours(int64_t site, int64_t* out, int* isHex) const {
        const int sub = (int)(site & 1);
        const int64_t q = site >> 1;
        const int i = (int)(q % nx), j = (int)((q / nx) % ny), k = (int)(q / ((int64_t)nx * ny));
        int n = 0;
        auto add = [&](int I, int J, int K, int S, int hex) {
            if (I < 0 || J < 0 || K < 0 || I >= nx || J >= ny || K >= nz) return;
            out[n] = 2 * ((int64_t)I + (int64_t)nx * (J + (int64_t)ny * K)) + S;
            isHex[n] = hex;
            ++n;
        };
        for (int d = 0; d < 8; ++d) {
            const int di = d & 1, dj = (d >> 1) & 1, dk = (d >> 2) & 1;
            if (sub == 0) add(i - 1 + di, j - 1 + dj, k - 1 + dk, 1, 1);
            else add(i + di, j + dj, k + dk, 0, 1);
        }
        add(i - 1, j, k, sub, 0); add(i + 1, j, k, sub, 0);
        add(i, j - 1, k, sub, 0); add(i, j + 1, k, sub, 0);
        add(i, j, k - 1, sub, 0); add(i, j, k + 1, sub, 0);
        return n;
    }
};
bool make_bcc(int nx, int ny, int nz, uint64_t seed, int shuffleBlock, Bcc& b) {
    if (nx < 1 || ny < 1 || nz < 1) return false;
    const int64_t N = 2ll * nx * ny * nz;
    if (N > 0x7fffffff) return false;
    b.nx = nx; b.ny = ny; b.nz = nz; b.N = N;
    std::vector<std::pair<uint64_t, int32_t>> key((size_t)N);
#pragma omp parallel for schedule(static)
    for (int64_t site = 0; site < N; ++site) {
        const int sub = (int)(site & 1);
        const int64_t q = site >> 1;
        const uint64_t X = 2 * (q % nx) + sub, Y = 2 * ((q / nx) % ny) + sub, Z = 2 * (q / ((int64_t)nx * ny)) + sub;
        key[site] = {part1by2(X) | (part1by2(Y) << 1) | (part1by2(Z) << 2), (int32_t)site};
    }
    std::sort(key.begin(), key.end());
    b.siteOfCell.resize((size_t)N);
    for (int64_t c = 0; c < N; ++c) b.siteOfCell[c] = key[c].second;
    if (shuffleBlock > 1)
        for (int64_t b0 = 0; b0 < N; b0 += shuffleBlock) {
            const int64_t n = std::min<int64_t>(shuffleBlock, N - b0);
            for (int64_t i = n - 1; i > 0; --i) {
                const uint64_t r = splitmix64(seed ^ (uint64_t)(b0 + i) * 0x9E3779B97F4A7C15ull) % (uint64_t)(i + 1);
                std::swap(b.siteOfCell[b0 + i], b.siteOfCell[b0 + (int64_t)r]);
            }
        }
    b.cellOfSite.resize((size_t)N);
    for (int64_t c = 0; c < N; ++c) b.cellOfSite[b.siteOfCell[c]] = (int32_t)c;
    return true;
}
inline void site_xyz(const Bcc& b, int64_t site, double a, double* p) {
    const int sub = (int)(site & 1);
    const int64_t q = site >> 1;
    p[0] = ((double)(q % b.nx) + 0.5 * sub) * a;
    p[1] = ((double)((q / b.nx) % b.ny) + 0.5 * sub) * a;
    p[2] = ((double)(q / ((int64_t)b.nx * b.ny)) + 0.5 * sub) * a;
}
Bcc* g_bcc = nullptr;
}  // namespace

// two-phase: sizes (builds and caches the numbering), then fill
int b200mesh_bcc_sizes(int nx, int ny, int nz, uint64_t seed, int shuffleBlock, int64_t* sizes) {
    delete g_bcc;
    g_bcc = new Bcc();
    if (!make_bcc(nx, ny, nz, seed, shuffleBlock, *g_bcc)) { delete g_bcc; g_bcc = nullptr; return 1; }
    const Bcc& b = *g_bcc;
    int64_t F = 0;
#pragma omp parallel for schedule(static) reduction(+ : F)
    for (int64_t c = 0; c < b.N; ++c) {
        int64_t nb[14]; int hx[14];
        const int n = b.neighbours(b.siteOfCell[c], nb, hx);
        for (int k = 0; k < n; ++k) if (b.cellOfSite[nb[k]] > c) ++F;
    }
    if (F > 0x7fffffff) return 2;
    sizes[0] = b.N;
    sizes[1] = F;
    return 0;
}

int b200mesh_bcc_fill(uint64_t seed, double a, double gamma0, double psiOverDt, int32_t* lower,
                      int32_t* upper_addr, double* gamma_f, double* magSf, double* delta, double* diag0,
                      double* diag, double* upper, double* source, double* xstar, double* xyz) {
    if (!g_bcc) return 1;
    const Bcc& b = *g_bcc;
    const int64_t N = b.N;
    const double Shex = 3.0 * std::sqrt(3.0) / 16.0 * a * a, Ssq = a * a / 8.0;
    const double dhex = 1.0 / (std::sqrt(3.0) / 2.0 * a), dsq = 1.0 / a, V = 0.5 * a * a * a;
    std::vector<double> gam((size_t)N);
    std::vector<int64_t> fstart((size_t)N + 1, 0);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < N; ++c) {
        const int64_t site = b.siteOfCell[c];
        double p[3];
        site_xyz(b, site, a, p);
        xyz[3 * c] = p[0]; xyz[3 * c + 1] = p[1]; xyz[3 * c + 2] = p[2];
        const double xi = uni(seed, (uint64_t)site), xj = uni(seed ^ 0xABCDEF1234567ull, (uint64_t)site);
        gam[c] = gamma0 * (1.0 + 0.9 * std::sin(2 * kPi * p[0]) * std::sin(2 * kPi * p[1]) * std::sin(2 * kPi * p[2])) *
                 std::pow(10.0, 0.5 * xi);
        xstar[c] = std::sin(kPi * p[0]) * std::cos(2 * kPi * p[1]) * std::sin(3 * kPi * p[2]) + 0.5 * p[0] + 0.01 * xj;
        int64_t nb[14]; int hx[14];
        const int n = b.neighbours(site, nb, hx);
        int cnt = 0;
        for (int k = 0; k < n; ++k) if (b.cellOfSite[nb[k]] > c) ++cnt;
        fstart[c + 1] = cnt;
        const int j = (int)((site >> 1) / b.nx % b.ny);
        diag0[c] = psiOverDt * V;
        if (j == b.ny - 1) diag0[c] += 0.0;   // filled below once gamma is known
    }
    for (int64_t c = 0; c < N; ++c) fstart[c + 1] += fstart[c];
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < N; ++c) {
        const int64_t site = b.siteOfCell[c];
        const int j = (int)((site >> 1) / b.nx % b.ny);
        if (j == b.ny - 1) diag0[c] += gam[c] * Ssq * (2.0 * dsq);   // fixed-value top patch
        int64_t nb[14]; int hx[14];
        const int n = b.neighbours(site, nb, hx);
        std::pair<int32_t, int> up[16];
        int cnt = 0;
        for (int k = 0; k < n; ++k) {
            const int32_t o = b.cellOfSite[nb[k]];
            if (o > c) up[cnt++] = {o, hx[k]};
        }
        std::sort(up, up + cnt);
        int64_t f = fstart[c];
        for (int k = 0; k < cnt; ++k, ++f) {
            lower[f] = (int32_t)c;
            upper_addr[f] = up[k].first;
            gamma_f[f] = 0.5 * (gam[c] + gam[up[k].first]);
            magSf[f] = up[k].second ? Shex : Ssq;
            delta[f] = up[k].second ? dhex : dsq;
            upper[f] = -1.0 * (delta[f] * (gamma_f[f] * magSf[f]));
        }
    }
    const int64_t F = fstart[N];
    for (int64_t c = 0; c < N; ++c) diag[c] = 0.0;
    for (int64_t f = 0; f < F; ++f) { diag[lower[f]] -= upper[f]; diag[upper_addr[f]] -= upper[f]; }
    for (int64_t c = 0; c < N; ++c) { diag[c] += diag0[c]; source[c] = diag[c] * xstar[c]; }
    for (int64_t f = 0; f < F; ++f) {
        source[upper_addr[f]] += upper[f] * xstar[lower[f]];
        source[lower[f]] += upper[f] * xstar[upper_addr[f]];
    }
    return 0;
}

}  // extern "C"
