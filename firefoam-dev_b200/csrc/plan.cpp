// plan.cpp -- see plan.hpp.  Host-side, once per mesh.
#include "plan.hpp"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <numeric>

namespace b200 {

namespace {

// CSR over natural cells: for cell c the faces in OpenFOAM's own visiting order of
// lduMatrix::Amul (OF-dev lduMatrixATmul.C): first the faces where c is the
// neighbour (upperAddr == c; ascending face index == losort order), then the faces c owns.
struct CellFaces {
    std::vector<int64_t> start;   // [N+1]
    std::vector<int32_t> face;    // [2F]
    std::vector<int32_t> other;   // [2F]
    std::vector<int32_t> nLowerNat;  // [N] number of faces where c is the neighbour
};

void build_cell_faces(int32_t N, int32_t F, const int32_t* l, const int32_t* u, CellFaces& cf) {
    cf.start.assign((size_t)N + 1, 0);
    cf.nLowerNat.assign((size_t)N, 0);
    for (int32_t f = 0; f < F; ++f) {
        cf.start[(size_t)l[f] + 1]++;
        cf.start[(size_t)u[f] + 1]++;
        cf.nLowerNat[u[f]]++;
    }
    for (int32_t c = 0; c < N; ++c) cf.start[(size_t)c + 1] += cf.start[c];
    cf.face.resize((size_t)2 * F);
    cf.other.resize((size_t)2 * F);
    std::vector<int64_t> posLow(cf.start.begin(), cf.start.end() - 1);
    std::vector<int64_t> posUp((size_t)N);
    for (int32_t c = 0; c < N; ++c) posUp[c] = cf.start[c] + cf.nLowerNat[c];
    for (int32_t f = 0; f < F; ++f) {
        int64_t a = posUp[l[f]]++;
        cf.face[a] = f;
        cf.other[a] = u[f];
        int64_t b = posLow[u[f]]++;
        cf.face[b] = f;
        cf.other[b] = l[f];
    }
}

// Reverse Cuthill-McKee order of the cell-cell graph: order[k] = k-th cell.  Every connected
// component is started from a pseudo-peripheral cell (two BFS sweeps from its first unvisited
// cell), neighbours are visited in ascending degree.
void rcm_order(int32_t N, const CellFaces& cf, std::vector<int32_t>& order) {
    order.clear();
    order.reserve((size_t)N);
    std::vector<uint8_t> seen((size_t)N, 0);
    std::vector<int32_t> level, nbrs;
    auto deg = [&](int32_t c) { return (int32_t)(cf.start[c + 1] - cf.start[c]); };
    // BFS from s over unseen cells without marking them permanently; returns a far, low-degree cell
    std::vector<int32_t> queue, touched;
    std::vector<int32_t> dist((size_t)N, -1);
    auto far_cell = [&](int32_t s) {
        queue.clear();
        touched.clear();
        queue.push_back(s);
        dist[s] = 0;
        touched.push_back(s);
        size_t head = 0;
        while (head < queue.size()) {
            const int32_t c = queue[head++];
            for (int64_t e = cf.start[c]; e < cf.start[c + 1]; ++e) {
                const int32_t o = cf.other[e];
                if (!seen[o] && dist[o] < 0) {
                    dist[o] = dist[c] + 1;
                    queue.push_back(o);
                    touched.push_back(o);
                }
            }
        }
        const int32_t dmax = dist[queue.back()];
        int32_t best = queue.back();
        for (size_t i = queue.size(); i-- > 0 && dist[queue[i]] == dmax;)
            if (deg(queue[i]) < deg(best)) best = queue[i];
        for (int32_t c : touched) dist[c] = -1;
        return best;
    };
    for (int32_t s0 = 0; s0 < N; ++s0) {
        if (seen[s0]) continue;
        int32_t s = far_cell(s0);
        s = far_cell(s);
        size_t head = order.size();
        order.push_back(s);
        seen[s] = 1;
        while (head < order.size()) {
            const int32_t c = order[head++];
            nbrs.clear();
            for (int64_t e = cf.start[c]; e < cf.start[c + 1]; ++e) {
                const int32_t o = cf.other[e];
                if (!seen[o]) {
                    seen[o] = 1;
                    nbrs.push_back(o);
                }
            }
            std::sort(nbrs.begin(), nbrs.end(), [&](int32_t a, int32_t b) {
                const int32_t da = deg(a), db = deg(b);
                return da != db ? da < db : a < b;
            });
            for (int32_t o : nbrs) order.push_back(o);
        }
    }
    std::reverse(order.begin(), order.end());
}

// mean |pos(l) - pos(u)| over the faces (pos == nullptr: natural numbering)
double mean_span(int32_t F, const int32_t* l, const int32_t* u, const int32_t* pos) {
    if (F == 0) return 0.0;
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int32_t f = 0; f < F; ++f) {
        const int64_t a = pos ? pos[l[f]] : l[f], b = pos ? pos[u[f]] : u[f];
        s += (double)(a > b ? a - b : b - a);
    }
    return s / F;
}

// Coalescing metric of a row order: the kernels run one row per thread, a warp = 32 consecutive rows,
// and the j-th neighbour gathers of a warp form one memory request.  Returns the mean number of
// distinct 32-byte sectors (4 doubles) per such request over a sample of warps: ~8-9 on a
// lexicographic hex mesh, ~30 (of 32) on a locally shuffled numbering.
// order: position -> cell (nullptr: natural), pos: cell -> position (nullptr: natural).
double sectors_per_gather(int32_t N, const CellFaces& cf, const int32_t* order, const int32_t* pos) {
    const int32_t nSlices = (N + 31) / 32;
    if (nSlices == 0) return 0.0;
    const int32_t step = std::max(1, nSlices / 4096);
    double total = 0.0, requests = 0.0;
#pragma omp parallel for reduction(+ : total, requests) schedule(static)
    for (int32_t s = 0; s < nSlices; s += step) {
        int32_t sec[32];
        for (int32_t j = 0;; ++j) {
            int n = 0;
            for (int32_t k = s * 32; k < std::min(N, s * 32 + 32); ++k) {
                const int32_t c = order ? order[k] : k;
                if (cf.start[c] + j < cf.start[c + 1]) {
                    const int32_t o = cf.other[cf.start[c] + j];
                    sec[n++] = (pos ? pos[o] : o) >> 2;
                }
            }
            if (n == 0) break;
            std::sort(sec, sec + n);
            total += (double)(std::unique(sec, sec + n) - sec);
            requests += 1.0;
        }
    }
    return requests > 0 ? total / requests : 0.0;
}

// Greedy sequential multicolouring (first-fit), cells visited in `base` order (natural if empty).
int32_t colour_greedy(int32_t N, const CellFaces& cf, const std::vector<int32_t>& base,
                      std::vector<int32_t>& colour) {
    colour.assign((size_t)N, -1);
    int32_t nCol = 0;
    std::vector<int32_t> mark;  // mark[k] == c  <=> colour k is used by a neighbour of c
    for (int32_t k0 = 0; k0 < N; ++k0) {
        const int32_t c = base.empty() ? k0 : base[k0];
        for (int64_t e = cf.start[c]; e < cf.start[c + 1]; ++e) {
            int32_t k = colour[cf.other[e]];
            if (k >= 0) {
                if ((size_t)k >= mark.size()) mark.resize((size_t)k + 1, -1);
                mark[k] = c;
            }
        }
        int32_t k = 0;
        while ((size_t)k < mark.size() && mark[k] == c) ++k;
        colour[c] = k;
        if (k + 1 > nCol) nCol = k + 1;
        if ((size_t)k >= mark.size()) mark.resize((size_t)k + 1, -1);
    }
    return nCol;
}

// Dependency levels of the natural-order recurrences of DICPreconditioner (OF-dev
// DICPreconditioner.C): row u depends on every row l < u it shares a face
// with.  level(u) = 1 + max level(l).  Faces are sorted by l, and every face INTO l has a
// smaller owner, so one ascending pass over the faces is enough.
int32_t colour_levels(int32_t N, int32_t F, const int32_t* l, const int32_t* u,
                      std::vector<int32_t>& colour) {
    colour.assign((size_t)N, 0);
    int32_t nLev = N > 0 ? 1 : 0;
    for (int32_t f = 0; f < F; ++f) {
        int32_t cand = colour[l[f]] + 1;
        if (cand > colour[u[f]]) {
            colour[u[f]] = cand;
            if (cand + 1 > nLev) nLev = cand + 1;
        }
    }
    return nLev;
}

}  // namespace

std::string build_plan(Ordering ordering, int32_t N, int32_t F, const int32_t* l, const int32_t* u,
                       int32_t nIfaces, const IfaceIn* ifaces, HostPlan& P, Renumber renumber, int32_t tileRows,
                       bool sortColumns) {
    if (N < 0 || F < 0 || nIfaces < 0) return "negative size";
    if (F > 0 && (!l || !u)) return "null lowerAddr/upperAddr";
    if (nIfaces > 0 && !ifaces) return "null interface list";
    for (int32_t f = 0; f < F; ++f) {
        if (l[f] < 0 || u[f] >= N || l[f] >= u[f])
            return "face " + std::to_string(f) + ": need 0 <= lowerAddr < upperAddr < nCells";
        if (f > 0 && (l[f] < l[f - 1]))
            return "face " + std::to_string(f) + ": lowerAddr not in upper-triangular (ascending) order";
    }
    for (int32_t k = 0; k < nIfaces; ++k) {
        if (ifaces[k].nFaces < 0 || (ifaces[k].nFaces > 0 && !ifaces[k].faceCells))
            return "interface " + std::to_string(k) + ": bad size or null faceCells";
        for (int32_t i = 0; i < ifaces[k].nFaces; ++i)
            if (ifaces[k].faceCells[i] < 0 || ifaces[k].faceCells[i] >= N)
                return "interface " + std::to_string(k) + ": faceCells out of range";
    }

    P = HostPlan();
    P.ordering = ordering;
    P.N = N;
    P.F = F;

    CellFaces cf;
    build_cell_faces(N, F, l, u, cf);

    // ---- baseOrder order: natural, or reverse Cuthill-McKee when the given numbering is cache-hostile ----
    // (never for Levels: DIC-exact reproduces the recurrences of the NATURAL face order)
    std::vector<int32_t> baseOrder;   // baseOrder position -> cell (empty == natural)
    P.spanNatural = P.spanUsed = mean_span(F, l, u, nullptr);
    if (ordering != Ordering::Levels && renumber != Renumber::Off && N > 1 && F > 0) {
        // Two symptoms of a cache-hostile numbering: a mean face span far above a mesh plane
        // (~N^(2/3): long-range jumps, L2 misses), or warps whose gathers do not coalesce (local
        // shuffles: L1 sector-throughput bound).  RCM is kept only if it clearly improves either.
        const double banded = 4.0 * std::pow((double)N, 2.0 / 3.0) + 64.0;
        P.sectorsNatural = P.sectorsUsed = sectors_per_gather(N, cf, nullptr, nullptr);
        if (renumber == Renumber::Force || P.spanNatural > banded || P.sectorsNatural > 12.0) {
            rcm_order(N, cf, baseOrder);
            std::vector<int32_t> pos((size_t)N);
            for (int32_t k = 0; k < N; ++k) pos[baseOrder[k]] = k;
            const double spanRcm = mean_span(F, l, u, pos.data());
            const double secRcm = sectors_per_gather(N, cf, baseOrder.data(), pos.data());
            const bool better = (P.spanNatural > banded && spanRcm < 0.5 * P.spanNatural && secRcm < 1.1 * P.sectorsNatural) ||
                                (secRcm < 0.85 * P.sectorsNatural && spanRcm < 2.0 * std::max(P.spanNatural, banded));
            if (renumber == Renumber::Force || better) {
                P.renumbered = true;
                P.spanUsed = spanRcm;
                P.sectorsUsed = secRcm;
            } else {
                baseOrder.clear();
            }
        }
    }

    // ---- row order ----------------------------------------------------------------------
    if (ordering == Ordering::Natural) {
        P.nColours = 1;
        P.colourStart = {0, N};
        if (P.renumbered) {
            P.perm = baseOrder;
            P.iperm.resize((size_t)N);
            for (int32_t k = 0; k < N; ++k) P.iperm[baseOrder[k]] = k;
        }
    } else {
        std::vector<int32_t> colour;
        P.nColours = (ordering == Ordering::MultiColour) ? colour_greedy(N, cf, baseOrder, colour)
                                                        : colour_levels(N, F, l, u, colour);
        if (N == 0) P.nColours = 1;
        P.colourStart.assign((size_t)P.nColours + 1, 0);
        for (int32_t c = 0; c < N; ++c) P.colourStart[(size_t)colour[c] + 1]++;
        for (int32_t k = 0; k < P.nColours; ++k) P.colourStart[k + 1] += P.colourStart[k];
        P.perm.resize((size_t)N);
        P.iperm.resize((size_t)N);
        // tiles of `tileRows` consecutive base-order cells (multicolour only, power of two, >= 2 tiles)
        const int32_t C = P.nColours;
        int32_t tr = 0;
        if (ordering == Ordering::MultiColour && tileRows > 0 && N > 2 * (int64_t)tileRows) {
            tr = 32;
            while (tr < tileRows) tr <<= 1;
        }
        P.tileRows = tr;
        P.nTiles = tr ? (int32_t)(((int64_t)N + tr - 1) / tr) : 1;
        P.segStart.assign((size_t)P.nTiles * C + 1, 0);
        auto tileOf = [&](int32_t k) { return tr ? k / tr : 0; };
        for (int32_t k = 0; k < N; ++k) {
            const int32_t c = baseOrder.empty() ? k : baseOrder[k];
            P.segStart[(size_t)tileOf(k) * C + colour[c] + 1]++;
        }
        for (size_t i = 0; i + 1 < P.segStart.size(); ++i) P.segStart[i + 1] += P.segStart[i];
        std::vector<int32_t> pos(P.segStart.begin(), P.segStart.end() - 1);
        P.rowColour.resize((size_t)N);
        for (int32_t k = 0; k < N; ++k) {  // stable: base (natural / RCM) order inside a (tile, colour) segment
            const int32_t c = baseOrder.empty() ? k : baseOrder[k];
            int32_t r = pos[(size_t)tileOf(k) * C + colour[c]]++;
            P.perm[r] = c;
            P.iperm[c] = r;
            P.rowColour[r] = colour[c];
        }
    }
    const bool ident = P.perm.empty();
    auto rowOf = [&](int32_t c) { return ident ? c : P.iperm[c]; };
    auto cellOf = [&](int32_t r) { return ident ? r : P.perm[r]; };

    // ---- sliced ELL ---------------------------------------------------------------------
    P.nSlices = (N + 31) / 32;
    P.sliceBase.assign((size_t)P.nSlices + 1, 0);
    P.rowLen.assign((size_t)N, 0);
    for (int32_t s = 0; s < P.nSlices; ++s) {
        int64_t mx = 0;
        for (int32_t r = s * 32; r < std::min(N, s * 32 + 32); ++r) {
            int32_t c = cellOf(r);
            int64_t len = cf.start[c + 1] - cf.start[c];
            if (len > 65535) return "row with more than 65535 faces is unsupported";
            mx = std::max(mx, len);
        }
        P.sliceBase[s + 1] = P.sliceBase[s] + mx * 32;
    }
    P.nEntries = P.sliceBase[P.nSlices];
    if (P.nEntries > (int64_t)0x7fffffff0LL) return "matrix too large";
    P.col.resize((size_t)P.nEntries);
    P.faceOf.assign((size_t)P.nEntries, -1);
#pragma omp parallel
    {
        std::vector<std::pair<int32_t, int32_t>> ent;  // (face, other row)
#pragma omp for schedule(static)
        for (int32_t r = 0; r < N; ++r) {
            const int32_t c = cellOf(r);
            const int64_t base = P.sliceBase[r / 32] + (r % 32);
            const int64_t width = (P.sliceBase[r / 32 + 1] - P.sliceBase[r / 32]) / 32;
            int32_t j = 0, nLower = 0;
            // cf lists the faces where c is the neighbour (ascending), then the faces c owns
            // (ascending).  In natural order that already is [earlier | later], each ascending;
            // in a permuted order re-sort the few entries by face and split by row index.
            const int64_t e0 = cf.start[c], e1 = cf.start[c + 1];
            ent.clear();
            for (int64_t e = e0; e < e1; ++e) ent.emplace_back(cf.face[e], rowOf(cf.other[e]));
            if (!ident) std::sort(ent.begin(), ent.end());
            // DIC-class sweeps reproduce no OpenFOAM summation order: by column, so that the j-th gathers of
            // neighbouring rows are neighbours in memory (plan.hpp sortColumns)
            if (sortColumns && ordering == Ordering::MultiColour)
                std::sort(ent.begin(), ent.end(), [](const std::pair<int32_t, int32_t>& a,
                                                     const std::pair<int32_t, int32_t>& b) {
                    return a.second != b.second ? a.second < b.second : a.first < b.first;
                });
            // A renumbered Natural plan serves Amul / sumA / negSumDiag only: keep the whole row in
            // ascending face order (OpenFOAM's visiting order) -- no [earlier | later] grouping,
            // which would change the order of the row sum.
            const bool grouped = !(ordering == Ordering::Natural && !ident);
            // "earlier" = eliminated earlier: smaller row index, or -- in a (tiled) multicolour order,
            // where storage order and elimination order differ -- an earlier colour
            const bool byColour = !P.rowColour.empty();
            auto earlier = [&](int32_t other) {
                return byColour ? P.rowColour[other] < P.rowColour[r] : other < r;
            };
            for (auto& fe : ent)
                if (grouped && earlier(fe.second)) {
                    P.col[base + 32 * (int64_t)j] = fe.second;
                    P.faceOf[base + 32 * (int64_t)j] = fe.first;
                    ++j;
                    ++nLower;
                }
            for (auto& fe : ent)
                if (!grouped || !earlier(fe.second)) {
                    P.col[base + 32 * (int64_t)j] = fe.second;
                    P.faceOf[base + 32 * (int64_t)j] = fe.first;
                    ++j;
                }
            P.rowLen[r] = (uint32_t)nLower | ((uint32_t)j << 16);
            for (int64_t jj = j; jj < width; ++jj) P.col[base + 32 * jj] = r;
        }
    }
    // padding rows of the last slice
    for (int64_t r = N; r < (int64_t)P.nSlices * 32; ++r) {
        const int64_t base = P.sliceBase[r / 32] + (r % 32);
        const int64_t width = (P.sliceBase[r / 32 + 1] - P.sliceBase[r / 32]) / 32;
        for (int64_t jj = 0; jj < width; ++jj) P.col[base + 32 * jj] = 0;
    }

    // ---- 16-bit column offsets per (slice, entry) ---------------------------------------------
    if (N > 0 && P.nEntries > 0) {
        P.colBase.assign((size_t)(P.nEntries / 32), -1);
        P.col16.assign((size_t)P.nEntries, 0);
        int64_t fit = 0, total = 0;
#pragma omp parallel for schedule(static) reduction(+ : fit, total)
        for (int32_t sl = 0; sl < P.nSlices; ++sl) {
            const int64_t b0 = P.sliceBase[sl];
            const int64_t width = (P.sliceBase[sl + 1] - b0) / 32;
            const int32_t r0 = sl * 32, r1 = std::min(N, sl * 32 + 32);
            for (int64_t j = 0; j < width; ++j) {
                int32_t lo = INT32_MAX, hi = -1;
                for (int32_t r = r0; r < r1; ++r)
                    if ((int64_t)(P.rowLen[r] >> 16) > j) {
                        const int32_t c = P.col[b0 + 32 * j + (r - r0)];
                        lo = std::min(lo, c);
                        hi = std::max(hi, c);
                    }
                if (hi < 0) continue;          // no row of the slice has a j-th entry
                ++total;
                if ((int64_t)hi - lo < 65536) {
                    ++fit;
                    P.colBase[(size_t)(b0 / 32 + j)] = lo;
                    for (int32_t r = r0; r < r1; ++r)
                        if ((int64_t)(P.rowLen[r] >> 16) > j)
                            P.col16[b0 + 32 * j + (r - r0)] = (uint16_t)(P.col[b0 + 32 * j + (r - r0)] - lo);
                }
            }
        }
        P.col16Fraction = total ? (double)fit / (double)total : 0.0;
        if (fit == 0) {
            std::vector<int32_t>().swap(P.colBase);
            std::vector<uint16_t>().swap(P.col16);
        }
    }

    // ---- symmetric single-read layout -----------------------------------------------------
    // Built from the cell-face lists, split by ROW INDEX (lower = smaller index: that row's thread
    // streamed the value moments earlier), whatever grouping the full-row ELL above uses.  Each group
    // in ascending face order: in the natural order that is OpenFOAM's visiting order of the whole row.
    // Not for level orders (neighbours far upstream) nor renumbered natural plans (the row sum must
    // stay in pure face order there, which a [lower | upper] split does not give: those get the
    // face-ordered single-read layout, SrPlan, below).
    if (N > 0 && N <= (1 << 27) && ordering != Ordering::Levels && !(ordering == Ordering::Natural && !ident)) {
        SymPlan& Sp = P.sym;
        Sp.rowLen.assign((size_t)N, 0u);
        int64_t wU = 0, wL = 0;
#pragma omp parallel for schedule(static) reduction(max : wU, wL)
        for (int32_t r = 0; r < N; ++r) {
            const int32_t c = cellOf(r);
            int32_t nLo = 0, nT = 0;
            for (int64_t e = cf.start[c]; e < cf.start[c + 1]; ++e, ++nT)
                if (rowOf(cf.other[e]) < r) ++nLo;
            Sp.rowLen[r] = (uint32_t)nLo | ((uint32_t)nT << 16);
            wU = std::max<int64_t>(wU, nT - nLo);
            wL = std::max<int64_t>(wL, nLo);
        }
        const int64_t nU = (int64_t)P.nSlices * 32 * wU, nL = (int64_t)P.nSlices * 32 * wL;
        if (wU <= 32 && nU < 0x7fffffffLL && nL < 0x7fffffffLL) {
            Sp.WU = (int32_t)wU;
            Sp.WL = (int32_t)wL;
            Sp.nU = nU;
            Sp.nL = nL;
            Sp.uCol.assign((size_t)nU, 0);
            Sp.uFace.assign((size_t)nU, -1);
            Sp.lRef.assign((size_t)nL, 0);
#pragma omp parallel
            {
                std::vector<std::pair<int32_t, int32_t>> ent;  // (face, other row)
                // pass 1: upper entries of every row
#pragma omp for schedule(static)
                for (int32_t r = 0; r < N; ++r) {
                    const int32_t c = cellOf(r);
                    ent.clear();
                    for (int64_t e = cf.start[c]; e < cf.start[c + 1]; ++e)
                        ent.emplace_back(cf.face[e], rowOf(cf.other[e]));
                    if (!ident) std::sort(ent.begin(), ent.end());
                    const int64_t ub = (int64_t)(r / 32) * 32 * wU + (r % 32);
                    int64_t j = 0;
                    for (auto& fe : ent) {
                        if (fe.second > r) {
                            Sp.uCol[ub + 32 * j] = fe.second;
                            Sp.uFace[ub + 32 * j] = fe.first;
                            ++j;
                        }
                    }
                    for (; j < wU; ++j) Sp.uCol[ub + 32 * j] = r;
                }
                // pass 2: lower entries as references (owner row a << 5 | q-th upper entry of a)
#pragma omp for schedule(static)
                for (int32_t r = 0; r < N; ++r) {
                    const int32_t c = cellOf(r);
                    ent.clear();
                    for (int64_t e = cf.start[c]; e < cf.start[c + 1]; ++e)
                        ent.emplace_back(cf.face[e], rowOf(cf.other[e]));
                    if (!ident) std::sort(ent.begin(), ent.end());
                    const int64_t lb = (int64_t)(r / 32) * 32 * wL + (r % 32);
                    int64_t j = 0;
                    for (auto& fe : ent) {
                        if (fe.second < r) {
                            const int32_t a = fe.second;
                            const int64_t ab = (int64_t)(a / 32) * 32 * wU + (a % 32);
                            const int32_t aU = (int32_t)(Sp.rowLen[a] >> 16) - (int32_t)(Sp.rowLen[a] & 0xffffu);
                            int32_t q = 0;
                            for (int32_t k = 0; k < aU; ++k)
                                if (Sp.uFace[ab + 32 * (int64_t)k] == fe.first) { q = k; break; }
                            Sp.lRef[lb + 32 * j] = ((uint32_t)a << 5) | (uint32_t)q;
                            ++j;
                        }
                    }
                }
            }
            Sp.valid = true;
        } else {
            P.sym = SymPlan();
        }
    }

    // ---- single-read face-ordered layout (renumbered natural plans; plan.hpp SrPlan) -----------
    if (ordering == Ordering::Natural && !ident && N > 0 && N <= (1 << 27)) {
        SrPlan& R = P.sr;
        // own count of every row (neighbour has the larger row index) and the per-slice own widths
        std::vector<int32_t> nOwnRow((size_t)N, 0);
        bool fits = true;
#pragma omp parallel for schedule(static) reduction(&& : fits)
        for (int32_t r = 0; r < N; ++r) {
            const int32_t c = cellOf(r);
            int32_t k = 0;
            for (int64_t e = cf.start[c]; e < cf.start[c + 1]; ++e)
                if (rowOf(cf.other[e]) > r) ++k;
            nOwnRow[r] = k;
            fits = fits && k <= 31;          // q < 31 addresses the own values 0..30; an own count of 31 is still fine
        }
        if (fits) {
            R.ownBase.assign((size_t)P.nSlices + 1, 0);
            for (int32_t sl = 0; sl < P.nSlices; ++sl) {
                int64_t mx = 0;
                for (int32_t r = sl * 32; r < std::min(N, sl * 32 + 32); ++r) mx = std::max<int64_t>(mx, nOwnRow[r]);
                R.ownBase[sl + 1] = R.ownBase[sl] + 32 * mx;
            }
            R.nOwn = R.ownBase[P.nSlices];
            R.ownFace.assign((size_t)R.nOwn, -1);
            R.meta.assign((size_t)P.nEntries, 0u);
            // pass 1: own slots in ascending face order (the full-row ELL of a renumbered natural plan is in
            // ascending face order already: its entries ARE the row's visiting order)
#pragma omp parallel for schedule(static)
            for (int32_t r = 0; r < N; ++r) {
                const int64_t base = P.sliceBase[r / 32] + (r % 32);
                const int64_t ob = R.ownBase[r / 32] + (r % 32);
                const int32_t n = (int32_t)(P.rowLen[r] >> 16);
                int64_t j = 0;
                for (int32_t k = 0; k < n; ++k) {
                    const int64_t e = base + 32 * (int64_t)k;
                    if (P.col[e] > r) R.ownFace[ob + 32 * j++] = P.faceOf[e];
                }
            }
            // pass 2: the meta words; a reference finds its face among the owner's own slots
            bool ok = true;
#pragma omp parallel for schedule(static) reduction(&& : ok)
            for (int32_t r = 0; r < N; ++r) {
                const int64_t base = P.sliceBase[r / 32] + (r % 32);
                const int32_t n = (int32_t)(P.rowLen[r] >> 16);
                for (int32_t k = 0; k < n; ++k) {
                    const int64_t e = base + 32 * (int64_t)k;
                    const int32_t a = P.col[e];
                    uint32_t q = 31u;
                    if (a < r) {
                        const int64_t ab = R.ownBase[a / 32] + (a % 32);
                        int32_t found = -1;
                        for (int32_t t = 0; t < nOwnRow[a]; ++t)
                            if (R.ownFace[ab + 32 * (int64_t)t] == P.faceOf[e]) { found = t; break; }
                        if (found < 0 || found > 30) { ok = false; found = 0; }
                        q = (uint32_t)found;
                    }
                    R.meta[e] = ((uint32_t)a << 5) | q;
                }
            }
            R.valid = ok;
            if (!ok) P.sr = SrPlan();
        }
    }

    // ---- interfaces ---------------------------------------------------------------------
    P.nIfaces = nIfaces;
    P.nbrRank.resize((size_t)nIfaces);
    P.patchStart.assign((size_t)nIfaces + 1, 0);
    for (int32_t k = 0; k < nIfaces; ++k) {
        P.nbrRank[k] = ifaces[k].nbrRank;
        P.patchStart[k + 1] = P.patchStart[k] + ifaces[k].nFaces;
    }
    const int32_t nSlots = P.patchStart[nIfaces];
    P.slotRow.resize((size_t)nSlots);
    for (int32_t k = 0; k < nIfaces; ++k)
        for (int32_t i = 0; i < ifaces[k].nFaces; ++i)
            P.slotRow[P.patchStart[k] + i] = rowOf(ifaces[k].faceCells[i]);
    // CSR row -> slots; stable sort keeps (patch, face) order inside a row, which is the
    // order of OpenFOAM's updateMatrixInterfaces loop over patches
    std::vector<int32_t> order((size_t)nSlots);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](int32_t a, int32_t b) { return P.slotRow[a] < P.slotRow[b]; });
    P.bSlot = order;
    P.bRow.clear();
    P.bStart.clear();
    for (int32_t i = 0; i < nSlots; ++i) {
        int32_t r = P.slotRow[order[i]];
        if (P.bRow.empty() || P.bRow.back() != r) {
            P.bRow.push_back(r);
            P.bStart.push_back(i);
        }
    }
    P.bStart.push_back(nSlots);
    P.nBRows = (int32_t)P.bRow.size();
    return std::string();
}

}  // namespace b200
