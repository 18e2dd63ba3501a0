/*
 * bicg_oracle.c -- TEST INFRASTRUCTURE.  CPU restatement of `PBiCG` + `DILU` on asymmetric lduMatrices, the other
 * solver the reference's cases select for their transport equations.
 *
 * Only tests/ and __graft_entry__.smoke() may load this; the product (libb200pcg.so) never does.
 *
 * What it restates.  The reference (LeiXu84/fireFoam-dev 17.11.10) selects
 *   cases/wallFireSpread2D/system/fvSolution:66-73   Yi { solver PBiCG; preconditioner DILU; tolerance 1e-8; relTol 0; }
 *   cases/pyrolysis1D/system/fvSolution, the panelRegion dictionaries of both, cases/singleBox/system/pyrolysisRegion/fvSolution
 * and its two older golden logs of the steckler case were produced with it:
 *   cases/steckler/original/log.fireFoam and original/darwinIntel64/log.fireFoam: 207 `DILUPBiCG:` lines each
 *   (:157-161 are the first five).
 * The arithmetic lives in the un-vendored OpenFOAM (dev @ 940e28f6 for the current build, CHANGELOG:1-3; 2.4.x for the
 * older logs), absent from /root/reference and from this image.  The functions below follow the published algorithm
 * of the upstream files:
 *   PBiCG.C              PBiCG::solve               bi-conjugate gradients with left preconditioning, literal control flow
 *   DILUPreconditioner.C calcReciprocalD / precondition / preconditionT   (forward loop in losort order, see below)
 *   diagonalPreconditioner.C, noPreconditioner.C
 *   lduMatrixATmul.C     lduMatrix::Amul / Tmul / sumA with lower != upper;  lduMatrixSolver.C normFactor
 *
 * PARITY UNPINNED by the reference: the DILUPBiCG lines of its logs need matrices that only the whole solver can
 * assemble.  Two anchors exist and are tested (tests/test_bicg_oracle.py): the line that needs no matrix -- zero field,
 * zero source: `Initial residual = 0, Final residual = 0, No Iterations 0` (original/darwinIntel64/log.fireFoam:161)
 * -- and, on a SYMMETRIC matrix, DILU == DIC and BiCG == CG: on the digit-pinned steckler ph_rgh system this file
 * reproduces the DICPCG lines of pcg_oracle.c (29 iterations, log.fireFoam:92).  Everything else is checked against an
 * independent dense formulation.
 *
 * Plain C, double precision, int32 labels, no FMA contraction (-ffp-contract=off).  One rank.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct bi_sys {
    int32_t nCells, nFaces;
    const int32_t* lower;       /* lowerAddr (owner)     */
    const int32_t* upper;       /* upperAddr (neighbour) */
    const double* diag;
    const double* upperCoeffs;  /* A[l][u] */
    const double* lowerCoeffs;  /* A[u][l]; == upperCoeffs for a symmetric matrix */
    const double* source;
    double* psi;                /* in: x0, out: x */
} bi_sys;

typedef struct bi_controls {
    double tolerance, relTol;
    int32_t maxIter, minIter;
    int32_t precond;            /* 0 none, 1 diagonal, 2 DILU */
    int32_t pad;
} bi_controls;

typedef struct bi_perf {
    double initialResidual, finalResidual, normFactor;
    int32_t nIterations, converged, singular, pad;
} bi_perf;

/* lduMatrix::Amul / Tmul (OF-dev lduMatrixATmul.C) */
static void bi_amul(const bi_sys* m, double* y, const double* x) {
    for (int32_t c = 0; c < m->nCells; ++c) y[c] = m->diag[c] * x[c];
    for (int32_t f = 0; f < m->nFaces; ++f) {
        y[m->upper[f]] += m->lowerCoeffs[f] * x[m->lower[f]];
        y[m->lower[f]] += m->upperCoeffs[f] * x[m->upper[f]];
    }
}
static void bi_tmul(const bi_sys* m, double* y, const double* x) {
    for (int32_t c = 0; c < m->nCells; ++c) y[c] = m->diag[c] * x[c];
    for (int32_t f = 0; f < m->nFaces; ++f) {
        y[m->upper[f]] += m->upperCoeffs[f] * x[m->lower[f]];
        y[m->lower[f]] += m->lowerCoeffs[f] * x[m->upper[f]];
    }
}
static void bi_sumA(const bi_sys* m, double* s) {
    for (int32_t c = 0; c < m->nCells; ++c) s[c] = m->diag[c];
    for (int32_t f = 0; f < m->nFaces; ++f) {
        s[m->upper[f]] += m->lowerCoeffs[f];
        s[m->lower[f]] += m->upperCoeffs[f];
    }
}

/* lduAddressing::losortAddr (OF-dev lduAddressing.C): faces ordered by upperAddr, ascending face within a cell */
static int32_t* bi_losort(const bi_sys* m) {
    const int32_t N = m->nCells, F = m->nFaces;
    int32_t* start = (int32_t*)calloc((size_t)N + 2, sizeof(int32_t));
    int32_t* lo = (int32_t*)malloc((size_t)(F > 0 ? F : 1) * sizeof(int32_t));
    for (int32_t f = 0; f < F; ++f) start[m->upper[f] + 1]++;
    for (int32_t c = 0; c < N; ++c) start[c + 1] += start[c];
    for (int32_t f = 0; f < F; ++f) lo[start[m->upper[f]]++] = f;
    free(start);
    return lo;
}

/* DILUPreconditioner::calcReciprocalD (OF-dev DILUPreconditioner.C) */
static void dilu_calc_rd(const bi_sys* m, double* rD) {
    for (int32_t c = 0; c < m->nCells; ++c) rD[c] = m->diag[c];
    for (int32_t f = 0; f < m->nFaces; ++f)
        rD[m->upper[f]] -= m->upperCoeffs[f] * m->lowerCoeffs[f] / rD[m->lower[f]];
    for (int32_t c = 0; c < m->nCells; ++c) rD[c] = 1.0 / rD[c];
}
/* DILUPreconditioner::precondition: forward loop in losort order, reverse loop in face order */
static void dilu_precondition(const bi_sys* m, const int32_t* losort, const double* rD, double* w, const double* r) {
    const int32_t F = m->nFaces;
    for (int32_t c = 0; c < m->nCells; ++c) w[c] = rD[c] * r[c];
    for (int32_t i = 0; i < F; ++i) {
        const int32_t s = losort[i];
        w[m->upper[s]] -= rD[m->upper[s]] * m->lowerCoeffs[s] * w[m->lower[s]];
    }
    for (int32_t f = F - 1; f >= 0; --f)
        w[m->lower[f]] -= rD[m->lower[f]] * m->upperCoeffs[f] * w[m->upper[f]];
}
/* DILUPreconditioner::preconditionT: forward loop in face order, reverse loop in reverse losort order */
static void dilu_preconditionT(const bi_sys* m, const int32_t* losort, const double* rD, double* w, const double* r) {
    const int32_t F = m->nFaces;
    for (int32_t c = 0; c < m->nCells; ++c) w[c] = rD[c] * r[c];
    for (int32_t f = 0; f < F; ++f)
        w[m->upper[f]] -= rD[m->upper[f]] * m->upperCoeffs[f] * w[m->lower[f]];
    for (int32_t i = F - 1; i >= 0; --i) {
        const int32_t s = losort[i];
        w[m->lower[s]] -= rD[m->lower[s]] * m->lowerCoeffs[s] * w[m->upper[s]];
    }
}

static int bi_converged(const bi_perf* p, double tol, double relTol) {
    return (p->finalResidual < tol) || (relTol > 1e-20 && p->finalResidual < relTol * p->initialResidual);
}

/* PBiCG::solve (OF-dev PBiCG.C) -- literal control flow */
int orc_pbicg_solve(const bi_sys* m, const bi_controls* ctl, bi_perf* out) {
    const int32_t N = m->nCells;
    const size_t nb = (size_t)(N > 0 ? N : 1) * sizeof(double);
    double *pA = (double*)malloc(nb), *pT = (double*)calloc((size_t)(N > 0 ? N : 1), sizeof(double));
    double *wA = (double*)malloc(nb), *wT = (double*)malloc(nb), *rA = (double*)malloc(nb), *rT = (double*)malloc(nb);
    double* rD = NULL;
    int32_t* losort = NULL;
    double* psi = m->psi;
    bi_perf perf;
    memset(&perf, 0, sizeof(perf));
    double wArT = 1e20, wArTold = wArT;     /* great_ */

    bi_amul(m, wA, psi);
    bi_tmul(m, wT, psi);
    for (int32_t c = 0; c < N; ++c) { rA[c] = m->source[c] - wA[c]; rT[c] = m->source[c] - wT[c]; }
    {   /* normFactor(psi, source, wA, pA) */
        bi_sumA(m, pA);
        double sPsi = 0.0;
        for (int32_t c = 0; c < N; ++c) sPsi += psi[c];
        const double xRef = sPsi / (double)N;       /* gAverage */
        for (int32_t c = 0; c < N; ++c) pA[c] *= xRef;
        double nf = 0.0;
        for (int32_t c = 0; c < N; ++c) nf += fabs(wA[c] - pA[c]) + fabs(m->source[c] - pA[c]);
        perf.normFactor = nf + 1e-20;
    }
    {
        double s = 0.0;
        for (int32_t c = 0; c < N; ++c) s += fabs(rA[c]);
        perf.initialResidual = s / perf.normFactor;
        perf.finalResidual = perf.initialResidual;
    }
    if (ctl->minIter > 0 || !bi_converged(&perf, ctl->tolerance, ctl->relTol)) {
        if (ctl->precond != 0) {
            rD = (double*)malloc(nb);
            if (ctl->precond == 1) for (int32_t c = 0; c < N; ++c) rD[c] = 1.0 / m->diag[c];
            else { dilu_calc_rd(m, rD); losort = bi_losort(m); }
        }
        do {
            wArTold = wArT;
            if (ctl->precond == 0) { memcpy(wA, rA, (size_t)N * sizeof(double)); memcpy(wT, rT, (size_t)N * sizeof(double)); }
            else if (ctl->precond == 1) for (int32_t c = 0; c < N; ++c) { wA[c] = rD[c] * rA[c]; wT[c] = rD[c] * rT[c]; }
            else { dilu_precondition(m, losort, rD, wA, rA); dilu_preconditionT(m, losort, rD, wT, rT); }
            {
                double s = 0.0;
                for (int32_t c = 0; c < N; ++c) s += wA[c] * rT[c];
                wArT = s;
            }
            if (perf.nIterations == 0) {
                for (int32_t c = 0; c < N; ++c) { pA[c] = wA[c]; pT[c] = wT[c]; }
            } else {
                const double beta = wArT / wArTold;
                for (int32_t c = 0; c < N; ++c) { pA[c] = wA[c] + beta * pA[c]; pT[c] = wT[c] + beta * pT[c]; }
            }
            bi_amul(m, wA, pA);
            bi_tmul(m, wT, pT);
            double wApT = 0.0;
            for (int32_t c = 0; c < N; ++c) wApT += wA[c] * pT[c];
            if (!(fabs(wApT) / perf.normFactor > 1e-300)) {     /* checkSingularity: <= vSmall_ */
                perf.singular = 1;
                break;
            }
            const double alpha = wArT / wApT;
            for (int32_t c = 0; c < N; ++c) {
                psi[c] += alpha * pA[c];
                rA[c] -= alpha * wA[c];
                rT[c] -= alpha * wT[c];
            }
            {
                double s = 0.0;
                for (int32_t c = 0; c < N; ++c) s += fabs(rA[c]);
                perf.finalResidual = s / perf.normFactor;
            }
        } while ((perf.nIterations++ < ctl->maxIter && !bi_converged(&perf, ctl->tolerance, ctl->relTol)) ||
                 perf.nIterations < ctl->minIter);
    }
    perf.converged = bi_converged(&perf, ctl->tolerance, ctl->relTol);
    *out = perf;
    free(pA); free(pT); free(wA); free(wT); free(rA); free(rT); free(rD); free(losort);
    return 0;
}

/* stand-alone pieces for unit tests */
void orc_asym_tmul(const bi_sys* m, const double* x, double* y) { bi_tmul(m, y, x); }
void orc_dilu(const bi_sys* m, const double* r, double* rD, double* w, double* wT) {
    int32_t* lo = bi_losort(m);
    dilu_calc_rd(m, rD);
    dilu_precondition(m, lo, rD, w, r);
    dilu_preconditionT(m, lo, rD, wT, r);
    free(lo);
}
