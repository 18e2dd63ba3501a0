/*
 * smooth_oracle.c -- TEST INFRASTRUCTURE.  CPU restatement of the reference's OTHER linear solves of a
 * time step: `smoothSolver` + `symGaussSeidel` / `GaussSeidel` on ASYMMETRIC lduMatrices.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the product
 * (libb200pcg.so) never does.
 *
 * What it restates.  The reference (LeiXu84/fireFoam-dev 17.11.10) solves U, Yi, h and k with
 *   cases/steckler/system/fvSolution:48-61   solver smoothSolver; smoother symGaussSeidel; tolerance 1e-6 | 1e-8;
 *                                            relTol 0; maxIter 10   (nSweeps defaults to 1)
 * from solver/UEqn.H:19-30, solver/YEEqn.H:60,111 and the thermo / turbulence libraries, 261 times in the golden
 * log cases/steckler/original/linux64/log.fireFoam (:172-178 are the first seven).  The arithmetic lives in the
 * un-vendored dependency OpenFOAM-dev @ 940e28f63681c7e5b292096d8fd35a71acd52599 (CHANGELOG:1-3,
 * solver/Make/options:47), absent from /root/reference and from this image.  The functions below follow the
 * published algorithm of the named upstream files:
 *   smoothSolver.C                smoothSolver::solve          control flow (nSweeps < 0: fixed sweeps)
 *   GaussSeidelSmoother.C         GaussSeidelSmoother::smooth  forward sweep, coupled patches as explicit (Jacobi)
 *   symGaussSeidelSmoother.C      symGaussSeidelSmoother::smooth  forward then reverse sweep
 *   lduMatrixATmul.C              lduMatrix::Amul / sumA / residual with lower != upper
 *   lduMatrixSolver.C             lduMatrix::solver::normFactor
 *
 * PARITY UNPINNED.  The reference's log pins these solves only by residual lines whose matrices need the whole
 * solver (UEqn / YEEqn / EEqn assembly, turbulence, combustion, radiation) and cannot be rebuilt here.  The one
 * line that depends on nothing else IS reproduced (tests/test_smooth_oracle.py): a zero field with a zero source
 * prints `Initial residual = 0, Final residual = 0, No Iterations 0` (log.fireFoam:176,178 -- H2O, CO2).  All other
 * checks of this file are against an independent formulation (dense triangular solves), not against the reference.
 *
 * Plain C, double precision, int32 labels, no FMA contraction (-ffp-contract=off).  Multi-rank runs are
 * emulated with one pthread per rank like pcg_oracle.c; reductions in ascending rank order.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct sm_iface {
    int32_t nbrRank;
    int32_t nFaces;
    const int32_t* faceCells;
    const double* bouCoeffs;   /* interfaceBouCoeffs[k] */
} sm_iface;

typedef struct sm_rank {
    int32_t nCells, nFaces;
    const int32_t* lower;      /* lowerAddr (owner)     */
    const int32_t* upper;      /* upperAddr (neighbour) */
    const double* diag;
    const double* upperCoeffs; /* A[l][u]: Apsi[l] += upper*psi[u] */
    const double* lowerCoeffs; /* A[u][l]: Apsi[u] += lower*psi[l]; == upperCoeffs for a symmetric matrix */
    const double* source;
    double* psi;               /* in: x0, out: x */
    int32_t nIfaces;
    const sm_iface* ifaces;
} sm_rank;

typedef struct sm_controls {
    double tolerance, relTol;
    int32_t maxIter, minIter;
    int32_t nSweeps;           /* smoothSolver::readControls: nSweeps 1 */
    int32_t smoother;          /* 0 GaussSeidel, 1 symGaussSeidel */
} sm_controls;

typedef struct sm_perf {
    double initialResidual, finalResidual, normFactor;
    int32_t nIterations, converged, singular, pad;
} sm_perf;

typedef struct sm_shared {
    int R;
    const sm_rank* ranks;
    sm_controls ctl;
    pthread_barrier_t bar;
    double* red;
    double*** sendbuf;         /* [R][nIfaces] -> packed psi[faceCells] */
    int** partner;
    sm_perf perf;
} sm_shared;

static double sm_reduce(sm_shared* sh, int rank, double v) {
    if (sh->R == 1) return v;
    sh->red[rank] = v;
    pthread_barrier_wait(&sh->bar);
    double s = sh->red[0];
    for (int r = 1; r < sh->R; ++r) s += sh->red[r];
    pthread_barrier_wait(&sh->bar);
    return s;
}

/* initMatrixInterfaces + updateMatrixInterfaces (OF-dev lduMatrixUpdateMatrixInterfaces.C,
 * processorFvPatchField.C): result[faceCells[i]] -= sign*bouCoeffs[i]*psi_nbr[i].  sign = -1 restates the
 * negated coefficients (mBouCoeffs) of the smoothers and of lduMatrix::residual. */
static void sm_send(sm_shared* sh, int rank, const double* x) {
    const sm_rank* m = &sh->ranks[rank];
    for (int k = 0; k < m->nIfaces; ++k) {
        double* sb = sh->sendbuf[rank][k];
        for (int32_t i = 0; i < m->ifaces[k].nFaces; ++i) sb[i] = x[m->ifaces[k].faceCells[i]];
    }
}
static void sm_update(sm_shared* sh, int rank, double* result, double sign) {
    const sm_rank* m = &sh->ranks[rank];
    if (sh->R > 1) pthread_barrier_wait(&sh->bar);
    for (int k = 0; k < m->nIfaces; ++k) {
        const sm_iface* I = &m->ifaces[k];
        const double* rb = sh->sendbuf[I->nbrRank][sh->partner[rank][k]];
        for (int32_t i = 0; i < I->nFaces; ++i) {
            const double c = sign < 0 ? -I->bouCoeffs[i] : I->bouCoeffs[i];
            result[I->faceCells[i]] -= c * rb[i];
        }
    }
    if (sh->R > 1) pthread_barrier_wait(&sh->bar);
}

/* lduMatrix::Amul (OF-dev lduMatrixATmul.C) with lower != upper */
static void sm_amul(sm_shared* sh, int rank, double* y, const double* x) {
    const sm_rank* m = &sh->ranks[rank];
    sm_send(sh, rank, x);
    for (int32_t c = 0; c < m->nCells; ++c) y[c] = m->diag[c] * x[c];
    for (int32_t f = 0; f < m->nFaces; ++f) {
        y[m->upper[f]] += m->lowerCoeffs[f] * x[m->lower[f]];
        y[m->lower[f]] += m->upperCoeffs[f] * x[m->upper[f]];
    }
    sm_update(sh, rank, y, +1.0);
}

/* lduMatrix::sumA (OF-dev lduMatrixATmul.C) */
static void sm_sumA(const sm_rank* m, double* s) {
    for (int32_t c = 0; c < m->nCells; ++c) s[c] = m->diag[c];
    for (int32_t f = 0; f < m->nFaces; ++f) {
        s[m->upper[f]] += m->lowerCoeffs[f];
        s[m->lower[f]] += m->upperCoeffs[f];
    }
    for (int k = 0; k < m->nIfaces; ++k)
        for (int32_t i = 0; i < m->ifaces[k].nFaces; ++i)
            s[m->ifaces[k].faceCells[i]] -= m->ifaces[k].bouCoeffs[i];
}

/* lduMatrix::residual (OF-dev lduMatrixATmul.C): rA = source - A psi, coupled coefficients negated */
static void sm_residual(sm_shared* sh, int rank, double* rA, const double* psi) {
    const sm_rank* m = &sh->ranks[rank];
    sm_send(sh, rank, psi);
    for (int32_t c = 0; c < m->nCells; ++c) rA[c] = m->source[c] - m->diag[c] * psi[c];
    for (int32_t f = 0; f < m->nFaces; ++f) {
        rA[m->upper[f]] -= m->lowerCoeffs[f] * psi[m->lower[f]];
        rA[m->lower[f]] -= m->upperCoeffs[f] * psi[m->upper[f]];
    }
    sm_update(sh, rank, rA, -1.0);
}

/* GaussSeidelSmoother::smooth / symGaussSeidelSmoother::smooth (OF-dev GaussSeidelSmoother.C,
 * symGaussSeidelSmoother.C).  ownStart = lduAddressing::ownerStartAddr(). */
static void sm_smooth(sm_shared* sh, int rank, double* psi, const int32_t* ownStart, double* bPrime, int sym,
                      int nSweeps) {
    const sm_rank* m = &sh->ranks[rank];
    const int32_t N = m->nCells;
    const int32_t* u = m->upper;
    for (int sweep = 0; sweep < nSweeps; ++sweep) {
        for (int32_t c = 0; c < N; ++c) bPrime[c] = m->source[c];
        sm_send(sh, rank, psi);
        sm_update(sh, rank, bPrime, -1.0);
        double psii;
        int32_t fStart, fEnd = ownStart[0];
        for (int32_t c = 0; c < N; ++c) {
            fStart = fEnd;
            fEnd = ownStart[c + 1];
            psii = bPrime[c];
            for (int32_t f = fStart; f < fEnd; ++f) psii -= m->upperCoeffs[f] * psi[u[f]];
            psii /= m->diag[c];
            for (int32_t f = fStart; f < fEnd; ++f) bPrime[u[f]] -= m->lowerCoeffs[f] * psii;
            psi[c] = psii;
        }
        if (!sym) continue;
        fStart = ownStart[N];
        for (int32_t c = N - 1; c >= 0; --c) {
            fEnd = fStart;
            fStart = ownStart[c];
            psii = bPrime[c];
            for (int32_t f = fStart; f < fEnd; ++f) psii -= m->upperCoeffs[f] * psi[u[f]];
            psii /= m->diag[c];
            for (int32_t f = fStart; f < fEnd; ++f) bPrime[u[f]] -= m->lowerCoeffs[f] * psii;
            psi[c] = psii;
        }
    }
}

static int sm_converged(const sm_perf* p, double tol, double relTol) {
    return (p->finalResidual < tol) || (relTol > 1e-20 && p->finalResidual < relTol * p->initialResidual);
}

typedef struct sm_worker { sm_shared* sh; int rank; } sm_worker;

/* smoothSolver::solve (OF-dev smoothSolver.C) -- literal control flow */
static void* sm_solve_worker(void* arg) {
    sm_worker* w = (sm_worker*)arg;
    sm_shared* sh = w->sh;
    const int rank = w->rank;
    const sm_rank* m = &sh->ranks[rank];
    const sm_controls* ctl = &sh->ctl;
    const int32_t N = m->nCells;
    const size_t nb = (size_t)(N > 0 ? N : 1) * sizeof(double);
    double* Apsi = (double*)malloc(nb);
    double* temp = (double*)malloc(nb);
    double* bPrime = (double*)malloc(nb);
    int32_t* ownStart = (int32_t*)calloc((size_t)N + 2, sizeof(int32_t));
    for (int32_t f = 0; f < m->nFaces; ++f) ownStart[m->lower[f] + 1]++;
    for (int32_t c = 0; c < N; ++c) ownStart[c + 1] += ownStart[c];
    double* psi = m->psi;
    sm_perf perf;
    memset(&perf, 0, sizeof(perf));

    if (ctl->nSweeps < 0) {
        /* fixed number of sweeps, no residual evaluation */
        sm_smooth(sh, rank, psi, ownStart, bPrime, ctl->smoother, -ctl->nSweeps);
        perf.nIterations -= ctl->nSweeps;
    } else {
        sm_amul(sh, rank, Apsi, psi);
        {
            sm_sumA(m, temp);
            double sPsi = 0.0;
            for (int32_t c = 0; c < N; ++c) sPsi += psi[c];
            const double gs = sm_reduce(sh, rank, sPsi);
            const double gn = sm_reduce(sh, rank, (double)N);
            const double xRef = gs / gn;
            for (int32_t c = 0; c < N; ++c) temp[c] *= xRef;
            double nf = 0.0;
            for (int32_t c = 0; c < N; ++c) nf += fabs(Apsi[c] - temp[c]) + fabs(m->source[c] - temp[c]);
            perf.normFactor = sm_reduce(sh, rank, nf) + 1e-20;
        }
        {
            double s = 0.0;
            for (int32_t c = 0; c < N; ++c) s += fabs(m->source[c] - Apsi[c]);
            perf.initialResidual = sm_reduce(sh, rank, s) / perf.normFactor;
            perf.finalResidual = perf.initialResidual;
        }
        if (ctl->minIter > 0 || !sm_converged(&perf, ctl->tolerance, ctl->relTol)) {
            do {
                sm_smooth(sh, rank, psi, ownStart, bPrime, ctl->smoother, ctl->nSweeps);
                sm_residual(sh, rank, temp, psi);
                double s = 0.0;
                for (int32_t c = 0; c < N; ++c) s += fabs(temp[c]);
                perf.finalResidual = sm_reduce(sh, rank, s) / perf.normFactor;
            } while (((perf.nIterations += ctl->nSweeps) < ctl->maxIter &&
                      !sm_converged(&perf, ctl->tolerance, ctl->relTol)) ||
                     perf.nIterations < ctl->minIter);
        }
        perf.converged = sm_converged(&perf, ctl->tolerance, ctl->relTol);
    }
    if (rank == 0) sh->perf = perf;
    free(Apsi); free(temp); free(bPrime); free(ownStart);
    return NULL;
}

static int sm_setup(sm_shared* sh, int R, const sm_rank* ranks) {
    memset(sh, 0, sizeof(*sh));
    sh->R = R;
    sh->ranks = ranks;
    sh->red = (double*)calloc((size_t)R, sizeof(double));
    sh->sendbuf = (double***)calloc((size_t)R, sizeof(double**));
    sh->partner = (int**)calloc((size_t)R, sizeof(int*));
    for (int r = 0; r < R; ++r) {
        const int nI = ranks[r].nIfaces;
        sh->sendbuf[r] = (double**)calloc((size_t)(nI > 0 ? nI : 1), sizeof(double*));
        sh->partner[r] = (int*)calloc((size_t)(nI > 0 ? nI : 1), sizeof(int));
        for (int k = 0; k < nI; ++k)
            sh->sendbuf[r][k] = (double*)calloc((size_t)(ranks[r].ifaces[k].nFaces > 0 ? ranks[r].ifaces[k].nFaces : 1),
                                               sizeof(double));
    }
    for (int r = 0; r < R; ++r)
        for (int k = 0; k < ranks[r].nIfaces; ++k) {
            const int q = ranks[r].ifaces[k].nbrRank;
            if (q < 0 || q >= R) return 1;
            int found = -1;
            for (int j = 0; j < ranks[q].nIfaces; ++j)
                if (ranks[q].ifaces[j].nbrRank == r) { found = j; break; }
            if (found < 0 || ranks[q].ifaces[found].nFaces != ranks[r].ifaces[k].nFaces) return 2;
            sh->partner[r][k] = found;
        }
    if (R > 1) pthread_barrier_init(&sh->bar, NULL, (unsigned)R);
    return 0;
}
static void sm_teardown(sm_shared* sh) {
    for (int r = 0; r < sh->R; ++r) {
        if (sh->sendbuf && sh->sendbuf[r]) {
            for (int k = 0; k < sh->ranks[r].nIfaces; ++k) free(sh->sendbuf[r][k]);
            free(sh->sendbuf[r]);
        }
        if (sh->partner) free(sh->partner[r]);
    }
    free(sh->sendbuf); free(sh->partner); free(sh->red);
    if (sh->R > 1) pthread_barrier_destroy(&sh->bar);
}

static int sm_run(int R, const sm_rank* ranks, sm_shared* sh, void* (*fn)(void*), sm_worker* ws) {
    if (R == 1) {
        fn(&ws[0]);
        return 0;
    }
    pthread_t* th = (pthread_t*)calloc((size_t)R, sizeof(pthread_t));
    for (int r = 0; r < R; ++r) pthread_create(&th[r], NULL, fn, &ws[r]);
    for (int r = 0; r < R; ++r) pthread_join(th[r], NULL);
    free(th);
    (void)ranks; (void)sh;
    return 0;
}

/* smoothSolver on R emulated ranks (R >= 1).  Returns 0, or > 0 when the interfaces do not pair up. */
int orc_smooth_solve(int R, const sm_rank* ranks, const sm_controls* ctl, sm_perf* perf) {
    sm_shared sh;
    int rc = sm_setup(&sh, R, ranks);
    if (rc) { sm_teardown(&sh); return rc; }
    sh.ctl = *ctl;
    sm_worker* ws = (sm_worker*)calloc((size_t)R, sizeof(sm_worker));
    for (int r = 0; r < R; ++r) { ws[r].sh = &sh; ws[r].rank = r; }
    sm_run(R, ranks, &sh, sm_solve_worker, ws);
    *perf = sh.perf;
    free(ws);
    sm_teardown(&sh);
    return 0;
}

/* Amul (what = 0) or residual (what = 1) alone on R emulated ranks: out[r] = A x[r]  |  source - A x[r] */
typedef struct sm_vec_job { sm_worker w; const double* x; double* y; int what; } sm_vec_job;
static void* sm_vec_worker(void* arg) {
    sm_vec_job* j = (sm_vec_job*)arg;
    if (j->what == 0) sm_amul(j->w.sh, j->w.rank, j->y, j->x);
    else sm_residual(j->w.sh, j->w.rank, j->y, j->x);
    return NULL;
}
int orc_asym_apply(int R, const sm_rank* ranks, const double* const* x, double* const* y, int what) {
    sm_shared sh;
    int rc = sm_setup(&sh, R, ranks);
    if (rc) { sm_teardown(&sh); return rc; }
    sm_vec_job* js = (sm_vec_job*)calloc((size_t)R, sizeof(sm_vec_job));
    pthread_t* th = (pthread_t*)calloc((size_t)R, sizeof(pthread_t));
    for (int r = 0; r < R; ++r) { js[r].w.sh = &sh; js[r].w.rank = r; js[r].x = x[r]; js[r].y = y[r]; js[r].what = what; }
    if (R == 1) sm_vec_worker(&js[0]);
    else {
        for (int r = 0; r < R; ++r) pthread_create(&th[r], NULL, sm_vec_worker, &js[r]);
        for (int r = 0; r < R; ++r) pthread_join(th[r], NULL);
    }
    free(js); free(th);
    sm_teardown(&sh);
    return 0;
}

/* one rank: sumA, and nSweeps sweeps of a smoother in place (no residual evaluation) */
void orc_asym_sumA(const sm_rank* m, double* s) { sm_sumA(m, s); }
