/*
 * b200pcg.h -- C ABI of the B200-native p_rgh pressure-correction hot path.
 *
 * This is the ONLY boundary between the host application (OpenFOAM/fireFoam through
 * adapter/B200PCG.C, or the Python ctypes mirror in firefoam-dev_b200/) and the sm_100a
 * CUDA library libb200pcg.so.  Plain C: pointers, sizes, PODs.  No C++ types, no
 * exceptions, no torch types.
 *
 * Each entry point names the reference interface it replaces.  The reference
 * (LeiXu84/fireFoam-dev 17.11.10) *calls* that interface at
 *     solver/pEqn.H:26-39        (assemble p_rghEqn, p_rghEqn.solve(...))
 *     solver/phrghEqn.H:43-48    (assemble ph_rghEqn, ph_rghEqn.solve())
 *     solver/pEqn.H:43-44        (p_rghEqn.flux())
 * and the implementation it reaches lives in un-vendored OpenFOAM-dev @ 940e28f6
 * (CHANGELOG:1-3): lduMatrix, PCG, DICPreconditioner, diagonalPreconditioner,
 * gaussLaplacianScheme, processorFvPatchField (restated in SURVEY.md Appendix A and in
 * oracle/pcg_oracle.c).
 *
 * Conventions
 *   - every function returns 0 on success, a B200_E* code otherwise; the text of the last
 *     error of a context is available from b200_last_error().
 *   - "host" entry points take host pointers, copy to the device and back inside the call;
 *     "_device" entry points take device pointers (same layout) and do no host<->device copy
 *     except an O(100 B) status read-back.
 *   - labels are int32 (WM_LABEL_SIZE=32), scalars are IEEE double (WM_PRECISION_OPTION=DP).
 *   - calls on one context must be serialised by the caller; with nranks > 1 every rank must
 *     issue the same sequence of collective calls (set_addressing, solve*), as OpenFOAM does.
 *   - non-convergence within maxIter and singularity are NOT errors; they are reported in
 *     b200_perf exactly like OpenFOAM's SolverPerformance.
 */
#ifndef B200PCG_C_ABI_H
#define B200PCG_C_ABI_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 2   /* 2: b200_dump grew lower / haveSmooth / smooth at its end; smoothSolver entry points */

/* error codes */
enum {
    B200_OK            = 0,
    B200_EINVAL        = 1,  /* bad argument (NULL, negative size, l>=u, unsorted faces ...)   */
    B200_ECUDA         = 2,  /* CUDA runtime error (message has cudaGetErrorString)           */
    B200_ENCCL         = 3,  /* NCCL error                                                    */
    B200_ENODEVICE     = 4,  /* no usable CUDA device: there is NO CPU fallback               */
    B200_ESTATE        = 5,  /* call out of order (solve before set_addressing ...)           */
    B200_EUNSUPPORTED  = 6,  /* unsupported interface type / row too long / label overflow    */
    B200_ENONFINITE    = 7,  /* non-finite residual encountered                               */
    B200_ENOMEM        = 8
};

/* preconditioner selector (fvSolution keyword `preconditioner`,
 * reference: cases/steckler/system/fvSolution:32) */
enum {
    B200_PRECOND_NONE      = 0,  /* `none`      -> noPreconditioner                              */
    B200_PRECOND_DIAGONAL  = 1,  /* `diagonal`  -> diagonalPreconditioner (bit-comparable path)  */
    B200_PRECOND_DIC_MC    = 2,  /* `DIC`       -> multicolour-ordered IC0 ("DIC-class"); the library picks
                                    the form: Eisenstat's (code 4) on bandwidth-bound systems with definite
                                    DIC pivots, else the three-kernel loop (code 5).  NOT OpenFOAM's DIC:
                                    iteration counts differ, the log says `DIC(mc)B200PCG`             */
    B200_PRECOND_DIC_EXACT = 3,  /* `DIC` + `B200{dicMode exact;}` -> level-scheduled DIC with
                                    the SAME elimination order as OpenFOAM's DICPreconditioner  */
    B200_PRECOND_DIC_MC_EIS = 4, /* `DIC` + `B200{dicMode eisenstat;}` -> the multicolour IC0 of code 2
                                    applied in Eisenstat's form: the two triangular sweeps also
                                    deliver A*p, so the iteration has no separate Amul (same iterates
                                    as code 5 up to rounding; 1e-8 solution parity bar)            */
    B200_PRECOND_DIC_MC_LOOP = 5 /* `DIC` + `B200{dicMode multicolour;}` -> the multicolour IC0 as a separate
                                    preconditioner apply (forward / backward colour sweeps) + Amul          */
};

typedef struct b200_ctx b200_ctx;

/* One coupled (processor) interface of this rank.  Mirrors what
 * lduMatrix::solver receives through `interfaces` + lduAddressing::patchAddr(k)
 * (OF-dev processorLduInterface). */
typedef struct b200_iface {
    int32_t        nbrRank;    /* neighbProcNo()                                   */
    int32_t        nFaces;     /* size of the patch                                */
    const int32_t* faceCells;  /* [nFaces] local cell of each patch face           */
    int32_t        tag;        /* message tag; unused by NCCL, kept for the adapter */
} b200_iface;

/* lduMatrix::solver::readControls (OF-dev lduMatrixSolver.C): maxIter 1000, minIter 0,
 * tolerance 1e-6, relTol 0 are the defaults the adapter fills in. */
typedef struct b200_controls {
    double  tolerance;
    double  relTol;
    int32_t maxIter;
    int32_t minIter;
    int32_t precond;      /* B200_PRECOND_*                                           */
    int32_t reserved;     /* must be 0                                                */
} b200_controls;

/* SolverPerformance<scalar> (OF-dev SolverPerformance.H): what the adapter needs to
 * build the `DICPCG:  Solving for p_rgh, Initial residual = ..` log line
 * (format: cases/steckler/original/linux64/log.fireFoam:92). */
typedef struct b200_perf {
    double  initialResidual;
    double  finalResidual;
    double  normFactor;       /* extra: lduMatrix::solver::normFactor value         */
    int32_t nIterations;
    int32_t converged;
    int32_t singular;
    int32_t nColours;         /* extra: colours/levels used by the DIC-class sweep   */
    double  solveMs;          /* extra: device time of the PCG loop (CUDA events)     */
    double  setupMs;          /* extra: device time of per-solve set-up               */
    double  h2dMs, d2hMs;     /* extra: copy times of the host entry points          */
} b200_perf;

/* ---- context ------------------------------------------------------------------------- */

/* Replaces: Pstream/MPI initialisation of the solver's communication (OF-dev UPstream.C).
 * device < 0 selects the current device.  nccl_uid: 128 bytes from b200_get_unique_id on
 * rank 0, broadcast by the caller (Pstream::scatter in the adapter, torch.distributed in the
 * harness); NULL iff nranks == 1. */
int b200_ctx_create(int device, int rank, int nranks, const void* nccl_uid, b200_ctx** out);
int b200_get_unique_id(void* uid128);
void b200_ctx_destroy(b200_ctx* ctx);
const char* b200_last_error(const b200_ctx* ctx);   /* ctx may be NULL: last create error */
int b200_abi_version(void);
int b200_device_count(void);                        /* 0 when no GPU/driver: callers must fail */

/* Replaces: lduAddressing (lowerAddr/upperAddr/patchAddr + lazily built losort/ownerStart;
 * OF-dev lduAddressing.C).  Idempotent for an unchanged mesh_key (upstream constructs a new
 * solver object per solve, so the adapter calls this every time).  Faces must be in
 * upper-triangular order (lowerAddr non-decreasing, l < u).  Builds the device-resident row
 * structure, the colourings (lazily, on first DIC-class solve) and the halo plans. */
int b200_set_addressing(b200_ctx* ctx, uint64_t mesh_key,
                        int32_t nCells, int32_t nFaces,
                        const int32_t* lowerAddr, const int32_t* upperAddr,
                        int32_t nIfaces, const b200_iface* ifaces);

/* ---- assembly: fvm::laplacian(gamma, vf) --------------------------------------------- */

/* Replaces: gaussLaplacianScheme<scalar,scalar>::fvmLaplacianUncorrected + lduMatrix::negSumDiag
 * (OF-dev gaussLaplacianScheme.C, lduMatrixOperations.C), called from solver/pEqn.H:32 (sign=-1,
 * because of `- fvm::laplacian`) and solver/phrghEqn.H:45 (sign=+1).
 *   upper_out[f] = sign * deltaCoeffs[f] * (gamma_f[f] * magSf[f])
 *   diag_inout[c] += -( sum of upper_out over faces of c )        (sorted-segment sums, no atomics)
 * diag_inout may hold the ddt contribution already (pass zeros for a pure Laplacian). */
int b200_assemble_laplacian(b200_ctx* ctx, const double* gamma_f, const double* magSf,
                            const double* deltaCoeffs, double sign,
                            double* upper_out, double* diag_inout);
int b200_assemble_laplacian_device(b200_ctx* ctx, const double* d_gamma_f, const double* d_magSf,
                                   const double* d_deltaCoeffs, double sign,
                                   double* d_upper_out, double* d_diag_inout);

/* ---- assembly of the whole p_rghEqn -------------------------------------- */

/* Replaces, on the device, the fvMatrix algebra of solver/pEqn.H:26-37
 *     fvm::ddt(psi, p_rgh) + fvc::ddt(psi, rho)*gh + fvc::ddt(psi)*pRef + fvc::div(phiHbyA)
 *   - fvm::laplacian(rhorAUf, p_rgh) == parcels.Srho() + surfaceFilm.Srho() + fvOptions(...)
 * (OF-dev EulerDdtScheme.C fvmDdt, fvMatrix.C operator+/-/==, surfaceIntegrate.C,
 * gaussLaplacianScheme.C, lduMatrixOperations.C negSumDiag) AND the boundary fold of
 * fvMatrix::solveSegregated (addBoundaryDiag / addBoundarySource), so that what comes
 * out is exactly what lduMatrix::solver::solve receives: upper, diag (with internalCoeffs) and
 * totalSource.  With psi == NULL, phi = phig, divSign = +1, lapSign = +1 it is the ph_rghEqn of
 * solver/phrghEqn.H:43-46.  Every sum is formed in OpenFOAM's face / patch order (sorted segments, no
 * atomics): bit-identical to the CPU operator sequence.
 *   diag   = rDeltaT*psi*V  -/+ negSumDiag(laplacian)  + sum internalCoeffs
 *   source = rDeltaT*psi0*p0*V - sum_k V*explicit_k  -/+ V*(surfaceIntegrate(phi)/V) + V*Su + sum boundaryCoeffs
 * The adapter evaluates the explicit cell fields (fvc::ddt(psi,rho)*gh, ...) and the patch coefficients
 * with OpenFOAM itself; they arrive as plain arrays.  At most 8 explicit fields. */
typedef struct b200_prgh_terms {
    double rDeltaT;                       /* 1/deltaT (Euler)                                      */
    const double *V, *psi, *psi0, *p0;    /* [nCells] mesh.V(), psi, psi.oldTime(), p_rgh.oldTime();
                                             psi == NULL: no ddt term                              */
    int32_t nExplicit, pad0;
    const double* const* explicitFields;  /* [nExplicit] pointers to [nCells]: source -= V*field   */
    const double* phi;                    /* [nFaces] flux of the fvc::div term, or NULL           */
    double divSign;                       /* -1: `+ fvc::div(phi)` on the left, +1: `== fvc::div(phi)` */
    const double *gamma_f, *magSf, *deltaCoeffs;   /* [nFaces] laplacian inputs                    */
    double lapSign;                       /* -1: `- fvm::laplacian`, +1: `fvm::laplacian`          */
    const double* Su;                     /* [nCells] explicit source `== Su` (source += V*Su), or NULL */
    int32_t nB, pad1;                     /* boundary faces (b200_set_boundary_faces)              */
    const int32_t* bCells;                /* unused by the device entry point (set once per mesh)  */
    const double *bPhi, *bInternal, *bBoundary;   /* [nB] boundary flux of phi, internalCoeffs (all
                                             patches), boundaryCoeffs (0 on coupled patches); NULL = none */
} b200_prgh_terms;

/* faceCells of every boundary face, patch by patch (mesh constant; idempotent for unchanged input) */
int b200_set_boundary_faces(b200_ctx* ctx, int32_t nB, const int32_t* bCells);
int b200_assemble_p_rgh(b200_ctx* ctx, const b200_prgh_terms* t, double* upper_out, double* diag_out,
                        double* source_out);
/* all array pointers inside *t are DEVICE pointers (the struct and explicitFields[] live on the host) */
int b200_assemble_p_rgh_device(b200_ctx* ctx, const b200_prgh_terms* t, double* d_upper_out,
                               double* d_diag_out, double* d_source_out);

/* ---- solve: lduMatrix::solver::solve ------------------------------------------------- */

/* Replaces: PCG::solve(psi, source, cmpt) with its Amul/normFactor/preconditioner/gSum*
 * calls (OF-dev PCG.C, lduMatrixATmul.C, lduMatrixSolver.C, DICPreconditioner.C,
 * diagonalPreconditioner.C), reached from solver/pEqn.H:39 and solver/phrghEqn.H:48.
 *   diag   [nCells]  matrix diagonal WITH boundary internalCoeffs already added
 *   upper  [nFaces]  symmetric off-diagonal
 *   ifaceBouCoeffs[k] [ifaces[k].nFaces]  interfaceBouCoeffs of coupled patch k (may be NULL if
 *                    nIfaces == 0)
 *   source [nCells]  totalSource
 *   psi    [nCells]  in: initial guess, out: solution
 */
int b200_solve(b200_ctx* ctx, const double* diag, const double* upper,
               const double* const* ifaceBouCoeffs, const double* source, double* psi,
               const b200_controls* ctl, b200_perf* perf);
int b200_solve_device(b200_ctx* ctx, const double* d_diag, const double* d_upper,
                      const double* const* d_ifaceBouCoeffs, const double* d_source, double* d_psi,
                      const b200_controls* ctl, b200_perf* perf);

/* Replaces: lduMatrix::Amul(Apsi, psi, interfaceBouCoeffs, interfaces, cmpt) (OF-dev
 * lduMatrixATmul.C) -- exposed so that parity tests can check the SpMV alone.  Collective when
 * nranks > 1.  Row sums are formed in OpenFOAM's face order without FMA contraction, so on one
 * rank the result is bit-identical to the CPU loop. */
int b200_amul(b200_ctx* ctx, const double* diag, const double* upper,
              const double* const* ifaceBouCoeffs, const double* psi, double* Apsi);

/* Replaces: fvMatrix<scalar>::flux() internal-face part (OF-dev fvMatrix.C; used at
 * solver/pEqn.H:43-44): flux[f] = upper[f]*psi[u[f]] - upper[f]*psi[l[f]]. */
int b200_flux(b200_ctx* ctx, const double* upper, const double* psi, double* flux_out);

/* ---- the other linear solves of a time step: smoothSolver on asymmetric matrices ---- */

/* smoother selector (fvSolution keyword `smoother`; reference: cases/steckler/system/fvSolution:51) */
enum {
    B200_SMOOTHER_GAUSS_SEIDEL     = 0,   /* `GaussSeidel`    -> GaussSeidelSmoother (forward sweep)            */
    B200_SMOOTHER_SYM_GAUSS_SEIDEL = 1    /* `symGaussSeidel` -> symGaussSeidelSmoother (forward + reverse)     */
};
/* order in which a sweep visits the cells */
enum {
    B200_SWEEP_MULTICOLOUR = 0,  /* default: multicolour ordering of the same smoother ("GS-class": one launch per
                                    colour, 2 colours on hex meshes).  NOT OpenFOAM's cell order: sweep counts and
                                    tolerance-limited solutions differ slightly, the log says `B200smoothSolver(mc)` */
    B200_SWEEP_EXACT       = 1   /* `B200{sweepMode exact;}`: level-scheduled sweeps in OpenFOAM's OWN elimination
                                    order (cell order), row sums in its face order: bit-identical psi after every
                                    sweep, identical sweep counts; one launch per dependency level (a parity tool on
                                    large meshes, like dicMode exact)                                              */
};

/* smoothSolver::readControls (OF-dev smoothSolver.C, lduMatrixSolver.C): the lduMatrix::solver controls + nSweeps 1 */
typedef struct b200_smooth_controls {
    double  tolerance;
    double  relTol;
    int32_t maxIter;
    int32_t minIter;
    int32_t nSweeps;      /* sweeps between two residual evaluations; < 0: exactly -nSweeps sweeps, no residuals  */
    int32_t smoother;     /* B200_SMOOTHER_*                                                                      */
    int32_t sweepMode;    /* B200_SWEEP_*                                                                         */
    int32_t reserved;     /* must be 0                                                                            */
} b200_smooth_controls;

/* Replaces: smoothSolver::solve(psi, source, cmpt) with its Amul / normFactor / residual / smoother calls (OF-dev
 * smoothSolver.C, GaussSeidelSmoother.C, symGaussSeidelSmoother.C, lduMatrixATmul.C, lduMatrixSolver.C) -- what
 * the reference selects for U, Yi, h and k (cases/steckler/system/fvSolution:48-61; solver/UEqn.H:19-30,
 * solver/YEEqn.H:60,111) on ASYMMETRIC lduMatrices.
 *   diag   [nCells]  matrix diagonal WITH boundary internalCoeffs already added
 *   upper  [nFaces]  A[l][u]   (Amul: Apsi[l] += upper*psi[u])
 *   lower  [nFaces]  A[u][l]   (Amul: Apsi[u] += lower*psi[l]); NULL: symmetric matrix, lower aliases upper
 *   source, psi, ifaceBouCoeffs as in b200_solve.
 * Collective when nranks > 1: processor patches are explicit (Jacobi-like) contributions refreshed once per sweep,
 * as upstream (bPrime = source; updateMatrixInterfaces with the negated interfaceBouCoeffs). */
int b200_smooth_solve(b200_ctx* ctx, const double* diag, const double* upper, const double* lower,
                      const double* const* ifaceBouCoeffs, const double* source, double* psi,
                      const b200_smooth_controls* ctl, b200_perf* perf);
int b200_smooth_solve_device(b200_ctx* ctx, const double* d_diag, const double* d_upper, const double* d_lower,
                             const double* const* d_ifaceBouCoeffs, const double* d_source, double* d_psi,
                             const b200_smooth_controls* ctl, b200_perf* perf);

/* Replaces: lduMatrix::Amul with lower != upper (OF-dev lduMatrixATmul.C) -- the SpMV of the asymmetric path alone,
 * for parity tests.  Row sums in OpenFOAM's face order, no FMA contraction: bit-identical to the CPU loop. */
int b200_amul_asym(b200_ctx* ctx, const double* diag, const double* upper, const double* lower,
                   const double* const* ifaceBouCoeffs, const double* psi, double* Apsi);

/* Replaces: PBiCG::solve(psi, source, cmpt) with its Amul / Tmul / normFactor / preconditioner calls (OF-dev PBiCG.C,
 * DILUPreconditioner.C, diagonalPreconditioner.C, lduMatrixATmul.C) -- what the reference's other cases select for
 * their transport equations (cases/wallFireSpread2D/system/fvSolution:66-73 `Yi { solver PBiCG; preconditioner DILU; }`,
 * cases/pyrolysis1D, the pyrolysis / panel regions) and what produced its 2.4.x golden logs of steckler (207
 * `DILUPBiCG:` lines in cases/steckler/original/darwinIntel64/log.fireFoam).  Matrix arguments as in
 * b200_smooth_solve.  ctl->precond: B200_PRECOND_NONE, B200_PRECOND_DIAGONAL, B200_PRECOND_DILU_MC (`DILU`: the
 * multicolour-ordered stand-in, "DILU-class", log name DILU(mc)B200PBiCG) or B200_PRECOND_DILU_EXACT
 * (`B200{diluMode exact;}`: level-scheduled DILU in OpenFOAM's own elimination order, identical iteration counts).
 * ifaceIntCoeffs: interfaceIntCoeffs of the coupled patches (Tmul uses them upstream).  One rank only in this version:
 * with nranks > 1 the call returns B200_EUNSUPPORTED. */
enum { B200_PRECOND_DILU_MC = 2, B200_PRECOND_DILU_EXACT = 3 };   /* the DIC codes, read on an asymmetric matrix */
int b200_bicg_solve(b200_ctx* ctx, const double* diag, const double* upper, const double* lower,
                    const double* const* ifaceBouCoeffs, const double* const* ifaceIntCoeffs, const double* source,
                    double* psi, const b200_controls* ctl, b200_perf* perf);
int b200_bicg_solve_device(b200_ctx* ctx, const double* d_diag, const double* d_upper, const double* d_lower,
                           const double* const* d_ifaceBouCoeffs, const double* const* d_ifaceIntCoeffs,
                           const double* d_source, double* d_psi, const b200_controls* ctl, b200_perf* perf);

/* ---- harness helpers (not part of the OpenFOAM-facing contract) ---------------------- */

/* pinned host memory for callers that want async H2D at full PCIe rate */
int  b200_host_alloc(void** p, size_t bytes);
void b200_host_free(void* p);
/* number of kernel launches issued by this context since creation (bench "gpu_launches") */
uint64_t b200_launch_count(const b200_ctx* ctx);
/* fixed-iteration timing aid: when > 0 the convergence test is ignored and exactly n loop
 * bodies are executed (bench only; 0 restores OpenFOAM semantics) */
int b200_debug_force_iterations(b200_ctx* ctx, int32_t n);
/* profiling aid: when on, every kernel class is timed with CUDA events and the totals are
 * retrievable as a JSON string (B200PCG_PROFILE=1 in the adapter) */
int b200_profile_enable(b200_ctx* ctx, int on);
const char* b200_profile_json(b200_ctx* ctx);
/* which kernel variants / launch geometry the context selected for the current mesh, as a JSON
 * string (bench "roofline.kernel"; valid until the next call on this context) */
const char* b200_describe(b200_ctx* ctx);

/* ---- matrix dump / replay format (layout in csrc/dump.cpp) ---------------- */

/* What lduMatrix::solver::solve receives on one rank (reference call sites solver/pEqn.H:39,
 * solver/phrghEqn.H:48), plus what it reported.  Written by the adapter when B200PCG_DUMP=<dir>
 * is set, one file per rank and solve; read back by tools/b200replay and
 * firefoam-dev_b200/replay.py.  Host-only: needs no GPU. */
typedef struct b200_dump {
    const char*    fieldName;       /* "p_rgh", "ph_rgh", ...                                  */
    int32_t        rank, nranks;
    int32_t        nCells, nFaces;
    const int32_t* lowerAddr;       /* [nFaces]                                                */
    const int32_t* upperAddr;       /* [nFaces]                                                */
    const double*  diag;            /* [nCells] with boundary internalCoeffs   */
    const double*  upper;           /* [nFaces]                                                */
    const double*  source;          /* [nCells] totalSource                                    */
    const double*  psi0;            /* [nCells] initial guess                                  */
    const double*  psiSolution;     /* [nCells] solution, or NULL                              */
    int32_t        nIfaces;
    const b200_iface* ifaces;       /* [nIfaces]                                               */
    const double* const* ifaceBouCoeffs;   /* [nIfaces][ifaces[k].nFaces]                      */
    b200_controls  controls;
    int32_t        havePerf;        /* perf + solverName below are valid                       */
    b200_perf      perf;            /* what the solver that ran in the host application reported */
    const char*    solverName;      /* e.g. "DICPCG"                                           */
    int32_t        solveIndex;      /* running number of the solve on this rank                */
    double         time;            /* runTime.value()                                         */
    /* ABI version 2: asymmetric systems and smoothSolver solves */
    const double*  lower;           /* [nFaces] A[u][l], or NULL: symmetric (lower aliases upper) */
    int32_t        haveSmooth;      /* written by B200smoothSolver: `smooth` holds the controls, `controls` is unused */
    int32_t        havePBiCG;       /* written by B200PBiCG: `controls` holds PBiCG's (precond: 0 none, 1 diagonal,
                                       B200_PRECOND_DILU_MC, B200_PRECOND_DILU_EXACT)                              */
    b200_smooth_controls smooth;
} b200_dump;

typedef struct b200_dump_file b200_dump_file;

int b200_dump_write(const char* path, const b200_dump* d);
/* reads a dump; the arrays of b200_dump_get() point into memory owned by *out */
int b200_dump_read(const char* path, b200_dump_file** out);
const b200_dump* b200_dump_get(const b200_dump_file* f);
const char* b200_dump_header_json(const b200_dump_file* f);
void b200_dump_free(b200_dump_file* f);
const char* b200_dump_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* B200PCG_C_ABI_H */
