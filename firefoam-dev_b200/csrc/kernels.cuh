// kernels.cuh -- hand-written sm_100a kernels of the p_rgh PCG hot path.
//
// Everything here is HBM-bandwidth-bound fp64 sparse/vector work (arithmetic intensity
// ~0.1-0.2 flop/B): no tensor cores.  Design rules applied throughout:
//   * persistent grid-stride kernels, grid = k x SM count, 256-thread blocks;
//   * every streaming array is read once, coalesced (sliced-ELL for the matrix, double2 for
//     vectors); gathered x[col] goes through the read-only path;
//   * reductions are deterministic: warp shuffles -> shared -> one partial per block ->
//     the LAST block to finish (integer ticket) sums the partials in a fixed order and runs
//     the scalar step of the CG recurrence on the device (no float atomics, no host sync);
//   * no FMA contraction (__dmul_rn/__dadd_rn) so that element-wise results are bit-identical
//     to OpenFOAM's scalar loops (gcc x86-64 without -mfma never fuses); only the order of the
//     global sums differs from the CPU.
//
// Reference algorithm text: SURVEY.md Appendix A (restating OF-dev PCG.C, lduMatrixATmul.C,
// lduMatrixSolver.C, DICPreconditioner.C, diagonalPreconditioner.C, gaussLaplacianScheme.C).
#pragma once
#include <cstdint>
#include <cooperative_groups.h>
#include <cuda_runtime.h>

namespace b200 {

constexpr int kBlock = 256;
constexpr int kMaxGrid = 148 * 16;   // upper bound of any reducing grid (partials capacity)
constexpr int kNSums = 2;

// CG recurrence steps executed on the device by the last block of a reducing kernel
enum Step : int {
    STEP_NONE = 0,
    STEP_SUMPSI,   // xRef = gAverage(psi)
    STEP_NORM,     // normFactor, initial residual, first convergence check
    STEP_WARA,     // wArA = (wA, rA); beta
    STEP_WAPA,     // wApA = (wA, pA); singularity; alpha
    STEP_RES,      // final residual; nIter++; convergence
    STEP_RES_WARA, // STEP_RES, then (if the loop continues) STEP_WARA of the next iteration
    // Eisenstat form of the DIC-class PCG (k_eis_* kernels)
    STEP_EIS_SIGN, // sign of the DIC pivots (sigma), or "unusable"
    STEP_EIS_RHO0, // rho_0 = (r^, r^); calibrates the true-residual predictor
    STEP_EIS_RHO,  // rho_k; beta; nIter++; decides whether this iteration needs a true-residual check
    STEP_EIS_RES,  // true residual |(D~+L) r^|_1 / normFactor; convergence; loop condition
    // smoothSolver (k_gs_* kernels)
    STEP_GS_RES    // final residual after nSweeps sweeps; nIter += nSweeps; convergence; loop condition
};

// Eisenstat form: how many iterations may pass between two evaluations of the true residual, as a function
// of q = predicted residual / convergence threshold (predictor: cRatio*sqrt(rho), re-calibrated at every
// evaluation).  Far from convergence the evaluation only re-calibrates the predictor; close to it every
// iteration is checked, so the loop stops within an iteration or two of the first crossing.
__host__ __device__ inline int eis_check_interval(double q) {
    return q < 1.5 ? 1 : (q < 4.0 ? 2 : (q < 32.0 ? 8 : 32));
}

struct Scalars {
    // controls (written by the host before each solve)
    double tol, relTol;
    double nGlobalCells;
    int maxIter, minIter, forceIters, nranks;
    int nSweeps, padS;     // smoothSolver: sweeps between two residual evaluations (STEP_GS_RES)
    // state
    double xRef, normFactor, initRes, finalRes;
    double wArA, wArAold, wApA, alpha, beta;
    int nIter, done, converged, singular, nonfinite, pendingPsi;
    // reduction plumbing
    double acc[kNSums];    // running totals across the kernels of one reduction
    double sums[kNSums];   // local totals (input of the all-reduce when nranks > 1)
    double gsums[kNSums];  // global totals (== sums when nranks == 1)
    double acc2;           // correction of (y, x) by the interface terms (k_iface_pre -> k_iface_apply)
    unsigned int ticket;
    unsigned int ticket2;  // block ticket of k_iface_pre, which reduces on the comm stream WHILE the Amul reduces on
                           // the compute stream (its running total is acc2, its partials a second arena)
    // Eisenstat form (STEP_EIS_*): predictor ratio  true residual / sqrt(|rho|)  at the last check,
    // iterations since that check, and whether the current iteration's check kernel has to run
    double cRatio;
    double sigma;          // +1 / -1: common sign of the DIC pivots (0: mixed or zero pivots -> error)
    int sinceCheck, needCheck;
    // wait accounting of the current solve (device nanosecond timer, one thread each): time the consumer's first
    // block spent waiting for its neighbours' halo flags, time the reducing kernels' last block spent waiting for
    // the peers' partial sums, and the number of waits of each kind -- what "exposed communication" is made of
    unsigned long long waitHaloNs, waitRedNs;
    unsigned int nHaloWaits, nRedWaits;
    // number of cross-rank reductions this rank has EXECUTED (peer_allreduce_step).  Every rank
    // executes the same sequence (identical totals -> identical `done` decisions), so the counters
    // stay equal across ranks, and consecutive executed reductions strictly alternate buffer parity --
    // which is what makes the two-slot PeerBuf safe (a host-side launch counter would not: launches
    // skipped by `done` would let two executed reductions share a parity).
    unsigned long long redSeq;
    // processor-patch halo over peer memory (k_pack_p2p / halo_acquire): number of exchanges this rank has
    // EXECUTED (same rule as redSeq: neighbours execute the same sequence, so their counters agree and consecutive
    // exchanges alternate the two halves of the receive buffer) and the producer's block ticket
    unsigned long long haloSeq;
    unsigned int haloTicket, pad2;
};

struct PeerBuf;
struct Reduce {
    Scalars* S;
    double* partials;   // [kNSums][kMaxGrid]
    int step;           // Step to run when this kernel completes the reduction (STEP_NONE:
                        // only accumulate into S->acc)
    // nranks > 1 with peer-memory reductions: the last block of the reducing kernel performs the
    // cross-rank sum itself (peer_allreduce_step); peers == nullptr -> host issues ncclAllReduce
    PeerBuf* const* peers;
    int rank;
};

// ---- scalar steps (OF-dev PCG.C, SolverPerformance.C) ---------------
__device__ __forceinline__ bool check_convergence(const Scalars* S) {
    // SolverPerformance::checkConvergence: final < Tolerance ||
    //   (RelTolerance > small_*pTraits::one && final < RelTolerance*initial); small_ = 1e-20
    return (S->finalRes < S->tol) || (S->relTol > 1e-20 && S->finalRes < S->relTol * S->initRes);
}

__device__ inline void scalar_step(int step, Scalars* S, const double* g) {
    switch (step) {
        case STEP_SUMPSI:
            S->xRef = g[0] / S->nGlobalCells;
            break;
        case STEP_NORM: {
            S->normFactor = g[0] + 1e-20;   // + small_
            S->initRes = g[1] / S->normFactor;
            S->finalRes = S->initRes;
            S->nIter = 0;
            S->wArA = 1e20;                 // great_
            S->wArAold = 1e20;
            S->singular = 0;
            S->nonfinite = 0;
            S->pendingPsi = 0;
            bool conv = check_convergence(S);
            S->converged = conv ? 1 : 0;
            bool enter = S->forceIters > 0 ? true : (S->minIter > 0 || !conv);
            S->done = enter ? 0 : 1;
            if (!(S->initRes == S->initRes) || fabs(S->initRes) > 1.7e308) {
                S->nonfinite = 1;
                S->done = 1;
            }
            break;
        }
        case STEP_WARA:
            S->wArAold = S->wArA;
            S->wArA = g[0];
            S->beta = (S->nIter == 0) ? 0.0 : S->wArA / S->wArAold;
            break;
        case STEP_WAPA:
            S->wApA = g[0];
            // checkSingularity(mag(wApA)/normFactor): singular <=> value <= vSmall_ (1e-300)
            if (S->forceIters == 0 && !(fabs(S->wApA) / S->normFactor > 1e-300)) {
                S->singular = 1;
                S->done = 1;
                S->pendingPsi = 0;
            } else {
                S->alpha = S->wArA / S->wApA;
                S->pendingPsi = 1;   // psi += alpha*pA is applied by the next k_p / k_psi_final
            }
            break;
        case STEP_RES:
        case STEP_RES_WARA: {
            S->finalRes = g[0] / S->normFactor;
            int old = S->nIter;
            S->nIter = old + 1;
            bool conv = check_convergence(S);
            S->converged = conv ? 1 : 0;
            bool cont;
            if (S->forceIters > 0) cont = S->nIter < S->forceIters;
            else cont = (old < S->maxIter && !conv) || (S->nIter < S->minIter);
            if (!(S->finalRes == S->finalRes) || fabs(S->finalRes) > 1.7e308) {
                S->nonfinite = 1;
                cont = false;
            }
            if (!cont) S->done = 1;
            else if (step == STEP_RES_WARA) {
                S->wArAold = S->wArA;
                S->wArA = g[1];
                S->beta = S->wArA / S->wArAold;
            }
            break;
        }
        case STEP_EIS_SIGN: {
            // g[0] = number of negative DIC pivots, g[1] = number of zero / non-finite ones (global counts)
            S->sigma = 1.0;
            if (S->done) break;                       // converged before the first iteration: nothing to scale
            if (g[1] > 0.0 || (g[0] > 0.0 && g[0] < S->nGlobalCells)) {
                S->sigma = 0.0;
                S->nonfinite = 3;                     // indefinite / singular pivots: reported as an error
                S->done = 1;
            } else if (g[0] > 0.0) {
                S->sigma = -1.0;
            }
            break;
        }
        case STEP_EIS_RHO0: {
            S->wArA = g[0];
            S->beta = 0.0;
            const double a = sqrt(fabs(g[0]));
            S->cRatio = a > 0.0 ? S->finalRes / a : 0.0;   // 0: predictor says "converged" -> always check
            S->sinceCheck = 0;
            S->needCheck = 0;
            break;
        }
        case STEP_EIS_RHO: {
            S->wArAold = S->wArA;
            S->wArA = g[0];
            S->beta = S->wArA / S->wArAold;
            const int old = S->nIter;
            S->nIter = old + 1;
            S->sinceCheck += 1;
            if (!(g[0] == g[0]) || fabs(g[0]) > 1.7e308) {
                S->nonfinite = 1;
                S->needCheck = 0;
                S->done = 1;
                break;
            }
            // the do/while cannot continue whatever the residual turns out to be -> final check
            const bool mustStop = S->forceIters > 0 ? (S->nIter >= S->forceIters) : !(old < S->maxIter);
            double thr = S->tol;
            if (S->relTol > 1e-20 && S->relTol * S->initRes > thr) thr = S->relTol * S->initRes;
            const int every = S->forceIters > 0 ? 32 : eis_check_interval(S->cRatio * sqrt(fabs(g[0])) / thr);
            S->needCheck = (mustStop || S->sinceCheck >= every) ? 1 : 0;
            break;
        }
        case STEP_EIS_RES: {
            S->finalRes = g[0] / S->normFactor;
            const bool conv = check_convergence(S);
            S->converged = conv ? 1 : 0;
            const int old = S->nIter - 1;
            bool cont;
            if (S->forceIters > 0) cont = S->nIter < S->forceIters;
            else cont = (old < S->maxIter && !conv) || (S->nIter < S->minIter);
            if (!(S->finalRes == S->finalRes) || fabs(S->finalRes) > 1.7e308) {
                S->nonfinite = 1;
                cont = false;
            }
            if (!cont) S->done = 1;
            const double a = sqrt(fabs(S->wArA));
            S->cRatio = a > 0.0 ? S->finalRes / a : 0.0;
            S->sinceCheck = 0;
            S->needCheck = 0;
            break;
        }
        case STEP_GS_RES: {
            // smoothSolver::solve (OF-dev smoothSolver.C): do { smooth(nSweeps); finalResidual = ... }
            // while (((nIterations += nSweeps) < maxIter && !converged) || nIterations < minIter)
            S->finalRes = g[0] / S->normFactor;
            S->nIter += S->nSweeps;
            const bool conv = check_convergence(S);
            S->converged = conv ? 1 : 0;
            bool cont;
            if (S->forceIters > 0) cont = S->nIter < S->forceIters;
            else cont = (S->nIter < S->maxIter && !conv) || (S->nIter < S->minIter);
            if (!(S->finalRes == S->finalRes) || fabs(S->finalRes) > 1.7e308) {
                S->nonfinite = 1;
                cont = false;
            }
            if (!cont) S->done = 1;
            break;
        }
        default:
            break;
    }
}

// After an all-reduce (nranks > 1): one thread runs the scalar step on the global sums.
__global__ void k_scalar_step(Scalars* S, int step) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (S->done && step != STEP_SUMPSI && step != STEP_NORM) return;
    if (step == STEP_EIS_RES && !S->needCheck) return;   // the check kernel did not run
    scalar_step(step, S, S->gsums);
}

// ---- one-shot all-reduce over NVLink peer memory, fused with the scalar step ------------------
// Replaces reduce(scalar, sumOp) of OF-dev UPstream.C for the 2-double CG reductions.  Every rank
// owns a PeerBuf in its HBM, mapped into all other ranks with CUDA IPC.  One warp per rank: lane r
// stores this rank's partial sums + a sequence flag into rank r's buffer (P2P stores through
// NVSwitch), then waits for rank r's contribution to arrive in the local buffer, and lane 0 adds
// the contributions in ASCENDING RANK ORDER -- the order of Pstream's linear gather for
// <= nProcsSimpleSum ranks -- so every rank forms bit-identical totals, and runs
// the scalar step.  Executed by the first warp of the LAST block of the kernel that completes the
// reduction (reduce_finish): no extra launch, instead of ncclAllReduce + a scalar kernel (~25 us).
constexpr int kMaxRanks = 32;
constexpr unsigned long long kPeerTimeoutNs = 5000000000ull;
struct PeerBuf {
    double vals[2][kMaxRanks][kNSums];
    unsigned long long flags[2][kMaxRanks];
};

// system-scope release store / acquire load of a 64-bit flag (peer memory over NVLink).  A release store orders
// the thread's earlier writes (and, cumulatively, writes it has synchronised with) before the flag; an acquire
// load orders the thread's later reads after it -- the full two-way membar.sys of __threadfence_system() on
// both sides of every flag cost ~4 us per wait on B200 (wait accounting at 4 GPUs, profiles/r02_*).
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// one warp (all 32 lanes must call)
__device__ __forceinline__ void peer_allreduce_step(Scalars* S, PeerBuf* const* peers, int rank, int nranks,
                                                    int step) {
    const int lane = threadIdx.x & 31;
    unsigned long long seq = 0;
    if (lane == 0) {
        seq = S->redSeq + 1ull;
        S->redSeq = seq;
    }
    seq = __shfl_sync(0xffffffffu, seq, 0);
    const int par = (int)(seq & 1ull);
    if (lane < nranks) {
        PeerBuf* dst = peers[lane];
#pragma unroll
        for (int i = 0; i < kNSums; ++i) dst->vals[par][rank][i] = S->sums[i];
        st_release_sys(&dst->flags[par][rank], seq);
    }
    double v[kNSums];
#pragma unroll
    for (int i = 0; i < kNSums; ++i) v[i] = 0.0;
    bool timedOut = false;
    unsigned long long tw0 = 0;
    if (lane == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tw0));
    if (lane < nranks) {
        PeerBuf* me = peers[rank];
        const unsigned long long* f = &me->flags[par][lane];
        // bounded by TIME (5 s on the device's nanosecond timer, sampled every 1024 polls), not by a poll count
        // whose duration depends on the clock and on the NVLink round trip
        unsigned long long spins = 0, t0 = 0;
        while (ld_acquire_sys(f) != seq) {
            if ((++spins & 1023ull) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > kPeerTimeoutNs) { timedOut = true; break; }
            }
        }
#pragma unroll
        for (int i = 0; i < kNSums; ++i)
            v[i] = *reinterpret_cast<volatile double*>(&me->vals[par][lane][i]);
    }
    const bool anyTimeout = __any_sync(0xffffffffu, timedOut);
    if (lane == 0) {
        unsigned long long tw1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tw1));
        S->waitRedNs += tw1 - tw0;
        S->nRedWaits += 1u;
    }
    double tot[kNSums];
#pragma unroll
    for (int i = 0; i < kNSums; ++i) {
        double t = 0.0;
        for (int r = 0; r < nranks; ++r) {
            const double vr = __shfl_sync(0xffffffffu, v[i], r);
            t = (r == 0) ? vr : __dadd_rn(t, vr);      // ((v0+v1)+v2)+...
        }
        tot[i] = t;
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kNSums; ++i) S->gsums[i] = tot[i];
        if (anyTimeout) {
            S->nonfinite = 2;   // peer never arrived: reported as an error by the host
            S->done = 1;
        } else {
            scalar_step(step, S, S->gsums);
        }
    }
}

// ---- deterministic block reduction + last-block finish -----------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double (*sh)[kBlock / 32]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = warp_sum(v[i]);
        if (lane == 0) sh[i][w] = s;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = (lane < kBlock / 32) ? sh[i][lane] : 0.0;
            s = warp_sum(s);
            v[i] = s;   // valid in thread 0
        }
    }
    __syncthreads();
}

// Every thread of every block must call this (uniformly).
template <int NV>
__device__ __forceinline__ void reduce_finish(double (&v)[NV], const Reduce& R) {
    __shared__ double sh[NV][kBlock / 32];
    __shared__ bool amLast;
    block_sum<NV>(v, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) R.partials[i * kMaxGrid + blockIdx.x] = v[i];
        __threadfence();
        unsigned int t = atomicAdd(&R.S->ticket, 1u);
        amLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!amLast) return;
    __threadfence();
    double t[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += kBlock)
            s = __dadd_rn(s, __ldcg(&R.partials[i * kMaxGrid + b]));
        t[i] = s;
    }
    block_sum<NV>(t, sh);
    Scalars* S = R.S;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) S->acc[i] = __dadd_rn(S->acc[i], t[i]);
        if (R.step != STEP_NONE) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                S->sums[i] = S->acc[i];
                S->acc[i] = 0.0;
            }
#pragma unroll
            for (int i = NV; i < kNSums; ++i) S->sums[i] = 0.0;
            if (S->nranks == 1) scalar_step(R.step, S, S->sums);
        }
    }
    if (R.step != STEP_NONE && R.peers != nullptr) {
        // cross-rank sum over NVLink peer memory + scalar step, by the first warp of this (last) block
        __syncthreads();
        if (threadIdx.x < 32) peer_allreduce_step(S, R.peers, R.rank, S->nranks, R.step);
    }
    if (threadIdx.x == 0) {
        S->ticket = 0u;
        __threadfence();
    }
}

// ---- gathers / scatters between natural and internal order -------------------------------
__global__ void k_gather(int n, const int* __restrict__ perm, const double* __restrict__ src,
                         double* __restrict__ dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        dst[i] = perm ? __ldg(&src[perm[i]]) : src[i];
}
__global__ void k_scatter(int n, const int* __restrict__ perm, const double* __restrict__ src,
                          double* __restrict__ dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (perm) dst[perm[i]] = src[i];
        else dst[i] = src[i];
    }
}
// matrix value fill: sliced-ELL value of entry e is upper[faceOf[e]] (0 for padding)
__global__ void k_fill_values(int64_t nEntries, const int* __restrict__ faceOf,
                              const double* __restrict__ upper, double* __restrict__ val) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nEntries;
         e += (int64_t)gridDim.x * blockDim.x) {
        int f = faceOf[e];
        val[e] = f >= 0 ? __ldg(&upper[f]) : 0.0;
    }
}

// ---- column of a full-row ELL entry ----------------------------------------------------------
// 32-bit column, or (plan.hpp) per-(slice, j) base + 16-bit offset: entry e = sliceBase + 32 j + lane,
// so e >> 5 indexes the base table.  The kernels that are bound by the bytes of the full-row ELL read
// 10 instead of 12 bytes per entry.  The 16-bit form is used only when EVERY slice entry of the plan
// fits: two independent loads and an add, no branch -- a per-entry fall-back to the 32-bit column was
// measured 50 % SLOWER than plain 32-bit columns (the test on the base puts a fourth dependent load
// and a branch into the gather chain).
struct EllCols {
    const int* col;             // [nEntries]
    const uint16_t* col16;      // [nEntries] or nullptr
    const int* colBase;         // [nEntries/32]
};
template <bool C16>
__device__ __forceinline__ int ell_col(const EllCols& E, int64_t e) {
    if (!C16) return __ldg(&E.col[e]);      // (read-only path: the struct members carry no __restrict__)
    return __ldg(&E.colBase[e >> 5]) + (int)__ldg(&E.col16[e]);
}

// ---- processor-patch halo exchange over NVLink peer memory ---------------------------------------
// Replaces initMatrixInterfaces / updateMatrixInterfaces' send + receive (OF-dev processorFvPatchField.C,
// UIPstream/UOPstream) -- and the ncclSend/ncclRecv pair this library used before -- by direct stores: every
// rank owns a receive buffer  double vals[2][nSlots]; unsigned long long flags[2][kMaxRanks]  in its HBM, mapped
// into its neighbours with CUDA IPC.  The PACK kernel of the sender gathers x[faceCells] and stores each value
// straight into the neighbour's buffer (the two sides of a processor patch list their faces in the same order,
// so patch face j of the sender is slot j of the matching patch of the receiver); its last block then raises
// the sender's flag in every neighbour.  The CONSUMER (interface fix-up of the Amul, halo term of the Eisenstat
// sweeps) waits for its neighbours' flags at its start.  No NCCL kernel is involved: an NCCL send/recv kernel
// (640 threads, ~60 K registers per CTA) cannot become resident beside the persistent Amul grid and was running
// AFTER it -- ~40 us of exposed exchange per iteration on 2 GPUs (profiles/r02_*): the pack kernel can.
// NCCL send/recv stays as the fall-back when peer mapping is unavailable (B200PCG_HALO=nccl forces it).
struct Halo {
    double* const* dst;                   // [2][nSlots] address of slot i's value in the neighbour's vals[par]
    unsigned long long* const* nbrFlag;   // [2][nNbr]   address of flags[par][my rank] in neighbour k's buffer
    const unsigned long long* localFlags; // this rank's flags[2][kMaxRanks]           (nullptr: NCCL mode)
    const double* localVals;              // this rank's vals[2][nSlots]
    const int* nbrRanks;                  // [nNbr]
    int nSlots, nNbr;
};

__global__ void __launch_bounds__(kBlock)
k_pack_p2p(Halo H, const int* __restrict__ slotRow, const double* __restrict__ x, Scalars* S) {
    if (S->done) return;
    __shared__ bool amLast;
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(&S->haloSeq) + 1ull;
    const int par = (int)(seq & 1ull);
    double* const* dst = H.dst + (size_t)par * H.nSlots;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H.nSlots; i += gridDim.x * blockDim.x)
        *dst[i] = __ldg(&x[slotRow[i]]);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();     // cumulative over the block's stores (ordered before it by the barrier)
        const unsigned int t = atomicAdd(&S->haloTicket, 1u);
        amLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!amLast) return;
    if (threadIdx.x < H.nNbr) {
        __threadfence_system();     // after the ticket that proves every block's fence has been executed
        st_release_sys(H.nbrFlag[par * H.nNbr + threadIdx.x], seq);
    }
    if (threadIdx.x == 0) {
        S->haloSeq = seq;
        S->haloTicket = 0u;
        __threadfence();
    }
}

// Every thread of every block of a consumer kernel calls this once (uniformly), after the kernel's `done` test.
// Returns the buffer that holds the neighbours' values of the current exchange.
__device__ __forceinline__ const double* halo_acquire(const Halo& H, const double* ncclRecv, Scalars* S) {
    if (H.localFlags == nullptr) return ncclRecv;
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(&S->haloSeq);
    const int par = (int)(seq & 1ull);
    unsigned long long tw0 = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tw0));
    if (threadIdx.x < H.nNbr) {
        const unsigned long long* f = H.localFlags + par * kMaxRanks + H.nbrRanks[threadIdx.x];
        unsigned long long spins = 0, t0 = 0;
        while (ld_acquire_sys(f) < seq) {
            if ((++spins & 1023ull) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > kPeerTimeoutNs) { S->nonfinite = 2; break; }   // reported by the host
            }
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long tw1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tw1));
        S->waitHaloNs += tw1 - tw0;
        S->nHaloWaits += 1u;
    }
    return H.localVals + (size_t)par * H.nSlots;
}

// The PRODUCER side fused into the tail of the kernel that writes the exchanged vector (k_p writes pA): every CTA
// stores the patch-face values of the rows IT has just written (list built on the host for the launch's grid)
// straight into the neighbours' receive buffers, then the last CTA raises the flags -- the same protocol as
// k_pack_p2p without its launch.  Every thread of the CTA calls this after a __syncthreads that follows the CTA's
// last write of x.
struct PackTail {
    const int* ctaSStart;   // [grid + 1]; nullptr: no fused pack
    const int* ctaS;        // [nSlots] slots of each CTA's rows
    const int* slotRow;
    Halo H;
};
__device__ __forceinline__ void pack_tail(const PackTail& T, const double* x, Scalars* S) {
    __shared__ bool amLastPack;
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(&S->haloSeq) + 1ull;
    const int par = (int)(seq & 1ull);
    double* const* dst = T.H.dst + (size_t)par * T.H.nSlots;
    const int i1 = T.ctaSStart[blockIdx.x + 1];
    for (int i = T.ctaSStart[blockIdx.x] + (int)threadIdx.x; i < i1; i += kBlock) {
        const int slot = T.ctaS[i];
        *dst[slot] = x[T.slotRow[slot]];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int t = atomicAdd(&S->haloTicket, 1u);
        amLastPack = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!amLastPack) return;
    if (threadIdx.x < T.H.nNbr) {
        __threadfence_system();
        st_release_sys(T.H.nbrFlag[par * T.H.nNbr + threadIdx.x], seq);
    }
    if (threadIdx.x == 0) {
        S->haloSeq = seq;
        S->haloTicket = 0u;
        __threadfence();
    }
}

// Interface fix-up fused into the tail of the Amul (peer-memory halos only).  The separate k_iface_fix launch --
// 25 us of its own at 4 GPUs, before any waiting (per-rank wait accounting, profiles/r02_*) -- disappears: after
// its chunk loop every CTA corrects the interface rows of the chunks IT computed (their y values were written by
// this CTA: a __syncthreads away), adds the correction of (y, x) to its own partial dot product, and the kernel's
// one reduce_finish performs the cross-rank reduction and the scalar step.  The rows of a CTA come from a list
// built on the host for the launch's grid (chunk c belongs to CTA c mod grid).  Same per-row operation order as
// k_iface_fix: slots in (patch, face) order.
struct IfaceTail {
    const int* ctaBStart;   // [grid + 1]; nullptr: no fused tail
    const int* ctaB;        // [nBRows] interface-row indices of each CTA, ascending
    const int* bRow;
    const int* bStart;
    const int* bSlot;
    const double* bou;
    Halo H;
};
// every thread of the CTA, after a __syncthreads that follows the CTA's last write of y
template <bool DOT>
__device__ __forceinline__ void iface_tail(const IfaceTail& T, const double* __restrict__ x, double* y, Scalars* S,
                                           double& dot) {
    // the neighbours' values were stored into this rank's receive buffer while the rows above were computed
    const double* recv = halo_acquire(T.H, nullptr, S);
    const int i1 = T.ctaBStart[blockIdx.x + 1];
    for (int i = T.ctaBStart[blockIdx.x] + (int)threadIdx.x; i < i1; i += kBlock) {
        const int b = T.ctaB[i];
        const int r = T.bRow[b];
        const double y0 = y[r];
        double acc = y0;
        for (int e = T.bStart[b]; e < T.bStart[b + 1]; ++e) {
            const int slot = T.bSlot[e];
            acc = __dadd_rn(acc, -__dmul_rn(T.bou[slot], __ldcg(&recv[slot])));
        }
        y[r] = acc;
        if (DOT) dot = __dadd_rn(dot, __dmul_rn(__dadd_rn(acc, -y0), x[r]));
    }
}

// ---- lduMatrix::Amul / sumA (OF-dev lduMatrixATmul.C) ---------------------
// One row per thread, rows of a warp = one ELL slice -> the j-th loads of a warp are
// contiguous.  Row sum order: diag*x, then faces in ascending face order with the row as
// neighbour, then as owner == the order in which OpenFOAM's face loop updates Apsi[row].
// INIT additionally forms sumA (same loop, x == 1).  DOT accumulates (y, x).
template <bool INIT, bool DOT, bool C16>
__global__ void __launch_bounds__(kBlock)
k_spmv(int N, const int64_t* __restrict__ sliceBase, const uint32_t* __restrict__ rowLen,
       EllCols E, const double* __restrict__ val,
       const double* __restrict__ diag, const double* __restrict__ x, double* __restrict__ y,
       double* __restrict__ sA, Reduce R, IfaceTail T) {
    if (R.S->done) return;
    double dot[1] = {0.0};
    const int lane = threadIdx.x & 31;
    const int nSlices = (N + 31) >> 5;
    const int warpsPerGrid = (gridDim.x * kBlock) >> 5;
    for (int s = (blockIdx.x * kBlock + threadIdx.x) >> 5; s < nSlices; s += warpsPerGrid) {
        const int r = (s << 5) + lane;
        if (r < N) {
            const int64_t base = sliceBase[s] + lane;
            const int n = (int)(rowLen[r] >> 16);
            const double xr = x[r];
            const double d = diag[r];
            double acc = __dmul_rn(d, xr);
            double sa = d;
            int j = 0;
            for (; j + 4 <= n; j += 4) {
                const int64_t e = base + 32 * (int64_t)j;
                const int c0 = ell_col<C16>(E, e), c1 = ell_col<C16>(E, e + 32), c2 = ell_col<C16>(E, e + 64),
                          c3 = ell_col<C16>(E, e + 96);
                const double a0 = val[e], a1 = val[e + 32], a2 = val[e + 64], a3 = val[e + 96];
                const double x0 = __ldg(&x[c0]), x1 = __ldg(&x[c1]), x2 = __ldg(&x[c2]),
                             x3 = __ldg(&x[c3]);
                acc = __dadd_rn(acc, __dmul_rn(a0, x0));
                acc = __dadd_rn(acc, __dmul_rn(a1, x1));
                acc = __dadd_rn(acc, __dmul_rn(a2, x2));
                acc = __dadd_rn(acc, __dmul_rn(a3, x3));
                if (INIT) {
                    sa = __dadd_rn(sa, a0);
                    sa = __dadd_rn(sa, a1);
                    sa = __dadd_rn(sa, a2);
                    sa = __dadd_rn(sa, a3);
                }
            }
            for (; j < n; ++j) {
                const int64_t e = base + 32 * (int64_t)j;
                const int c0 = ell_col<C16>(E, e);
                const double a0 = val[e];
                acc = __dadd_rn(acc, __dmul_rn(a0, __ldg(&x[c0])));
                if (INIT) sa = __dadd_rn(sa, a0);
            }
            y[r] = acc;
            if (INIT) sA[r] = sa;
            if (DOT) dot[0] = __dadd_rn(dot[0], __dmul_rn(acc, xr));
        }
    }
    if (T.ctaBStart != nullptr) {
        __syncthreads();       // rows of this CTA's slices were written by several of its warps
        iface_tail<DOT>(T, x, y, R.S, dot[0]);
    }
    if (DOT) reduce_finish<1>(dot, R);
}

// Symmetric single-read Amul (SymPlan in plan.hpp): each coefficient is streamed from HBM once.
// Upper entries: coalesced sliced-ELL (value, column).  Lower entries: 32-bit references
// (owner row << 5 | q) -> the value is re-read from the owner's q-th upper entry (an L2 hit on
// banded cell orders; rows of a warp reference consecutive owners, so the re-read is coalesced
// too).  Row-sum order == OpenFOAM's face order: bit-identical to the CPU loop.
// Loads are issued in three dependent levels, each as one batch of up to B independent loads per
// thread (references + upper (col, value); then x gathers + referenced values; then the FP chain),
// so a warp keeps ~4B loads in flight instead of one dependent chain at a time.
constexpr int kSymBatch = 4;
template <bool INIT, bool DOT>
__global__ void __launch_bounds__(kBlock, 4)
k_spmv_sym(int N, int WU, int WL, const uint32_t* __restrict__ rowLen,
           const int* __restrict__ uCol, const double* __restrict__ uVal,
           const uint32_t* __restrict__ lRef, const double* __restrict__ diag,
           const double* __restrict__ x, double* __restrict__ y, double* __restrict__ sA, Reduce R) {
    if (R.S->done) return;
    constexpr int B = kSymBatch;
    double dot[1] = {0.0};
    const uint32_t lane = threadIdx.x & 31;
    const int nSlices = (N + 31) >> 5;
    const int warpsPerGrid = (gridDim.x * kBlock) >> 5;
    const uint32_t strideU = 32u * (uint32_t)WU, strideL = 32u * (uint32_t)WL;
    for (int s = (blockIdx.x * kBlock + threadIdx.x) >> 5; s < nSlices; s += warpsPerGrid) {
        const int r = (s << 5) + (int)lane;
        if (r < N) {
            const uint32_t len = rowLen[r];
            const int nL = (int)(len & 0xffffu), nU = (int)(len >> 16) - nL;
            const uint32_t lb = (uint32_t)s * strideL + lane, ub = (uint32_t)s * strideU + lane;
            // level 1
            uint32_t pk[B];
            int uc[B];
            double uv[B];
#pragma unroll
            for (int k = 0; k < B; ++k) pk[k] = (k < nL) ? lRef[lb + 32u * k] : 0u;
#pragma unroll
            for (int k = 0; k < B; ++k) {
                uc[k] = (k < nU) ? uCol[ub + 32u * k] : r;
                uv[k] = (k < nU) ? uVal[ub + 32u * k] : 0.0;
            }
            const double xr = x[r];
            const double d = diag[r];
            // level 2
            double lv[B], lx[B], ux[B];
#pragma unroll
            for (int k = 0; k < B; ++k) {
                const uint32_t a = pk[k] >> 5;
                const uint32_t pos = (a >> 5) * strideU + ((pk[k] & 31u) << 5) + (a & 31u);
                lv[k] = (k < nL) ? uVal[pos] : 0.0;
                lx[k] = (k < nL) ? __ldg(&x[a]) : 0.0;
            }
#pragma unroll
            for (int k = 0; k < B; ++k) ux[k] = (k < nU) ? __ldg(&x[uc[k]]) : 0.0;
            // level 3: row sum in OpenFOAM's face order (lower faces ascending, then upper)
            double acc = __dmul_rn(d, xr);
            double sa = d;
#pragma unroll
            for (int k = 0; k < B; ++k)
                if (k < nL) {
                    acc = __dadd_rn(acc, __dmul_rn(lv[k], lx[k]));
                    if (INIT) sa = __dadd_rn(sa, lv[k]);
                }
            for (int j = B; j < nL; ++j) {
                const uint32_t p0 = lRef[lb + 32u * j];
                const uint32_t a = p0 >> 5;
                const double v0 = uVal[(a >> 5) * strideU + ((p0 & 31u) << 5) + (a & 31u)];
                acc = __dadd_rn(acc, __dmul_rn(v0, __ldg(&x[a])));
                if (INIT) sa = __dadd_rn(sa, v0);
            }
#pragma unroll
            for (int k = 0; k < B; ++k)
                if (k < nU) {
                    acc = __dadd_rn(acc, __dmul_rn(uv[k], ux[k]));
                    if (INIT) sa = __dadd_rn(sa, uv[k]);
                }
            for (int j = B; j < nU; ++j) {
                const double v0 = uVal[ub + 32u * j];
                acc = __dadd_rn(acc, __dmul_rn(v0, __ldg(&x[uCol[ub + 32u * j]])));
                if (INIT) sa = __dadd_rn(sa, v0);
            }
            y[r] = acc;
            if (INIT) sA[r] = sa;
            if (DOT) dot[0] = __dadd_rn(dot[0], __dmul_rn(acc, xr));
        }
    }
    if (DOT) reduce_finish<1>(dot, R);
}

// ---- single-read Amul for RENUMBERED natural plans (plan.hpp SrPlan) ----------------------------
// One list of 32-bit words per row, in ascending natural face order (= the order of OpenFOAM's face loop, so
// the row sum is bit-identical on any row order): column << 5 | q.  q == 31: the row owns the coefficient (its
// next own value, coalesced sliced ELL); q < 31: the q-th own value of row `column`, re-read through L2.  Every
// coefficient is streamed from HBM once: 8 F + 8 F + 28 N bytes instead of the full-row ELL's 24 F + 28 N.
// Batches of 4 entries: all words, then all positions (one cached ownBase lookup each), then values + gathers,
// then the adds in order.
template <bool INIT, bool DOT>
__global__ void __launch_bounds__(kBlock)
k_spmv_sr(int N, const int64_t* __restrict__ sliceBase, const uint32_t* __restrict__ rowLen,
          const uint32_t* __restrict__ meta, const int64_t* __restrict__ ownBase, const double* __restrict__ ownVal,
          const double* __restrict__ diag, const double* __restrict__ x, double* __restrict__ y,
          double* __restrict__ sA, Reduce R) {
    if (R.S->done) return;
    double dot[1] = {0.0};
    const int lane = threadIdx.x & 31;
    const int nSlices = (N + 31) >> 5;
    const int warpsPerGrid = (gridDim.x * kBlock) >> 5;
    for (int s = (blockIdx.x * kBlock + threadIdx.x) >> 5; s < nSlices; s += warpsPerGrid) {
        const int r = (s << 5) + lane;
        if (r < N) {
            const int64_t base = sliceBase[s] + lane;
            const int n = (int)(rowLen[r] >> 16);
            const double xr = x[r];
            const double d = diag[r];
            double acc = __dmul_rn(d, xr);
            double sa = d;
            int jown = 0;
            int j = 0;
            for (; j + 4 <= n; j += 4) {
                const int64_t e = base + 32 * (int64_t)j;
                uint32_t m[4];
                int64_t pos[4];
                double v[4], g[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) m[k] = meta[e + 32 * k];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t a = m[k] >> 5, q = m[k] & 31u;
                    const bool own = (q == 31u);
                    const int64_t ob = __ldg(&ownBase[own ? (uint32_t)s : (a >> 5)]);
                    pos[k] = ob + 32 * (int64_t)(own ? jown : (int)q) + (own ? lane : (int)(a & 31u));
                    jown += own ? 1 : 0;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    v[k] = ownVal[pos[k]];
                    g[k] = __ldg(&x[m[k] >> 5]);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    acc = __dadd_rn(acc, __dmul_rn(v[k], g[k]));
                    if (INIT) sa = __dadd_rn(sa, v[k]);
                }
            }
            for (; j < n; ++j) {
                const uint32_t m0 = meta[base + 32 * (int64_t)j];
                const uint32_t a = m0 >> 5, q = m0 & 31u;
                const bool own = (q == 31u);
                const int64_t ob = __ldg(&ownBase[own ? (uint32_t)s : (a >> 5)]);
                const double v0 = ownVal[ob + 32 * (int64_t)(own ? jown : (int)q) + (own ? lane : (int)(a & 31u))];
                jown += own ? 1 : 0;
                acc = __dadd_rn(acc, __dmul_rn(v0, __ldg(&x[a])));
                if (INIT) sa = __dadd_rn(sa, v0);
            }
            y[r] = acc;
            if (INIT) sA[r] = sa;
            if (DOT) dot[0] = __dadd_rn(dot[0], __dmul_rn(acc, xr));
        }
    }
    if (DOT) reduce_finish<1>(dot, R);
}

// ---- TMA-staged variant of the symmetric Amul -------------------------------------------------
// Same arithmetic and row order as k_spmv_sym, but the streaming operands of a chunk of 256 rows
// (row lengths, lower references, upper columns + values, x, diag: all contiguous in the sliced
// layout) are brought into shared memory by the bulk-copy engine (cp.async.bulk + mbarrier
// complete_tx), kSymStages chunks ahead of the math.  The DRAM latency of the first load level is
// thereby taken off the warps' dependent chain: a warp only waits for the L2-served gathers
// (referenced values, x of neighbours).  One elected thread issues the copies; all 256 threads
// consume; one __syncthreads per chunk recycles the stage.
constexpr int kSymStagesMax = 4;
constexpr int kChunkRows = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// bytes of one stage for given widths (host + device)
// (row lengths travel as one byte per row, nL | nU << 4: 4 % less Amul traffic than the 32-bit form)
__host__ __device__ inline size_t sym_stage_bytes(int WU, int WL) {
    return (size_t)kChunkRows * ((size_t)WU * 12 + (size_t)WL * 4 + 1 + 16);
}

// BL/BU: number of lower/upper entries handled by the unrolled, batched-load path.  TAIL = false
// asserts WL <= BL and WU <= BU (every row fits the unrolled path), which removes the tail loops
// and the unused predicated slots: the kernel is ~50 % issue-bound (ncu: 237 instructions per
// 32 rows with BL = BU = 4 on the 3+3-entry hex rows), so instructions matter as much as bytes.
template <bool DOT, int kSymStages, int BL, int BU, bool TAIL>
__global__ void __launch_bounds__(kBlock)
k_spmv_sym_tma(int N, int WU, int WL, const uint8_t* __restrict__ rowLen8,
               const int* __restrict__ uCol, const double* __restrict__ uVal,
               const uint32_t* __restrict__ lRef, const double* __restrict__ diag,
               const double* __restrict__ x, double* __restrict__ y, Reduce R, IfaceTail T) {
    if (R.S->done) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t stageBytes = sym_stage_bytes(WU, WL);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);          // [kSymStages]
    unsigned char* stage0 = smem_raw + 128;
    // per-stage layout: uVal | x | diag | uCol | lRef | rowLen  (sizes are multiples of 16 B)
    const size_t oVal = 0, oX = (size_t)kChunkRows * WU * 8, oDiag = oX + kChunkRows * 8,
                 oCol = oDiag + kChunkRows * 8, oRef = oCol + (size_t)kChunkRows * WU * 4,
                 oLen = oRef + (size_t)kChunkRows * WL * 4;
    const int nChunks = (N + kChunkRows - 1) / kChunkRows;
    const uint32_t strideU = 32u * (uint32_t)WU, strideL = 32u * (uint32_t)WL;

    if (tid == 0) {
        for (int s = 0; s < kSymStages; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int chunk, int s) {
        unsigned char* st = stage0 + (size_t)s * stageBytes;
        const size_t r0 = (size_t)chunk * kChunkRows;
        mbar_expect_tx(&bars[s], (uint32_t)stageBytes);
        bulk_g2s(st + oVal, uVal + r0 * WU, (uint32_t)(kChunkRows * WU * 8), &bars[s]);
        bulk_g2s(st + oX, x + r0, kChunkRows * 8, &bars[s]);
        bulk_g2s(st + oDiag, diag + r0, kChunkRows * 8, &bars[s]);
        bulk_g2s(st + oCol, uCol + r0 * WU, (uint32_t)(kChunkRows * WU * 4), &bars[s]);
        if (WL > 0) bulk_g2s(st + oRef, lRef + r0 * WL, (uint32_t)(kChunkRows * WL * 4), &bars[s]);
        bulk_g2s(st + oLen, rowLen8 + r0, kChunkRows, &bars[s]);
    };
    if (tid == 0)
        for (int s = 0; s < kSymStages; ++s) {
            const int c = blockIdx.x + s * gridDim.x;
            if (c < nChunks) issue(c, s);
        }

    double dot[1] = {0.0};
    int it = 0;
    for (int chunk = blockIdx.x; chunk < nChunks; chunk += gridDim.x, ++it) {
        const int s = it % kSymStages;
        mbar_wait(&bars[s], (uint32_t)((it / kSymStages) & 1));
        const unsigned char* st = stage0 + (size_t)s * stageBytes;
        const double* sVal = reinterpret_cast<const double*>(st + oVal);
        const double* sX = reinterpret_cast<const double*>(st + oX);
        const double* sDiag = reinterpret_cast<const double*>(st + oDiag);
        const int* sCol = reinterpret_cast<const int*>(st + oCol);
        const uint32_t* sRef = reinterpret_cast<const uint32_t*>(st + oRef);
        const uint8_t* sLen = st + oLen;
        const int r = chunk * kChunkRows + (int)tid;
        if (r < N) {
            const uint32_t len = sLen[tid];
            const int nL = (int)(len & 15u), nU = (int)(len >> 4);
            const uint32_t lb = warp * strideL + lane, ub = warp * strideU + lane;
            // gathers served by L2/L1 (the only exposed latency)
            double lv[BL > 0 ? BL : 1], lx[BL > 0 ? BL : 1], ux[BU > 0 ? BU : 1];
#pragma unroll
            for (int k = 0; k < BL; ++k) {
                const uint32_t pk = (k < nL) ? sRef[lb + 32u * k] : 0u;
                const uint32_t a = pk >> 5;
                const uint32_t pos = (a >> 5) * strideU + ((pk & 31u) << 5) + (a & 31u);
                lv[k] = (k < nL) ? uVal[pos] : 0.0;
                lx[k] = (k < nL) ? __ldg(&x[a]) : 0.0;
            }
#pragma unroll
            for (int k = 0; k < BU; ++k) ux[k] = (k < nU) ? __ldg(&x[sCol[ub + 32u * k]]) : 0.0;
            const double xr = sX[tid];
            double acc = __dmul_rn(sDiag[tid], xr);
#pragma unroll
            for (int k = 0; k < BL; ++k)
                if (k < nL) acc = __dadd_rn(acc, __dmul_rn(lv[k], lx[k]));
            if (TAIL)
                for (int j = BL; j < nL; ++j) {
                    const uint32_t p0 = sRef[lb + 32u * j];
                    const uint32_t a = p0 >> 5;
                    const double v0 = uVal[(a >> 5) * strideU + ((p0 & 31u) << 5) + (a & 31u)];
                    acc = __dadd_rn(acc, __dmul_rn(v0, __ldg(&x[a])));
                }
#pragma unroll
            for (int k = 0; k < BU; ++k)
                if (k < nU) acc = __dadd_rn(acc, __dmul_rn(sVal[ub + 32u * k], ux[k]));
            if (TAIL)
                for (int j = BU; j < nU; ++j)
                    acc = __dadd_rn(acc, __dmul_rn(sVal[ub + 32u * j], __ldg(&x[sCol[ub + 32u * j]])));
            y[r] = acc;
            if (DOT) dot[0] = __dadd_rn(dot[0], __dmul_rn(acc, xr));
        }
        __syncthreads();   // every thread is done with stage s
        if (tid == 0) {
            const int next = chunk + kSymStages * gridDim.x;
            if (next < nChunks) issue(next, s);
        }
    }
    if (T.ctaBStart != nullptr) iface_tail<DOT>(T, x, y, R.S, dot[0]);
    if (DOT) reduce_finish<1>(dot, R);
}

// (A shared-memory-WINDOW variant of this kernel -- x / value / column rings of the neighbouring chunks in shared
// memory, so that in-window neighbours cost no L2 traffic -- was built and measured in round 1: it removed the L2
// stalls (-19 % L2 sectors) but needed 1.7x the instructions and ran at 275 us against 205 us; the kernel is
// co-limited by instruction issue.  It is not part of the product any more: profiles/r01_v5_spmv_win_vs_tma.md.)

// ---- processor interfaces (OF-dev processorFvPatchField.C, lduMatrixUpdateMatrixInterfaces.C;
//      SURVEY.md A.4) ------------------------------------------------------------------------
__global__ void k_pack(int nSlots, const int* __restrict__ slotRow, const double* __restrict__ x,
                       double* __restrict__ sendbuf, const Scalars* S) {
    if (S->done) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nSlots; i += gridDim.x * blockDim.x)
        sendbuf[i] = __ldg(&x[slotRow[i]]);
}
// result[faceCells[i]] -= coeffs[i]*nbr[i]; one thread per distinct interface row, its slots
// in (patch, face) order: a sorted-segment reduction.  MODE 0: Amul fix-up (+ correction of
// the (y,x) dot product), MODE 1: sumA fix-up (nbr == 1).
template <int MODE, bool DOT>
__global__ void __launch_bounds__(kBlock)
k_iface_fix(int nBRows, const int* __restrict__ bRow, const int* __restrict__ bStart,
            const int* __restrict__ bSlot, const double* __restrict__ bou,
            const double* recvNccl, Halo H, const double* __restrict__ x,
            double* __restrict__ y, Reduce R) {
    if (R.S->done) return;
    const double* recv = (MODE == 0) ? halo_acquire(H, recvNccl, R.S) : recvNccl;
    double dot[1] = {0.0};
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nBRows; b += gridDim.x * blockDim.x) {
        const int r = bRow[b];
        const double y0 = y[r];
        double acc = y0;
        for (int e = bStart[b]; e < bStart[b + 1]; ++e) {
            const int slot = bSlot[e];
            if (MODE == 0) acc = __dadd_rn(acc, -__dmul_rn(bou[slot], __ldcg(&recv[slot])));
            else acc = __dadd_rn(acc, -bou[slot]);
        }
        y[r] = acc;
        if (DOT) dot[0] = __dadd_rn(dot[0], __dmul_rn(__dadd_rn(acc, -y0), x[r]));
    }
    if (DOT) reduce_finish<1>(dot, R);
}

// The interface update split in two so that only a short kernel is left on the critical path behind the Amul
// (the one-kernel form above costs ~21 us of its own at 2 GPUs, 25+ at 4: a chain of dependent loads, two block
// reductions and the cross-rank reduction, all serial behind the Amul):
//   k_iface_pre   COMM stream, concurrent with the Amul: waits for the neighbours' halo flags, forms the products
//                 prod[e] = coeffs*nbr of every patch face in the row CSR's order and the correction of the dot
//                 product  sum_rows x[r] * (-sum_e prod[e])  (needs x and the halo, not y), reduced over its own
//                 ticket / partials arena into S->acc2.
//   k_iface_apply compute stream, after the Amul: y[r] = ((y[r] - prod[e0]) - prod[e1]) ... -- the same operation
//                 order as updateMatrixInterfaces, so Amul stays bit-identical -- while block 0 adds acc2 to the
//                 Amul's running total and performs the cross-rank reduction + scalar step at once: it depends on
//                 no row of this kernel.
template <bool DOT>
__global__ void __launch_bounds__(kBlock)
k_iface_pre(int nBRows, const int* __restrict__ bRow, const int* __restrict__ bStart, const int* __restrict__ bSlot,
            const double* __restrict__ bou, const double* recvNccl, Halo H, const double* __restrict__ x,
            double* __restrict__ prod, double* __restrict__ partials2, Scalars* S) {
    if (S->done) return;
    const double* recv = halo_acquire(H, recvNccl, S);
    double dot[1] = {0.0};
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nBRows; b += gridDim.x * blockDim.x) {
        double h = 0.0;
        for (int e = bStart[b]; e < bStart[b + 1]; ++e) {
            const int slot = bSlot[e];
            const double p = __dmul_rn(bou[slot], __ldcg(&recv[slot]));
            prod[e] = p;
            h = __dadd_rn(h, p);
        }
        if (DOT) dot[0] = __dadd_rn(dot[0], -__dmul_rn(h, x[bRow[b]]));
    }
    if (!DOT) return;
    __shared__ double sh[1][kBlock / 32];
    __shared__ bool amLast;
    block_sum<1>(dot, sh);
    if (threadIdx.x == 0) {
        partials2[blockIdx.x] = dot[0];
        __threadfence();
        amLast = (atomicAdd(&S->ticket2, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!amLast) return;
    __threadfence();
    double t[1] = {0.0};
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += kBlock) t[0] = __dadd_rn(t[0], __ldcg(&partials2[b]));
    block_sum<1>(t, sh);
    if (threadIdx.x == 0) {
        S->acc2 = t[0];
        S->ticket2 = 0u;
        __threadfence();
    }
}

template <bool DOT>
__global__ void __launch_bounds__(kBlock)
k_iface_apply(int nBRows, const int* __restrict__ bRow, const int* __restrict__ bStart,
              const double* __restrict__ prod, double* __restrict__ y, Reduce R) {
    Scalars* S = R.S;
    if (S->done) return;
    if (DOT && blockIdx.x == 0) {
        // (y, x) = the Amul's running total + the interface correction; then exactly what reduce_finish does when a
        // reduction completes
        if (threadIdx.x == 0) {
            S->sums[0] = __dadd_rn(S->acc[0], S->acc2);
            S->acc[0] = 0.0;
            S->acc2 = 0.0;
#pragma unroll
            for (int i = 1; i < kNSums; ++i) S->sums[i] = 0.0;
            if (S->nranks == 1) scalar_step(R.step, S, S->sums);
        }
        if (R.peers != nullptr) {
            __syncthreads();
            if (threadIdx.x < 32) peer_allreduce_step(S, R.peers, R.rank, S->nranks, R.step);
        }
    }
    // block 0 joins the row loop after its reduction; the other blocks cover the rows meanwhile
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nBRows; b += gridDim.x * blockDim.x) {
        const int r = bRow[b];
        double acc = y[r];
        for (int e = bStart[b]; e < bStart[b + 1]; ++e) acc = __dadd_rn(acc, -prod[e]);
        y[r] = acc;
    }
}

// ---- vector kernels ------------------------------------------------------------------------
// All vectors are cudaMalloc'ed (256-byte aligned) -> double2 accesses are aligned.
#define B200_VEC_LOOP(N, body2, body1)                                                     \
    {                                                                                      \
        const int n2 = (N) >> 1;                                                           \
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2;                        \
             i += gridDim.x * blockDim.x) { body2 }                                        \
        if (((N) & 1) && blockIdx.x == 0 && threadIdx.x == 0) { const int i = (N)-1; body1 } \
    }

// gSum(psi) for gAverage (OF-dev FieldFunctions.C)
__global__ void __launch_bounds__(kBlock)
k_sum(int N, const double* __restrict__ a, Reduce R) {
    double s[1] = {0.0};
    B200_VEC_LOOP(N,
        { double2 v = reinterpret_cast<const double2*>(a)[i];
          s[0] = __dadd_rn(s[0], __dadd_rn(v.x, v.y)); },
        { s[0] = __dadd_rn(s[0], a[i]); })
    reduce_finish<1>(s, R);
}

// rA = source - wA; normFactor and initial residual sums (OF-dev lduMatrixSolver.C normFactor,
// PCG.C): sums[0] = sum(|wA - xRef*sumA| + |source - xRef*sumA|),
// sums[1] = sum |rA|
__global__ void __launch_bounds__(kBlock)
k_norm_resid(int N, const double* __restrict__ wA, const double* __restrict__ sA,
             const double* __restrict__ src, double* __restrict__ rA, Reduce R) {
    double s[2] = {0.0, 0.0};
    const double xRef = R.S->xRef;
#define B200_NR1(W, SA, B, ROUT)                                                   \
    {                                                                              \
        const double t = __dmul_rn((SA), xRef);                                    \
        s[0] = __dadd_rn(s[0], __dadd_rn(fabs(__dadd_rn((W), -t)), fabs(__dadd_rn((B), -t)))); \
        const double rr_ = __dadd_rn((B), -(W));                                   \
        s[1] = __dadd_rn(s[1], fabs(rr_));                                         \
        ROUT = rr_;                                                                \
    }
    B200_VEC_LOOP(N,
        { double2 w = reinterpret_cast<const double2*>(wA)[i];
          double2 a = reinterpret_cast<const double2*>(sA)[i];
          double2 b = reinterpret_cast<const double2*>(src)[i];
          double2 r;
          B200_NR1(w.x, a.x, b.x, r.x)
          B200_NR1(w.y, a.y, b.y, r.y)
          reinterpret_cast<double2*>(rA)[i] = r; },
        { double r; B200_NR1(wA[i], sA[i], src[i], r) rA[i] = r; })
#undef B200_NR1
    reduce_finish<2>(s, R);
}

__global__ void __launch_bounds__(kBlock)
k_recip(int N, const double* d, double* rD) {   // d may alias rD (in-place)
    B200_VEC_LOOP(N,
        { double2 v = reinterpret_cast<const double2*>(d)[i];
          v.x = __ddiv_rn(1.0, v.x); v.y = __ddiv_rn(1.0, v.y);
          reinterpret_cast<double2*>(rD)[i] = v; },
        { rD[i] = __ddiv_rn(1.0, d[i]); })
}

// diagonalPreconditioner::precondition fused with gSumProd(wA, rA)
// (OF-dev diagonalPreconditioner.C, PCG.C).  PRECOND=false: `none` -> (rA, rA) only.
template <bool PRECOND>
__global__ void __launch_bounds__(kBlock)
k_precond_dot(int N, const double* __restrict__ rD, const double* __restrict__ rA,
              double* __restrict__ wA, Reduce R) {
    if (R.S->done) return;
    double s[1] = {0.0};
    B200_VEC_LOOP(N,
        { double2 r = reinterpret_cast<const double2*>(rA)[i];
          double2 w = r;
          if (PRECOND) {
              double2 d = reinterpret_cast<const double2*>(rD)[i];
              w.x = __dmul_rn(d.x, r.x); w.y = __dmul_rn(d.y, r.y);
              reinterpret_cast<double2*>(wA)[i] = w;
          }
          s[0] = __dadd_rn(s[0], __dmul_rn(w.x, r.x));
          s[0] = __dadd_rn(s[0], __dmul_rn(w.y, r.y)); },
        { double r = rA[i]; double w = r;
          if (PRECOND) { w = __dmul_rn(rD[i], r); wA[i] = w; }
          s[0] = __dadd_rn(s[0], __dmul_rn(w, r)); })
    reduce_finish<1>(s, R);
}

// ---- fused PCG vector kernels ---------------------------------------------------------------
// The loop body of OF-dev PCG.C is regrouped so that every vector is touched as few times as
// possible, WITHOUT changing any element-wise operation or its rounding:
//   k_p : psi += alpha_prev*pA (deferred from the previous iteration)      [if nIter > 0]
//         pA  = z + beta*pA, z = rD*rA | rA | wA        (ZMODE 1 diagonal | 0 none | 2 DIC-class)
//   Amul: wA = A pA, wApA                                                  (k_spmv*)
//   k_r : rA -= alpha*wA; sum|rA|; and for ZMODE 0/1 the NEXT iteration's wArA = sum(z*rA)
//   k_psi_final: the last deferred psi update, once, after the loop.
// diagonal: 48N + 32N bytes of vector traffic per iteration instead of 24N + 24N + 48N.
#define B200_Z(R_, D_, Z_) (ZMODE == 0 ? (R_) : (ZMODE == 1 ? __dmul_rn((D_), (R_)) : (Z_)))
template <int ZMODE>
__global__ void __launch_bounds__(kBlock)
k_p(int N, double* __restrict__ psi, double* __restrict__ pA, const double* __restrict__ rA,
    const double* __restrict__ rD, const double* __restrict__ zbuf, Scalars* S, PackTail T) {
    if (S->done) return;
    const bool first = (S->nIter == 0);
    const double beta = S->beta, alpha = S->alpha;
    B200_VEC_LOOP(N,
        { double2 r = make_double2(0.0, 0.0); double2 d = r; double2 z = r;
          if (ZMODE != 2) r = reinterpret_cast<const double2*>(rA)[i];
          if (ZMODE == 1) d = reinterpret_cast<const double2*>(rD)[i];
          if (ZMODE == 2) z = reinterpret_cast<const double2*>(zbuf)[i];
          double2 p;
          p.x = B200_Z(r.x, d.x, z.x); p.y = B200_Z(r.y, d.y, z.y);
          if (!first) {
              const double2 po = reinterpret_cast<const double2*>(pA)[i];
              double2 x = reinterpret_cast<double2*>(psi)[i];
              x.x = __dadd_rn(x.x, __dmul_rn(alpha, po.x));
              x.y = __dadd_rn(x.y, __dmul_rn(alpha, po.y));
              reinterpret_cast<double2*>(psi)[i] = x;
              p.x = __dadd_rn(p.x, __dmul_rn(beta, po.x));
              p.y = __dadd_rn(p.y, __dmul_rn(beta, po.y));
          }
          reinterpret_cast<double2*>(pA)[i] = p; },
        { const double r1 = (ZMODE != 2) ? rA[i] : 0.0;
          const double d1 = (ZMODE == 1) ? rD[i] : 0.0;
          const double z1 = (ZMODE == 2) ? zbuf[i] : 0.0;
          double p = B200_Z(r1, d1, z1);
          if (!first) {
              const double po = pA[i];
              psi[i] = __dadd_rn(psi[i], __dmul_rn(alpha, po));
              p = __dadd_rn(p, __dmul_rn(beta, po));
          }
          pA[i] = p; })
    if (T.ctaSStart != nullptr) {
        __syncthreads();       // the rows of this CTA's list were written by its own threads
        pack_tail(T, pA, S);
    }
}
#undef B200_Z

// ZMODE 2 (DIC-class) with nFirst > 0 additionally runs the forward sweep of the FIRST colour (rows
// [0, nFirst): no earlier neighbours, wA = rD*rA) in place, saving a launch and a pass over rA/rD;
// wA (which held A*pA) is only overwritten after its own element has been consumed.
// Rows of the first colour: [segStart[t*C], segStart[t*C + 1]) in every tile t = row >> tileShift
// (one tile: tileShift = 31).  segStart == nullptr: no fusion.
struct FirstColour {
    const int* segStart;
    int C, tileShift;
    __device__ __forceinline__ bool has(int row) const {
        return segStart != nullptr && row < __ldg(&segStart[(row >> tileShift) * C + 1]);
    }
};

template <int ZMODE>
__global__ void __launch_bounds__(kBlock)
k_r(int N, double* __restrict__ rA, double* wA, const double* __restrict__ rD, FirstColour fc, Reduce R) {
    if (R.S->done) return;
    const double alpha = R.S->alpha;
    double s[2] = {0.0, 0.0};
    B200_VEC_LOOP(N,
        { double2 r = reinterpret_cast<double2*>(rA)[i];
          const double2 w = reinterpret_cast<const double2*>(wA)[i];
          r.x = __dadd_rn(r.x, -__dmul_rn(alpha, w.x));
          r.y = __dadd_rn(r.y, -__dmul_rn(alpha, w.y));
          reinterpret_cast<double2*>(rA)[i] = r;
          if (ZMODE == 2) {
              const bool f0 = fc.has(2 * i);
              const bool f1 = fc.has(2 * i + 1);
              if (f0 || f1) {
                  const double2 d = reinterpret_cast<const double2*>(rD)[i];
                  double2 z = w;
                  if (f0) z.x = __dmul_rn(d.x, r.x);
                  if (f1) z.y = __dmul_rn(d.y, r.y);
                  reinterpret_cast<double2*>(wA)[i] = z;
              }
          }
          s[0] = __dadd_rn(s[0], __dadd_rn(fabs(r.x), fabs(r.y)));
          if (ZMODE == 1) {
              const double2 d = reinterpret_cast<const double2*>(rD)[i];
              s[1] = __dadd_rn(s[1], __dmul_rn(__dmul_rn(d.x, r.x), r.x));
              s[1] = __dadd_rn(s[1], __dmul_rn(__dmul_rn(d.y, r.y), r.y));
          } else if (ZMODE == 0) {
              s[1] = __dadd_rn(s[1], __dmul_rn(r.x, r.x));
              s[1] = __dadd_rn(s[1], __dmul_rn(r.y, r.y));
          } },
        { const double r = __dadd_rn(rA[i], -__dmul_rn(alpha, wA[i]));
          rA[i] = r;
          if (ZMODE == 2 && fc.has(i)) wA[i] = __dmul_rn(rD[i], r);
          s[0] = __dadd_rn(s[0], fabs(r));
          if (ZMODE == 1) s[1] = __dadd_rn(s[1], __dmul_rn(__dmul_rn(rD[i], r), r));
          else if (ZMODE == 0) s[1] = __dadd_rn(s[1], __dmul_rn(r, r)); })
    reduce_finish<2>(s, R);
}

__global__ void __launch_bounds__(kBlock)
k_psi_final(int N, double* __restrict__ psi, const double* __restrict__ pA, const Scalars* S) {
    if (!S->pendingPsi) return;
    const double alpha = S->alpha;
    B200_VEC_LOOP(N,
        { double2 x = reinterpret_cast<double2*>(psi)[i];
          const double2 p = reinterpret_cast<const double2*>(pA)[i];
          x.x = __dadd_rn(x.x, __dmul_rn(alpha, p.x));
          x.y = __dadd_rn(x.y, __dmul_rn(alpha, p.y));
          reinterpret_cast<double2*>(psi)[i] = x; },
        { psi[i] = __dadd_rn(psi[i], __dmul_rn(alpha, pA[i])); })
}

// ---- DIC-class preconditioner (OF-dev DICPreconditioner.C) -----------------
// Rows are ordered colour-major; "lower" entries of a row = neighbours of an earlier colour
// (or an earlier dependency level in DIC-exact mode).  One launch per colour, rows of a colour
// are independent.  Operation order inside a row is OpenFOAM's:
//   calcReciprocalD:  rD[u] -= upper*upper/rD[l]         (faces ascending)
//   forward:          wA[u] -= rD[u]*upper*wA[l]         (faces ascending)
//   backward:         wA[l] -= rD[l]*upper*wA[u]         (faces descending)
// Rows of one colour: segment (tile t, colour c) = [segStart[t*C + c], segStart[t*C + c + 1]) for every
// tile (plan.hpp); work item w = (tile, block-in-segment), `bps` blocks share one segment.  With one
// tile and bps == gridDim.x this is the plain grid-stride loop over the colour's row range.
struct ColourRows {
    const int* segStart;
    int C, c, nTiles, bps;
};
#define B200_FOR_COLOUR_ROWS(CR, r)                                                                  \
    for (int w_ = blockIdx.x; w_ < (CR).nTiles * (CR).bps; w_ += gridDim.x)                          \
        for (int t_ = w_ / (CR).bps, e_ = (CR).segStart[t_ * (CR).C + (CR).c + 1],                    \
                 r = (CR).segStart[t_ * (CR).C + (CR).c] + (w_ - t_ * (CR).bps) * kBlock + (int)threadIdx.x; \
             r < e_; r += (CR).bps * kBlock)

template <bool C16>
__global__ void __launch_bounds__(kBlock)
k_dic_calc_rd(ColourRows cr, const int64_t* __restrict__ sliceBase,
              const uint32_t* __restrict__ rowLen, EllCols E,
              const double* __restrict__ val, const double* __restrict__ diag,
              double* __restrict__ rD) {
    B200_FOR_COLOUR_ROWS(cr, r) {
        const int64_t base = sliceBase[r >> 5] + (r & 31);
        const int nLower = (int)(rowLen[r] & 0xffffu);
        double d = diag[r];
        for (int j = 0; j < nLower; ++j) {
            const int64_t e = base + 32 * (int64_t)j;
            const double a = val[e];
            // rD of an earlier colour was written by an earlier launch: plain (coherent) load
            d = __dadd_rn(d, -__ddiv_rn(__dmul_rn(a, a), rD[ell_col<C16>(E, e)]));
        }
        rD[r] = d;
    }
}

// forward sweep over one colour.  DOT: this colour's wA is final (last colour) -> add (wA, rA).
template <bool DOT, bool C16>
__global__ void __launch_bounds__(kBlock)
k_dic_fwd(ColourRows cr, const int64_t* __restrict__ sliceBase,
          const uint32_t* __restrict__ rowLen, EllCols E,
          const double* __restrict__ val, const double* __restrict__ rD,
          const double* __restrict__ rA, double* wA, Reduce R) {
    if (R.S->done) return;
    double s[1] = {0.0};
    B200_FOR_COLOUR_ROWS(cr, r) {
        const int64_t base = sliceBase[r >> 5] + (r & 31);
        const int nLower = (int)(rowLen[r] & 0xffffu);
        const double d = rD[r];
        const double rr = rA[r];
        double w = __dmul_rn(d, rr);
        for (int j = 0; j < nLower; ++j) {
            const int64_t e = base + 32 * (int64_t)j;
            w = __dadd_rn(w, -__dmul_rn(__dmul_rn(d, val[e]), wA[ell_col<C16>(E, e)]));
        }
        wA[r] = w;
        if (DOT) s[0] = __dadd_rn(s[0], __dmul_rn(w, rr));
    }
    if (DOT) reduce_finish<1>(s, R);
}

// backward sweep over one colour; wA of this colour becomes final -> always add (wA, rA).
template <bool C16>
__global__ void __launch_bounds__(kBlock)
k_dic_bwd(ColourRows cr, const int64_t* __restrict__ sliceBase,
          const uint32_t* __restrict__ rowLen, EllCols E,
          const double* __restrict__ val, const double* __restrict__ rD,
          const double* __restrict__ rA, double* wA, Reduce R) {
    if (R.S->done) return;
    double s[1] = {0.0};
    B200_FOR_COLOUR_ROWS(cr, r) {
        const int64_t base = sliceBase[r >> 5] + (r & 31);
        const uint32_t len = rowLen[r];
        const int nLower = (int)(len & 0xffffu), nTotal = (int)(len >> 16);
        const double d = rD[r];
        double w = wA[r];
        for (int j = nTotal - 1; j >= nLower; --j) {
            const int64_t e = base + 32 * (int64_t)j;
            w = __dadd_rn(w, -__dmul_rn(__dmul_rn(d, val[e]), wA[ell_col<C16>(E, e)]));
        }
        wA[r] = w;
        s[0] = __dadd_rn(s[0], __dmul_rn(w, rA[r]));
    }
    reduce_finish<1>(s, R);
}

// ---- Eisenstat form of the DIC-class PCG (B200_PRECOND_DIC_MC_EIS) ----------------------------
// The DIC-class preconditioner is M = (D~ + L) D~^-1 (D~ + L^T): L = the matrix's OWN strictly-lower
// part in the colour-major elimination order, D~ = the DIC diagonal.  For exactly that shape
// Eisenstat's identity removes the separate Amul from the iteration.  The system is first scaled
// symmetrically so that the DIC diagonal becomes the identity (once per solve, in the plan's own copy
// of the coefficients):
//     s_i = 1/sqrt|D~_i|,  sigma = sign(D~) (+1: SPD p_rghEqn; -1: the un-negated ph_rghEqn)
//     A- = sigma S A S,  L-_ij = sigma s_i s_j L_ij,  D-_i = D_i / D~_i,  x = S x-,  r- = sigma S r
//     M- = (I + L-)(I + L-^T)
// PCG on A preconditioned by M is then PLAIN CG on  A^ = (I+L-)^-1 A- (I+L-^T)^-1  in the variables
//     r^ = (I+L-)^-1 r-,   p^ = (I+L-^T) p-,   A- = (I+L-) + (D- - 2I) + (I+L-^T) [+ B-: processor interfaces]
//     t  = (I+L-^T)^-1 p^                                   backward sweep (t = the scaled search direction)
//     w^ = A^ p^ = t + (I+L-)^-1 (p^ + (D- - 2) t [+ B- t])   forward sweep
// One iteration:  k_eis_p -> backward sweeps -> [halo exchange of t] -> forward sweeps (+ (p^, w^)) ->
// k_eis_r (+ rho = (r^, r^)).  Every matrix entry is read ONCE per sweep, no vector of the
// preconditioner (rD) is read at all, and the solution increment sum(alpha t) accumulates in xa
// (psi = psi0 + S xa after the loop).  Same iterates as the three-kernel DIC-class loop in exact
// arithmetic (same preconditioner, same Krylov space); the rounding differs, which the DIC-class parity
// bar (solution within 1e-8 relative L2 at the same residual tolerance) allows.
// OpenFOAM's convergence test needs |r|_1 of the TRUE residual r = sigma S^-1 (I+L-) r^, a third pass
// over L.  k_eis_res evaluates it only when the device-side predictor  cRatio * sqrt(rho)  (cRatio is
// re-calibrated at every evaluation) comes within kEisMargin of the threshold, and at least every
// kEisEvery iterations; the solve only ever stops on an evaluated true residual.
// Storage: rh = r^, ph = p^ (rows below lastStart), t (rows of the LAST colour, >= lastStart, have no
// later neighbours: t == p^ there, kept in t only), y (= (I+L-)^-1 rhs; last-colour rows, which no sweep
// gathers, hold w^ = t + y), sv = s, eb = D- - 2, xa.

// DIC pivots: how many are negative, how many unusable (zero / non-finite) -> STEP_EIS_SIGN
__global__ void __launch_bounds__(kBlock)
k_eis_sign(int N, const double* __restrict__ dT, Reduce R) {
    double s[2] = {0.0, 0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const double d = dT[i];
        if (d < 0.0) s[0] = __dadd_rn(s[0], 1.0);
        if (!(fabs(d) > 0.0 && fabs(d) < 1.7e308)) s[1] = __dadd_rn(s[1], 1.0);
    }
    reduce_finish<2>(s, R);
}

// once per solve: sv = 1/sqrt|D~| (in place of D~), eb = D/D~ - 2, r- = sigma s r (in place), xa = 0
__global__ void __launch_bounds__(kBlock)
k_eis_setup(int N, const double* __restrict__ diag, double* __restrict__ dTsv, double* __restrict__ eb,
            double* __restrict__ rh, double* __restrict__ xa, const Scalars* S) {
    if (S->done) return;
    const double sigma = S->sigma;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const double dt = dTsv[i];
        const double s = __ddiv_rn(1.0, __dsqrt_rn(fabs(dt)));
        dTsv[i] = s;
        eb[i] = __dadd_rn(__ddiv_rn(diag[i], dt), -2.0);
        rh[i] = __dmul_rn(__dmul_rn(sigma, s), rh[i]);
        xa[i] = 0.0;
    }
}

// once per solve: the plan's coefficient copy becomes L- (both triangles): val *= sigma * (s_row * s_col).
// s_row*s_col is formed first so that the two copies of a coefficient stay bit-identical.
template <bool C16>
__global__ void __launch_bounds__(kBlock)
k_eis_scale_vals(int N, const int64_t* __restrict__ sliceBase, const uint32_t* __restrict__ rowLen,
                 EllCols E, double* __restrict__ val, const double* __restrict__ sv, const Scalars* S) {
    if (S->done) return;
    const double sigma = S->sigma;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) {
        const int64_t base = sliceBase[r >> 5] + (r & 31);
        const int n = (int)(rowLen[r] >> 16);
        const double sr = sv[r];
        for (int j = 0; j < n; ++j) {
            const int64_t e = base + 32 * (int64_t)j;
            const double ss = __dmul_rn(sr, __ldg(&sv[ell_col<C16>(E, e)]));
            val[e] = __dmul_rn(val[e], __dmul_rn(sigma, ss));
        }
    }
}
// nranks > 1: interface coefficients of B- (recv = the neighbours' s on the patch faces)
__global__ void __launch_bounds__(kBlock)
k_eis_scale_bou(int nSlots, const int* __restrict__ slotRow, const double* __restrict__ sv,
                const double* recvNccl, Halo H, double* __restrict__ bou, Scalars* S) {
    if (S->done) return;
    const double* recv = halo_acquire(H, recvNccl, S);
    const double sigma = S->sigma;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nSlots; i += gridDim.x * blockDim.x) {
        const double ss = __dmul_rn(__ldg(&sv[slotRow[i]]), __ldcg(&recv[i]));
        bou[i] = __dmul_rn(bou[i], __dmul_rn(sigma, ss));
    }
}
// like k_pack, but not gated by `done` bookkeeping differences: s of the interface rows
__global__ void k_eis_pack_s(int nSlots, const int* __restrict__ slotRow, const double* __restrict__ sv,
                             double* __restrict__ sendbuf) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nSlots; i += gridDim.x * blockDim.x)
        sendbuf[i] = __ldg(&sv[slotRow[i]]);
}

// resident CTAs per SM the batched sweeps are compiled for (caps the registers at 65536 / (256 * n)):
// 64 registers hold a batch of 6 without spills, a batch of 8 needs 80
constexpr int eis_sweep_ctas(int B) { return B == 0 ? 1 : (B <= 6 ? 4 : 3); }
// w -= sum of val*x[col] over the n ELL entries j0 .. j0+n-1 of a row (DESC: from the last one down).
// B > 0: batches of B entries -- all (column, value) loads, then all gathers, then the adds in order -- so
// that a row costs three dependent memory round trips instead of one per group of the unrolled loop
// (the sweeps are latency-bound: long_scoreboard, profiles/r01_v8_ncu_eisenstat_hex.md).  B == 0: plain loop.
template <int B, bool C16, bool DESC>
__device__ __forceinline__ double eis_row_sub(const EllCols& E, const double* __restrict__ val, const double* x,
                                              int64_t base, int j0, int n, double w) {
    if (B == 0) {
        for (int k = 0; k < n; ++k) {
            const int j = DESC ? (j0 + n - 1 - k) : (j0 + k);
            const int64_t e = base + 32 * (int64_t)j;
            w = __dadd_rn(w, -__dmul_rn(val[e], x[ell_col<C16>(E, e)]));
        }
        return w;
    }
    constexpr int BB = B > 0 ? B : 1;
    for (int done = 0; done < n; done += BB) {
        // entry k of the batch sits at a compile-time offset (+-32 k) from the batch's first entry
        const int jf = DESC ? (j0 + n - 1 - done) : (j0 + done);
        const int64_t ef = base + 32 * (int64_t)jf;
        const double* __restrict__ vp = val + ef;
        const int* __restrict__ cp = E.col + ef;
        const uint16_t* __restrict__ cp16 = E.col16 + ef;
        const int* __restrict__ bp = E.colBase + (ef >> 5);
        constexpr int STEP = DESC ? -1 : 1;
        int c[BB];
        double a[BB], g[BB];
        const int m = n - done;
#pragma unroll
        for (int k = 0; k < BB; ++k)
            if (k < m) {
                c[k] = C16 ? (__ldg(bp + STEP * k) + (int)__ldg(cp16 + STEP * 32 * k)) : __ldg(cp + STEP * 32 * k);
                a[k] = vp[STEP * 32 * k];
            }
#pragma unroll
        for (int k = 0; k < BB; ++k)
            if (k < m) g[k] = x[c[k]];
#pragma unroll
        for (int k = 0; k < BB; ++k)
            if (k < m) w = __dadd_rn(w, -__dmul_rn(a[k], g[k]));
    }
    return w;
}

// once per solve, one launch per colour, in place: rh = (I+L-)^-1 r-   (rows [r0, r1) of the colour)
template <bool C16>
__global__ void __launch_bounds__(kBlock)
k_eis_init_fwd(int r0, int r1, const int64_t* __restrict__ sliceBase, const uint32_t* __restrict__ rowLen,
               EllCols E, const double* __restrict__ val, double* rh, const Scalars* S) {
    if (S->done) return;
    for (int r = r0 + blockIdx.x * kBlock + threadIdx.x; r < r1; r += gridDim.x * kBlock) {
        const int64_t base = sliceBase[r >> 5] + (r & 31);
        const int nLower = (int)(rowLen[r] & 0xffffu);
        rh[r] = eis_row_sub<0, C16, false>(E, val, rh, base, 0, nLower, rh[r]);
    }
}

// rho_0 = (r^, r^)
__global__ void __launch_bounds__(kBlock)
k_eis_rho0(int N, const double* __restrict__ rh, Reduce R) {
    if (R.S->done) return;
    double s[1] = {0.0};
    B200_VEC_LOOP(N,
        { const double2 r = reinterpret_cast<const double2*>(rh)[i];
          s[0] = __dadd_rn(s[0], __dmul_rn(r.x, r.x));
          s[0] = __dadd_rn(s[0], __dmul_rn(r.y, r.y)); },
        { s[0] = __dadd_rn(s[0], __dmul_rn(rh[i], rh[i])); })
    reduce_finish<1>(s, R);
}

// xa += alpha_prev * t_prev (deferred, as psi in k_p);  p^ = r^ + beta p^.  Rows of the last colour keep
// p^ in t (t == p^ there): nothing else to do for their backward sweep.
__device__ __forceinline__ void eis_p_elem(int i, bool last, bool first, double alpha, double beta,
                                           const double* __restrict__ rh, double* __restrict__ ph, double* t,
                                           double* __restrict__ xa) {
    double p = rh[i];
    if (!first) {
        const double tv = t[i];
        const double po = last ? tv : ph[i];
        xa[i] = __dadd_rn(xa[i], __dmul_rn(alpha, tv));
        p = __dadd_rn(p, __dmul_rn(beta, po));
    }
    if (last) t[i] = p;
    else ph[i] = p;
}
__global__ void __launch_bounds__(kBlock)
k_eis_p(int N, int lastStart, double* __restrict__ xa, double* __restrict__ ph, double* t,
        const double* __restrict__ rh, const Scalars* S) {
    if (S->done) return;
    const bool first = (S->nIter == 0);
    const double beta = S->beta;
    const double alpha = S->alpha;
    B200_VEC_LOOP(N,
        { const bool anyLast = (2 * i + 1 >= lastStart);
          const bool bothLast = (2 * i >= lastStart);
          if (anyLast && !bothLast) {     // the one pair that straddles the colour boundary
              eis_p_elem(2 * i, false, first, alpha, beta, rh, ph, t, xa);
              eis_p_elem(2 * i + 1, true, first, alpha, beta, rh, ph, t, xa);
          } else {
              double2 p = reinterpret_cast<const double2*>(rh)[i];
              if (!first) {
                  const double2 tv = reinterpret_cast<const double2*>(t)[i];
                  double2 po = tv;
                  if (!bothLast) po = reinterpret_cast<const double2*>(ph)[i];
                  double2 x = reinterpret_cast<double2*>(xa)[i];
                  x.x = __dadd_rn(x.x, __dmul_rn(alpha, tv.x));
                  x.y = __dadd_rn(x.y, __dmul_rn(alpha, tv.y));
                  reinterpret_cast<double2*>(xa)[i] = x;
                  p.x = __dadd_rn(p.x, __dmul_rn(beta, po.x));
                  p.y = __dadd_rn(p.y, __dmul_rn(beta, po.y));
              }
              if (bothLast) reinterpret_cast<double2*>(t)[i] = p;
              else reinterpret_cast<double2*>(ph)[i] = p;
          } },
        { eis_p_elem(i, i >= lastStart, first, alpha, beta, rh, ph, t, xa); })
}

// backward sweep over the rows [r0, r1) of one colour: t = p^ - L-^T t.
// FWD0 == 1 (first colour, single rank): its rows have no earlier neighbours and D- == 1 there, so their
// forward sweep y = p^ + (D- - 2) t = p^ - t and their share of (p^, w^) ride along.
// FWD0 == 2 (first colour, nranks > 1, halo overlapped): the same for the rows WITHOUT a processor face;
// the interface rows (rowB >= 0) were swept by k_eis_bwd_rows before the halo exchange started and get
// their forward part from k_eis_fwd_rows once the halo term has arrived -- they are left alone here, so
// this kernel runs concurrently with the exchange of t.
// B > 0: the next row's row length / slice base / p^ are requested before the current row's gathers.
template <int FWD0, bool C16, int B, int CT>
__global__ void __launch_bounds__(kBlock, CT)
k_eis_bwd(int r0, int r1, const int64_t* __restrict__ sliceBase, const uint32_t* __restrict__ rowLen,
          EllCols E, const double* __restrict__ val, const double* __restrict__ ph, double* t,
          double* __restrict__ y, const int* __restrict__ rowB, Reduce R) {
    if (R.S->done) return;
    double s[1] = {0.0};
    const int stride = gridDim.x * kBlock;
    int r = r0 + blockIdx.x * kBlock + threadIdx.x;
    uint32_t len = 0;
    int64_t sb = 0;
    double p = 0.0;
    int bi = -1;
#define B200_EIS_BWD_LOAD(ROW, LEN, SB, P, BI)                        \
    {                                                                 \
        LEN = rowLen[ROW]; SB = sliceBase[(ROW) >> 5]; P = ph[ROW];   \
        if (FWD0 == 2) BI = rowB[ROW];                                \
    }
    if (r < r1) B200_EIS_BWD_LOAD(r, len, sb, p, bi)
    while (r < r1) {
        const int rn = r + stride;
        uint32_t lenN = 0;
        int64_t sbN = 0;
        double pN = 0.0;
        int biN = -1;
        if (B > 0 && rn < r1) B200_EIS_BWD_LOAD(rn, lenN, sbN, pN, biN)
        if (FWD0 != 2 || bi < 0) {
            const int nLower = (int)(len & 0xffffu), nTotal = (int)(len >> 16);
            const double w = eis_row_sub<B, C16, true>(E, val, t, sb + (r & 31), nLower, nTotal - nLower, p);
            t[r] = w;
            if (FWD0 != 0) {
                const double yv = __dadd_rn(p, -w);
                y[r] = yv;
                s[0] = __dadd_rn(s[0], __dmul_rn(p, __dadd_rn(w, yv)));
            }
        }
        if (B == 0 && rn < r1) B200_EIS_BWD_LOAD(rn, lenN, sbN, pN, biN)
        r = rn; len = lenN; sb = sbN; p = pN; bi = biN;
    }
#undef B200_EIS_BWD_LOAD
    if (FWD0 != 0) reduce_finish<1>(s, R);
}


// nranks > 1, halo overlapped: backward sweep of the first colour's interface rows only (bRow[0 .. nB0):
// bRow is ascending, so the first colour's interface rows are a prefix), ahead of the bulk of the colour:
// once they are done t is complete on every interface row and the exchange can start.
template <bool C16>
__global__ void __launch_bounds__(kBlock)
k_eis_bwd_rows(int nB0, const int* __restrict__ bRow, const int64_t* __restrict__ sliceBase,
               const uint32_t* __restrict__ rowLen, EllCols E, const double* __restrict__ val,
               const double* __restrict__ ph, double* t, const Scalars* S) {
    if (S->done) return;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nB0; b += gridDim.x * blockDim.x) {
        const int r = bRow[b];
        const uint32_t len = rowLen[r];
        const int nLower = (int)(len & 0xffffu), nTotal = (int)(len >> 16);
        t[r] = eis_row_sub<0, C16, true>(E, val, t, sliceBase[r >> 5] + (r & 31), nLower, nTotal - nLower, ph[r]);
    }
}

// halo term of the forward right-hand side (nranks > 1): hb[b] = (B- t)[bRow[b]] = -sum bou*t_nbr over the
// row's processor faces in (patch, face) order (the sorted-segment form of updateMatrixInterfaces)
template <bool ROWS>
__global__ void __launch_bounds__(kBlock)
k_eis_halo(int nBRows, const int* __restrict__ bStart, const int* __restrict__ bSlot,
           const double* __restrict__ bou, const double* recvNccl, Halo H, double* __restrict__ hb,
           Scalars* S, int nB0, const int* __restrict__ bRow, const double* __restrict__ ph,
           const double* __restrict__ t, double* __restrict__ y, Reduce R) {
    if (S->done) return;
    const double* recv = halo_acquire(H, recvNccl, S);
    double s[1] = {0.0};
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nBRows; b += gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int e = bStart[b]; e < bStart[b + 1]; ++e) {
            const int slot = bSlot[e];
            acc = __dadd_rn(acc, -__dmul_rn(bou[slot], __ldcg(&recv[slot])));
        }
        hb[b] = acc;
        if (ROWS && b < nB0) {
            // the forward part of the first colour's interface rows (what k_eis_fwd_rows does), now that their
            // halo term is known: one launch less behind the exchange
            const int r = bRow[b];
            const double p = ph[r], tv = t[r];
            const double yv = __dadd_rn(__dadd_rn(p, -tv), acc);
            y[r] = yv;
            s[0] = __dadd_rn(s[0], __dmul_rn(p, __dadd_rn(tv, yv)));
        }
    }
    if (ROWS) reduce_finish<1>(s, R);
}

// forward sweep over the rows [r0, r1) of one colour: y = p^ + (D- - 2) t [+ B- t] - L- y; every colour
// adds its share of (p^, w^), w^ = t + y.  LAST: p^ == t lives in t; the rows are gathered by no sweep
// and store w^.
template <bool LAST, bool HALO, bool C16, int B, int CT>
__global__ void __launch_bounds__(kBlock, CT)
k_eis_fwd(int r0, int r1, const int64_t* __restrict__ sliceBase, const uint32_t* __restrict__ rowLen,
          EllCols E, const double* __restrict__ val, const double* __restrict__ ph,
          const double* __restrict__ eb, const double* __restrict__ t, double* y,
          const int* __restrict__ rowB, const double* __restrict__ hb, Reduce R) {
    if (R.S->done) return;
    double s[1] = {0.0};
    const int stride = gridDim.x * kBlock;
    int r = r0 + blockIdx.x * kBlock + threadIdx.x;
    uint32_t len = 0;
    int64_t sb = 0;
    double p = 0.0, tv = 0.0, ev = 0.0;
#define B200_EIS_FWD_LOAD(ROW, LEN, SB, P, TV, EV)                   \
    {                                                                \
        LEN = rowLen[ROW]; SB = sliceBase[(ROW) >> 5]; EV = eb[ROW]; \
        TV = t[ROW];                                                 \
        P = LAST ? TV : ph[ROW];                                     \
    }
    if (r < r1) B200_EIS_FWD_LOAD(r, len, sb, p, tv, ev)
    while (r < r1) {
        const int rn = r + stride;
        uint32_t lenN = 0;
        int64_t sbN = 0;
        double pN = 0.0, tvN = 0.0, evN = 0.0;
        if (B > 0 && rn < r1) B200_EIS_FWD_LOAD(rn, lenN, sbN, pN, tvN, evN)
        const int nLower = (int)(len & 0xffffu);
        double rhs = __dadd_rn(p, __dmul_rn(ev, tv));
        if (HALO) {
            const int b = rowB[r];
            if (b >= 0) rhs = __dadd_rn(rhs, hb[b]);
        }
        const double w = eis_row_sub<B, C16, false>(E, val, y, sb + (r & 31), 0, nLower, rhs);
        const double wh = __dadd_rn(tv, w);
        y[r] = LAST ? wh : w;
        s[0] = __dadd_rn(s[0], __dmul_rn(p, wh));
        if (B == 0 && rn < r1) B200_EIS_FWD_LOAD(rn, lenN, sbN, pN, tvN, evN)
        r = rn; len = lenN; sb = sbN; p = pN; tv = tvN; ev = evN;
    }
#undef B200_EIS_FWD_LOAD
    reduce_finish<1>(s, R);
}

// r^ -= alpha w^ (w^ = t + y below lastStart, stored as such from lastStart on); rho = (r^, r^)
__global__ void __launch_bounds__(kBlock)
k_eis_r(int N, int lastStart, double* __restrict__ rh, const double* __restrict__ y,
        const double* __restrict__ t, Reduce R) {
    if (R.S->done) return;
    const double alpha = R.S->alpha;
    double s[1] = {0.0};
    B200_VEC_LOOP(N,
        { double2 r = reinterpret_cast<double2*>(rh)[i];
          double2 w = reinterpret_cast<const double2*>(y)[i];
          if (2 * i < lastStart) {
              const double2 tv = reinterpret_cast<const double2*>(t)[i];
              w.x = __dadd_rn(tv.x, w.x);
              if (2 * i + 1 < lastStart) w.y = __dadd_rn(tv.y, w.y);
          }
          r.x = __dadd_rn(r.x, -__dmul_rn(alpha, w.x));
          r.y = __dadd_rn(r.y, -__dmul_rn(alpha, w.y));
          reinterpret_cast<double2*>(rh)[i] = r;
          s[0] = __dadd_rn(s[0], __dmul_rn(r.x, r.x));
          s[0] = __dadd_rn(s[0], __dmul_rn(r.y, r.y)); },
        { double w = y[i];
          if (i < lastStart) w = __dadd_rn(t[i], w);
          const double r = __dadd_rn(rh[i], -__dmul_rn(alpha, w));
          rh[i] = r;
          s[0] = __dadd_rn(s[0], __dmul_rn(r, r)); })
    reduce_finish<1>(s, R);
}

// true residual of the iteration, only when the scalar step asked for it: sum |(I + L-) r^| / s
template <bool C16, int B>
__global__ void __launch_bounds__(kBlock, eis_sweep_ctas(B))
k_eis_res(int N, const int64_t* __restrict__ sliceBase, const uint32_t* __restrict__ rowLen, EllCols E,
          const double* __restrict__ val, const double* __restrict__ sv, const double* __restrict__ rh,
          Reduce R) {
    if (R.S->done || !R.S->needCheck) return;
    double s[1] = {0.0};
    for (int r = blockIdx.x * kBlock + threadIdx.x; r < N; r += gridDim.x * kBlock) {
        const int64_t base = sliceBase[r >> 5] + (r & 31);
        const int nLower = (int)(rowLen[r] & 0xffffu);
        // rh + L- rh == -(-rh - L- rh): the sweeps' batched row helper (negation is exact)
        const double acc = -eis_row_sub<B, C16, false>(E, val, rh, base, 0, nLower, -rh[r]);
        s[0] = __dadd_rn(s[0], __ddiv_rn(fabs(acc), sv[r]));
    }
    reduce_finish<1>(s, R);
}

// after the loop: psi = psi0 + S (xa [+ the last, still deferred alpha*t])
__global__ void __launch_bounds__(kBlock)
k_eis_final(int N, double* __restrict__ psi, const double* __restrict__ xa, const double* __restrict__ t,
            const double* __restrict__ sv, const Scalars* S) {
    if (S->nIter == 0 && !S->pendingPsi) return;   // no loop body ran: psi is returned untouched
    const bool pend = S->pendingPsi != 0;
    const double alpha = S->alpha;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        double x = xa[i];
        if (pend) x = __dadd_rn(x, __dmul_rn(alpha, t[i]));
        psi[i] = __dadd_rn(psi[i], __dmul_rn(sv[i], x));
    }
}

// ---- smoothSolver: GaussSeidel / symGaussSeidel on (a)symmetric lduMatrices ---------
// Replaces GaussSeidelSmoother::smooth / symGaussSeidelSmoother::smooth and lduMatrix::residual (OF-dev
// GaussSeidelSmoother.C, symGaussSeidelSmoother.C, lduMatrixATmul.C), which the reference selects for U, Yi, h
// and k (cases/steckler/system/fvSolution:48-61).  Upstream's sweep is a sequential loop over the cells,
//     psi_c = (bPrime_c - sum_{owned faces} upper_f psi_u) / diag_c;   bPrime_u -= lower_f psi_c   (scatter)
// i.e. row c needs the NEW values of its lower neighbours and the OLD values of its upper ones.  In row form
// that is one formula for both directions,
//     psi_c = (b_c - sum_{lower nbrs} lower_f psi_l - sum_{upper nbrs} upper_f psi_u) / diag_c     in place,
// applied to the rows in an order that respects the dependencies: rows grouped by dependency LEVEL of the cell
// order (plan.hpp Ordering::Levels -- the forward sweep visits the levels ascending, the reverse sweep descending;
// during the reverse sweep a row's lower neighbours still hold their forward values, which is exactly what
// upstream's bPrime carries over) reproduces upstream BIT FOR BIT: same elimination order, and with the ELL
// entries stored [lower | upper], each in ascending face order, the same order of subtractions
// (tests/test_smooth_plan.py, numpy transliteration against the C oracle).  Grouped by COLOUR instead
// (Ordering::MultiColour) the same kernel is a multicolour Gauss-Seidel: a different ordering of the same
// smoother, one launch per colour.  Two launches per sweep are redundant and skipped, bit-neutrally: the last group
// of a forward sweep and the first group of the reverse sweep are the same rows with the same inputs (the reverse
// sweep starts at the group before it), and so are the last group of a reverse sweep (group 0) and the first group
// of the NEXT forward sweep -- after the first sweep of a solve a symGaussSeidel sweep is 2 C - 2 launches (two
// half-sweeps on a 2-colour hex mesh).

// matrix value fill of the full-row ELL for lower != upper: the entry of row r on face f carries upper[f] when
// r's cell owns the face (A[l][u] = upper), lower[f] when it is the face's neighbour (A[u][l] = lower)
// valT != nullptr (PBiCG): also the entries of the TRANSPOSED matrix, i.e. the other coefficient of the same face
__global__ void __launch_bounds__(kBlock)
k_fill_values_asym(int N, const int64_t* __restrict__ sliceBase, const uint32_t* __restrict__ rowLen,
                   const int* __restrict__ faceOf, const int* __restrict__ perm, const int* __restrict__ lowerAddr,
                   const double* __restrict__ upper, const double* __restrict__ lower, double* __restrict__ val,
                   double* __restrict__ valT) {
    for (int r = blockIdx.x * kBlock + threadIdx.x; r < N; r += gridDim.x * kBlock) {
        const int c = perm ? perm[r] : r;
        const int64_t base = sliceBase[r >> 5] + (r & 31);
        const int n = (int)(rowLen[r] >> 16);
        for (int j = 0; j < n; ++j) {
            const int64_t e = base + 32 * (int64_t)j;
            const int f = faceOf[e];
            const bool owns = (__ldg(&lowerAddr[f]) == c);
            const double up = __ldg(&upper[f]), lo = __ldg(&lower[f]);
            val[e] = owns ? up : lo;
            if (valT) valT[e] = owns ? lo : up;
        }
    }
}

// one group (level / colour) of a Gauss-Seidel sweep: rows [r0, r1), in place.  The next row's row length /
// slice base / b / diag are requested before the current row's gathers (as in k_eis_bwd).
// RES (multicolour mode, the LAST group an iteration updates): every neighbour of these rows already holds its final
// value of the iteration, so the rows' share of gSumMag(residual) is available here for free,
//     r_c = b_c - sum a x - diag_c psi_c = w - diag_c * fl(w / diag_c)          (a rounding-level quantity),
// and is accumulated into the running total that k_gs_resid completes over the other rows.  (Upstream evaluates
// the same quantity as (b - diag psi) - sum a x: the two differ in the last bits of an already rounding-level term.
// The level-scheduled mode does not use RES: its residual is evaluated in upstream's order over all rows.)
// HALO (nranks > 1): upstream treats a processor patch as an explicit, Jacobi-like contribution refreshed once per
// sweep (bPrime = source; updateMatrixInterfaces with the negated coefficients) -- the interface rows take their
// right-hand side from hbv (k_gs_bprime) instead of b.
// RES == 2 (two-colour plans, the FIRST group an iteration updates): nothing these rows read has changed since the
// end of the previous iteration, so  w - diag_c * psi_c(old)  IS their residual at the end of that iteration.  The
// pass completes the previous iteration's reduction (its other group contributed through RES == 1) and runs
// STEP_GS_RES -- no separate residual kernel at all -- while it writes the NEW values into a shadow array (xw != xo):
// if the step decides that the previous iteration converged, the old values are still there.
// x arrays: xg is gathered from (the OTHER group's rows), xo holds this group's old values (RES == 2 only), xw
// receives the new ones.
template <bool C16, int B, int CT, int RES, bool HALO>
__global__ void __launch_bounds__(kBlock, CT)
k_gs_rows(int r0, int r1, const int64_t* __restrict__ sliceBase, const uint32_t* __restrict__ rowLen,
          EllCols E, const double* __restrict__ val, const double* __restrict__ diag,
          const double* __restrict__ b, const int* __restrict__ rowB, const double* __restrict__ hbv,
          const double* xg, const double* xo, double* xw, Reduce R) {
    if (R.S->done) return;
    double s[1] = {0.0};
    const int stride = gridDim.x * kBlock;
    int r = r0 + blockIdx.x * kBlock + threadIdx.x;
    uint32_t len = 0;
    int64_t sb = 0;
    double bv = 0.0, dv = 1.0, ov = 0.0;
#define B200_GS_LOAD(ROW, LEN, SB, BV, DV, OV)                               \
    {                                                                        \
        LEN = rowLen[ROW]; SB = sliceBase[(ROW) >> 5]; DV = diag[ROW];       \
        int bi_ = -1;                                                        \
        if (HALO) bi_ = rowB[ROW];                                           \
        BV = (HALO && bi_ >= 0) ? hbv[bi_] : b[ROW];                         \
        if (RES == 2) OV = xo[ROW];                                          \
    }
    if (r < r1) B200_GS_LOAD(r, len, sb, bv, dv, ov)
    while (r < r1) {
        const int rn = r + stride;
        uint32_t lenN = 0;
        int64_t sbN = 0;
        double bvN = 0.0, dvN = 1.0, ovN = 0.0;
        if (B > 0 && rn < r1) B200_GS_LOAD(rn, lenN, sbN, bvN, dvN, ovN)
        const double w = eis_row_sub<B, C16, false>(E, val, xg, sb + (r & 31), 0, (int)(len >> 16), bv);
        const double xn = __ddiv_rn(w, dv);
        xw[r] = xn;
        if (RES == 1) s[0] = __dadd_rn(s[0], fabs(__dadd_rn(w, -__dmul_rn(dv, xn))));
        if (RES == 2) s[0] = __dadd_rn(s[0], fabs(__dadd_rn(w, -__dmul_rn(dv, ov))));
        if (B == 0 && rn < r1) B200_GS_LOAD(rn, lenN, sbN, bvN, dvN, ovN)
        r = rn; len = lenN; sb = sbN; bv = bvN; dv = dvN; ov = ovN;
    }
#undef B200_GS_LOAD
    if (RES != 0) reduce_finish<1>(s, R);
}

// nranks > 1, once per sweep: bPrime of the interface rows = source + sum bou*psi_nbr over the row's processor faces
// in (patch, face) order -- initMatrixInterfaces / updateMatrixInterfaces with mBouCoeffs = -interfaceBouCoeffs
// (bPrime[faceCells] -= (-bou)*psi_nbr; the negation is exact, so += bou*psi_nbr is the same bits)
__global__ void __launch_bounds__(kBlock)
k_gs_bprime(int nBRows, const int* __restrict__ bRow, const int* __restrict__ bStart, const int* __restrict__ bSlot,
            const double* __restrict__ bou, const double* recvNccl, Halo H, const double* __restrict__ b,
            double* __restrict__ hbv, Scalars* S) {
    if (S->done) return;
    const double* recv = halo_acquire(H, recvNccl, S);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nBRows; i += gridDim.x * blockDim.x) {
        double acc = b[bRow[i]];
        for (int e = bStart[i]; e < bStart[i + 1]; ++e) {
            const int slot = bSlot[e];
            acc = __dadd_rn(acc, __dmul_rn(bou[slot], __ldcg(&recv[slot])));
        }
        hbv[i] = acc;
    }
}

// gSumMag(lduMatrix::residual(psi, source)) over the rows [r0, r1): sum |b - A x| with the row sum in upstream's
// order ((b - diag x) - faces ascending), added to the running total; STEP_GS_RES advances the sweep counter and
// decides whether the loop goes on
// HALO (nranks > 1): the neighbours' psi of the exchange issued just before this kernel; the coupled terms are added
// last, in (patch, face) order, with the negated coefficients (lduMatrix::residual: rA[faceCells] -= (-bou)*psi_nbr)
template <bool C16, int B, int CT, bool HALO>
__global__ void __launch_bounds__(kBlock, CT)
k_gs_resid(int r0, int r1, const int64_t* __restrict__ sliceBase, const uint32_t* __restrict__ rowLen,
           EllCols E, const double* __restrict__ val, const double* __restrict__ diag,
           const double* __restrict__ b, const double* __restrict__ x, const int* __restrict__ rowB,
           const int* __restrict__ bStart, const int* __restrict__ bSlot, const double* __restrict__ bou,
           const double* recvNccl, Halo H, Reduce R) {
    if (R.S->done) return;
    const double* recv = nullptr;
    if (HALO) recv = halo_acquire(H, recvNccl, R.S);
    double s[1] = {0.0};
    const int stride = gridDim.x * kBlock;
    int r = r0 + blockIdx.x * kBlock + threadIdx.x;
    uint32_t len = 0;
    int64_t sb = 0;
    double bv = 0.0, dv = 1.0;
    int bi = -1;
#define B200_GSR_LOAD(ROW, LEN, SB, BV, DV, BI) \
    { LEN = rowLen[ROW]; SB = sliceBase[(ROW) >> 5]; BV = b[ROW]; DV = diag[ROW]; if (HALO) BI = rowB[ROW]; }
    if (r < r1) B200_GSR_LOAD(r, len, sb, bv, dv, bi)
    while (r < r1) {
        const int rn = r + stride;
        uint32_t lenN = 0;
        int64_t sbN = 0;
        double bvN = 0.0, dvN = 1.0;
        int biN = -1;
        if (B > 0 && rn < r1) B200_GSR_LOAD(rn, lenN, sbN, bvN, dvN, biN)
        double w = __dadd_rn(bv, -__dmul_rn(dv, x[r]));
        w = eis_row_sub<B, C16, false>(E, val, x, sb + (r & 31), 0, (int)(len >> 16), w);
        if (HALO && bi >= 0) {
            for (int e = bStart[bi]; e < bStart[bi + 1]; ++e) {
                const int slot = bSlot[e];
                w = __dadd_rn(w, __dmul_rn(bou[slot], __ldcg(&recv[slot])));
            }
        }
        s[0] = __dadd_rn(s[0], fabs(w));
        if (B == 0 && rn < r1) B200_GSR_LOAD(rn, lenN, sbN, bvN, dvN, biN)
        r = rn; len = lenN; sb = sbN; bv = bvN; dv = dvN; bi = biN;
    }
#undef B200_GSR_LOAD
    reduce_finish<1>(s, R);
}

// ---- PBiCG + DILU on asymmetric lduMatrices ------------------------------------------------
// Replaces PBiCG::solve and DILUPreconditioner::calcReciprocalD / precondition / preconditionT (OF-dev PBiCG.C,
// DILUPreconditioner.C), which the reference's cases select for their transport equations
// (cases/wallFireSpread2D/system/fvSolution:66-73; the 2.4.x golden logs of steckler: 207 `DILUPBiCG:` lines).
// DILU's face loops -- the forward one runs in losort order upstream -- are, row by row,
//     rD_u  = diag_u - sum_{lower nbrs l} upper_f lower_f / rD_l                         (then reciprocal)
//     w_u   = rD_u r_u - sum_{lower nbrs} (rD_u lower_f) w_l        forward, faces ascending within the row
//     w_l  -= (rD_l upper_f) w_u                                      backward, faces descending within the row
// i.e. the DIC sweeps with the row's own coefficient of each face, which the full-row ELL of an asymmetric matrix
// already holds (k_fill_values_asym); preconditionT is the same pair of sweeps over the TRANSPOSED values (valT: the
// other coefficient of each face), and Tmul is Amul over valT.  Rows grouped by dependency level (Ordering::Levels)
// reproduce upstream bit for bit (tests/test_bicg_plan.py); grouped by colour they are the DILU-class stand-in.
template <bool C16>
__global__ void __launch_bounds__(kBlock)
k_dilu_calc_rd(ColourRows cr, const int64_t* __restrict__ sliceBase, const uint32_t* __restrict__ rowLen, EllCols E,
               const double* __restrict__ val, const double* __restrict__ valT, const double* __restrict__ diag,
               double* __restrict__ rD) {
    B200_FOR_COLOUR_ROWS(cr, r) {
        const int64_t base = sliceBase[r >> 5] + (r & 31);
        const int nLower = (int)(rowLen[r] & 0xffffu);
        double d = diag[r];
        for (int j = 0; j < nLower; ++j) {
            const int64_t e = base + 32 * (int64_t)j;
            // (valT, val) of a lower entry = (upper_f, lower_f); rD of an earlier group: plain (coherent) load
            d = __dadd_rn(d, -__ddiv_rn(__dmul_rn(valT[e], val[e]), rD[ell_col<C16>(E, e)]));
        }
        rD[r] = d;
    }
}

// forward / backward triangular sweep over one group: the DIC-class sweeps without their dot products
template <bool C16>
__global__ void __launch_bounds__(kBlock)
k_tri_fwd(ColourRows cr, const int64_t* __restrict__ sliceBase, const uint32_t* __restrict__ rowLen, EllCols E,
          const double* __restrict__ val, const double* __restrict__ rD, const double* __restrict__ rIn, double* w,
          const Scalars* S) {
    if (S->done) return;
    B200_FOR_COLOUR_ROWS(cr, r) {
        const int64_t base = sliceBase[r >> 5] + (r & 31);
        const int nLower = (int)(rowLen[r] & 0xffffu);
        const double d = rD[r];
        double acc = __dmul_rn(d, rIn[r]);
        for (int j = 0; j < nLower; ++j) {
            const int64_t e = base + 32 * (int64_t)j;
            acc = __dadd_rn(acc, -__dmul_rn(__dmul_rn(d, val[e]), w[ell_col<C16>(E, e)]));
        }
        w[r] = acc;
    }
}
template <bool C16>
__global__ void __launch_bounds__(kBlock)
k_tri_bwd(ColourRows cr, const int64_t* __restrict__ sliceBase, const uint32_t* __restrict__ rowLen, EllCols E,
          const double* __restrict__ val, const double* __restrict__ rD, double* w, const Scalars* S) {
    if (S->done) return;
    B200_FOR_COLOUR_ROWS(cr, r) {
        const int64_t base = sliceBase[r >> 5] + (r & 31);
        const uint32_t len = rowLen[r];
        const int nLower = (int)(len & 0xffffu), nTotal = (int)(len >> 16);
        const double d = rD[r];
        double acc = w[r];
        for (int j = nTotal - 1; j >= nLower; --j) {
            const int64_t e = base + 32 * (int64_t)j;
            acc = __dadd_rn(acc, -__dmul_rn(__dmul_rn(d, val[e]), w[ell_col<C16>(E, e)]));
        }
        w[r] = acc;
    }
}

// gSumProd(a, b) -> step (wArT = (wA, rT): STEP_WARA; wApT = (wA, pT): STEP_WAPA)
__global__ void __launch_bounds__(kBlock)
k_dot2(int N, const double* __restrict__ a, const double* __restrict__ b, Reduce R) {
    if (R.S->done) return;
    double s[1] = {0.0};
    B200_VEC_LOOP(N,
        { const double2 x = reinterpret_cast<const double2*>(a)[i];
          const double2 y = reinterpret_cast<const double2*>(b)[i];
          s[0] = __dadd_rn(s[0], __dmul_rn(x.x, y.x));
          s[0] = __dadd_rn(s[0], __dmul_rn(x.y, y.y)); },
        { s[0] = __dadd_rn(s[0], __dmul_rn(a[i], b[i])); })
    reduce_finish<1>(s, R);
}

// diagonalPreconditioner on both residuals: wA = rD rA, wT = rD rT
__global__ void __launch_bounds__(kBlock)
k_bicg_diag(int N, const double* __restrict__ rD, const double* __restrict__ rA, const double* __restrict__ rT,
            double* __restrict__ wA, double* __restrict__ wT, const Scalars* S) {
    if (S->done) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const double d = rD[i];
        wA[i] = __dmul_rn(d, rA[i]);
        wT[i] = __dmul_rn(d, rT[i]);
    }
}

// rT = source - wT (the transpose residual of the initial guess)
__global__ void __launch_bounds__(kBlock)
k_sub(int N, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
        out[i] = __dadd_rn(a[i], -b[i]);
}

// search directions: pA = zA + beta pA, pT = zT + beta pT (first iteration: pA = zA, pT = zT)
__global__ void __launch_bounds__(kBlock)
k_bicg_p(int N, double* __restrict__ pA, const double* __restrict__ zA, double* __restrict__ pT,
         const double* __restrict__ zT, const Scalars* S) {
    if (S->done) return;
    const bool first = (S->nIter == 0);
    const double beta = S->beta;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        double a = zA[i], t = zT[i];
        if (!first) {
            a = __dadd_rn(a, __dmul_rn(beta, pA[i]));
            t = __dadd_rn(t, __dmul_rn(beta, pT[i]));
        }
        pA[i] = a;
        pT[i] = t;
    }
}

// psi += alpha pA; rA -= alpha wA; rT -= alpha wT; gSumMag(rA) -> STEP_RES
__global__ void __launch_bounds__(kBlock)
k_bicg_r(int N, double* __restrict__ psi, const double* __restrict__ pA, double* __restrict__ rA,
         const double* __restrict__ wA, double* __restrict__ rT, const double* __restrict__ wT, Reduce R) {
    if (R.S->done) return;
    const double alpha = R.S->alpha;
    double s[1] = {0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        psi[i] = __dadd_rn(psi[i], __dmul_rn(alpha, pA[i]));
        const double r = __dadd_rn(rA[i], -__dmul_rn(alpha, wA[i]));
        rA[i] = r;
        rT[i] = __dadd_rn(rT[i], -__dmul_rn(alpha, wT[i]));
        s[0] = __dadd_rn(s[0], fabs(r));
    }
    reduce_finish<1>(s, R);
}

// ---- assembly: gaussLaplacianScheme::fvmLaplacianUncorrected + negSumDiag ------------------
// (OF-dev gaussLaplacianScheme.C, lduMatrixOperations.C)
__global__ void __launch_bounds__(kBlock)
k_face_coeff(int F, const double* __restrict__ gamma, const double* __restrict__ magSf,
             const double* __restrict__ delta, double sign, double* __restrict__ upper) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x)
        upper[f] = __dmul_rn(sign, __dmul_rn(delta[f], __dmul_rn(gamma[f], magSf[f])));
}
// diag[c] += (0 - upper[f1] - upper[f2] ...) over the faces of c in face order: the two
// sorted segments (neighbour side via losort, then owner side) of the natural-order plan.
__global__ void __launch_bounds__(kBlock)
k_neg_sum_diag(int N, const int64_t* __restrict__ sliceBase, const uint32_t* __restrict__ rowLen,
               const int* __restrict__ faceOf, const int* __restrict__ perm,
               const double* __restrict__ upper, double* __restrict__ diag) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) {
        const int64_t base = sliceBase[r >> 5] + (r & 31);
        const int n = (int)(rowLen[r] >> 16);
        const int cell = perm ? perm[r] : r;   // diag is in the caller's (natural) cell order
        const double d0 = diag[cell];
        double d = 0.0;
        int j = 0;
        // batches of 4 independent (index, value) load pairs; the adds stay in face order
        for (; j + 4 <= n; j += 4) {
            const int64_t e = base + 32 * (int64_t)j;
            const int f0 = faceOf[e], f1 = faceOf[e + 32], f2 = faceOf[e + 64], f3 = faceOf[e + 96];
            const double u0 = __ldg(&upper[f0]), u1 = __ldg(&upper[f1]), u2 = __ldg(&upper[f2]),
                         u3 = __ldg(&upper[f3]);
            d = __dadd_rn(d, -u0);
            d = __dadd_rn(d, -u1);
            d = __dadd_rn(d, -u2);
            d = __dadd_rn(d, -u3);
        }
        if (j + 2 <= n) {
            const int64_t e = base + 32 * (int64_t)j;
            const int f0 = faceOf[e], f1 = faceOf[e + 32];
            const double u0 = __ldg(&upper[f0]), u1 = __ldg(&upper[f1]);
            d = __dadd_rn(d, -u0);
            d = __dadd_rn(d, -u1);
            j += 2;
        }
        if (j < n) d = __dadd_rn(d, -__ldg(&upper[faceOf[base + 32 * (int64_t)j]]));
        diag[cell] = __dadd_rn(d0, d);
    }
}

// negSumDiag in the CALLER'S cell order, for plans whose rows were renumbered (RCM): there the plan's rows are a
// permutation of the cells, and gathering upper[faceOf[e]] row by row jumps all over the face list (measured
// 0.26 of the roofline on the block-shuffled polyhedral workload).  In natural order the owner-side faces of a
// cell are CONTIGUOUS in the face list (upper-triangular order: faces sorted by owner) and the neighbour-side
// faces (losort: ascending face index, all below the cell's first owner face) belong to nearby owners, so one
// half of the reads streams and the other half stays in L2.  Same summation order as OpenFOAM's face loop:
// ascending face index.  ownerStart / losortStart / losort are OF-dev lduAddressing's own lazily-built lists.
__global__ void __launch_bounds__(kBlock)
k_neg_sum_diag_nat(int N, const int* __restrict__ ownerStart, const int* __restrict__ losortStart,
                   const int* __restrict__ losort, const double* __restrict__ upper, double* __restrict__ diag) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) {
        const int k0 = losortStart[c], k1 = losortStart[c + 1];
        const int f0 = ownerStart[c], f1 = ownerStart[c + 1];
        const double d0 = diag[c];
        double d = 0.0;
        int k = k0;
        for (; k + 4 <= k1; k += 4) {
            const int a0 = losort[k], a1 = losort[k + 1], a2 = losort[k + 2], a3 = losort[k + 3];
            const double u0 = __ldg(&upper[a0]), u1 = __ldg(&upper[a1]), u2 = __ldg(&upper[a2]), u3 = __ldg(&upper[a3]);
            d = __dadd_rn(d, -u0);
            d = __dadd_rn(d, -u1);
            d = __dadd_rn(d, -u2);
            d = __dadd_rn(d, -u3);
        }
        for (; k < k1; ++k) d = __dadd_rn(d, -__ldg(&upper[losort[k]]));
        int f = f0;
        for (; f + 4 <= f1; f += 4) {
            const double u0 = upper[f], u1 = upper[f + 1], u2 = upper[f + 2], u3 = upper[f + 3];
            d = __dadd_rn(d, -u0);
            d = __dadd_rn(d, -u1);
            d = __dadd_rn(d, -u2);
            d = __dadd_rn(d, -u3);
        }
        for (; f < f1; ++f) d = __dadd_rn(d, -upper[f]);
        diag[c] = __dadd_rn(d0, d);
    }
}

// ---- whole p_rghEqn: ddt + explicit terms + fvc::div + laplacian diagonal + boundary fold ------
// (reference solver/pEqn.H:26-37, solver/phrghEqn.H:43-46; OF-dev EulerDdtScheme.C,
// fvMatrix.C, surfaceIntegrate.C, lduMatrixOperations.C).  One row per thread: the face -> cell sums
// of fvc::div and negSumDiag are gathers over the row's faces in ascending face order (== the order
// in which OpenFOAM's face loops update the cell), the patch sums are gathers over the cell's
// boundary faces in patch order (CSR built by b200_set_boundary_faces): sorted segments, no atomics.
constexpr int kMaxExplicit = 8;
struct PrghDev {
    double rDeltaT, divSign;
    const double *V, *psi, *psi0, *p0, *phi, *Su, *bPhi, *bInt, *bBou;
    const double* ex[kMaxExplicit];
    int nExplicit;
    const int *bfStart, *bfOrder;   // [N+1], [nB]: boundary faces of each cell, patch order
};

__global__ void __launch_bounds__(kBlock)
k_prgh_cell(int N, const int64_t* __restrict__ sliceBase, const uint32_t* __restrict__ rowLen,
            const int* __restrict__ faceOf, const int* __restrict__ perm, const int* __restrict__ lowerAddr,
            const double* __restrict__ upper, PrghDev t, double* __restrict__ diag, double* __restrict__ source) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) {
        const int c = perm ? perm[r] : r;
        const int64_t base = sliceBase[r >> 5] + (r & 31);
        const int n = (int)(rowLen[r] >> 16);
        const double V = t.V[c];
        double dsum = 0.0, ivf = 0.0;
        for (int j = 0; j < n; ++j) {
            const int f = faceOf[base + 32 * (int64_t)j];
            dsum = __dadd_rn(dsum, -__ldg(&upper[f]));
            if (t.phi) {
                const double ph = __ldg(&t.phi[f]);
                ivf = __dadd_rn(ivf, (__ldg(&lowerAddr[f]) == c) ? ph : -ph);   // owner += phi, neighbour -= phi
            }
        }
        const int b0 = t.bfStart ? t.bfStart[c] : 0, b1 = t.bfStart ? t.bfStart[c + 1] : 0;
        double d = 0.0, s = 0.0;
        if (t.psi) {
            d = __dmul_rn(__dmul_rn(t.rDeltaT, t.psi[c]), V);
            s = __dmul_rn(__dmul_rn(__dmul_rn(t.rDeltaT, t.psi0[c]), t.p0[c]), V);
        }
        for (int k = 0; k < t.nExplicit; ++k) s = __dadd_rn(s, -__dmul_rn(V, t.ex[k][c]));
        if (t.phi) {
            if (t.bPhi)
                for (int b = b0; b < b1; ++b) ivf = __dadd_rn(ivf, t.bPhi[t.bfOrder[b]]);
            ivf = __ddiv_rn(ivf, V);
            const double q = __dmul_rn(V, ivf);
            s = __dadd_rn(s, t.divSign < 0 ? -q : q);
        }
        d = __dadd_rn(d, dsum);
        if (t.Su) s = __dadd_rn(s, __dmul_rn(V, t.Su[c]));
        if (t.bInt)
            for (int b = b0; b < b1; ++b) d = __dadd_rn(d, t.bInt[t.bfOrder[b]]);
        if (t.bBou)
            for (int b = b0; b < b1; ++b) s = __dadd_rn(s, t.bBou[t.bfOrder[b]]);
        diag[c] = d;
        source[c] = s;
    }
}

// ---- whole-solve kernel for small systems (one thread-block cluster) ---------------------------
// The reference's own cases are small (cases/steckler: 9 000 cells): there the multi-kernel loop is
// pure launch latency (~50 us per iteration = 3 launches + device-side scalar steps), no faster
// than one host core.  k_pcg_small runs the WHOLE of PCG::solve (OF-dev PCG.C) --
// initial residual, normFactor, preconditioner set-up, the iteration loop and its convergence
// test -- in ONE launch of one thread-block cluster (8 or 16 CTAs x 1024 threads, co-scheduled on
// one GPC).  Phases are separated by hardware cluster barriers (~0.2 us, release/acquire at
// cluster scope) instead of kernel boundaries; every CTA forms the global sums from the per-CTA
// partials in the same fixed order and advances its own copy of the CG scalars, so no broadcast
// is needed and every CTA takes the same branch.  Vectors and matrix stay L2-resident.
// Same row-sum order, same element-wise arithmetic and the same scalar_step() as the large path.
namespace cgx = cooperative_groups;
constexpr int kSmallBlock = 1024;
constexpr int kSmallMaxCtas = 16;

struct SmallArgs {
    int N, precond /* B200_PRECOND_* */, nColours;
    const int* colourStart;        // [nColours+1] (device)
    const int64_t* sliceBase;
    const uint32_t* rowLen;
    const int* col;
    const double* val;
    const double* diag;
    const double* src;
    double *psi, *rA, *pA, *wA, *rD;
    Scalars* S;
    double* partials;              // [2][kSmallMaxCtas][kNSums]
};

struct SmallCtx {
    cgx::cluster_group cluster;
    Scalars* sS;                   // this CTA's copy (shared memory)
    double* partials;
    double (*sh)[kSmallBlock / 32];
    unsigned nred;
    int nCtas, cta, gtid, nThreads;
};

// cluster-wide sum of NV per-thread values + scalar step; every thread of the cluster calls it
template <int NV>
__device__ __forceinline__ void small_reduce(SmallCtx& c, double (&v)[NV], int step) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const double s = warp_sum(v[i]);
        if (lane == 0) c.sh[i][w] = s;
    }
    __syncthreads();
    double* buf = c.partials + (size_t)(c.nred & 1u) * kSmallMaxCtas * kNSums;
    if (w == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = c.sh[i][lane];      // kSmallBlock / 32 == 32 warps
            s = warp_sum(s);
            if (lane == 0) buf[c.cta * kNSums + i] = s;
        }
    }
    c.cluster.sync();                      // partials of every CTA visible (release/acquire)
    if (threadIdx.x == 0) {
        double g[kNSums];
#pragma unroll
        for (int i = 0; i < kNSums; ++i) g[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double t = 0.0;
            for (int b = 0; b < c.nCtas; ++b) t = __dadd_rn(t, __ldcg(&buf[b * kNSums + i]));
            g[i] = t;
        }
        scalar_step(step, c.sS, g);
    }
    c.nred++;
    __syncthreads();
}

// (x is rewritten between phases of the same launch: plain coherent loads, never the read-only path)
__device__ __forceinline__ double small_row_amul(const SmallArgs& a, int r, const double* x,
                                                 double xr, double d, double* sa) {
    const int64_t base = a.sliceBase[r >> 5] + (r & 31);
    const int n = (int)(a.rowLen[r] >> 16);
    double acc = __dmul_rn(d, xr);
    double s = d;
    for (int j = 0; j < n; ++j) {
        const int64_t e = base + 32 * (int64_t)j;
        const double v = a.val[e];
        acc = __dadd_rn(acc, __dmul_rn(v, x[a.col[e]]));
        s = __dadd_rn(s, v);
    }
    if (sa) *sa = s;
    return acc;
}

// DIC-class apply (rank-local), colour by colour with a cluster barrier between colours; returns
// this thread's share of (wA, rA)
__device__ __forceinline__ double small_dic_apply(SmallCtx& c, const SmallArgs& a) {
    double dot = 0.0;
    const int C = a.nColours;
    for (int k = 0; k < C; ++k) {
        for (int r = a.colourStart[k] + c.gtid; r < a.colourStart[k + 1]; r += c.nThreads) {
            const int64_t base = a.sliceBase[r >> 5] + (r & 31);
            const int nLower = (int)(a.rowLen[r] & 0xffffu);
            const double d = a.rD[r], rr = a.rA[r];
            double w = __dmul_rn(d, rr);
            for (int j = 0; j < nLower; ++j) {
                const int64_t e = base + 32 * (int64_t)j;
                w = __dadd_rn(w, -__dmul_rn(__dmul_rn(d, a.val[e]), a.wA[a.col[e]]));
            }
            a.wA[r] = w;
            if (k == C - 1) dot = __dadd_rn(dot, __dmul_rn(w, rr));
        }
        if (k + 1 < C) c.cluster.sync();
    }
    for (int k = C - 2; k >= 0; --k) {
        c.cluster.sync();
        for (int r = a.colourStart[k] + c.gtid; r < a.colourStart[k + 1]; r += c.nThreads) {
            const int64_t base = a.sliceBase[r >> 5] + (r & 31);
            const uint32_t len = a.rowLen[r];
            const int nLower = (int)(len & 0xffffu), nTotal = (int)(len >> 16);
            const double d = a.rD[r];
            double w = a.wA[r];
            for (int j = nTotal - 1; j >= nLower; --j) {
                const int64_t e = base + 32 * (int64_t)j;
                w = __dadd_rn(w, -__dmul_rn(__dmul_rn(d, a.val[e]), a.wA[a.col[e]]));
            }
            a.wA[r] = w;
            dot = __dadd_rn(dot, __dmul_rn(w, a.rA[r]));
        }
    }
    return dot;
}

__global__ void __launch_bounds__(kSmallBlock, 1) k_pcg_small(SmallArgs a) {
    __shared__ Scalars sS;
    __shared__ double sh[kNSums][kSmallBlock / 32];
    SmallCtx c{cgx::this_cluster(), &sS, a.partials, sh, 0u, 0, 0, 0, 0};
    c.nCtas = (int)c.cluster.num_blocks();
    c.cta = (int)c.cluster.block_rank();
    c.nThreads = c.nCtas * kSmallBlock;
    c.gtid = c.cta * kSmallBlock + (int)threadIdx.x;
    if (threadIdx.x == 0) sS = *a.S;          // controls written by the host
    __syncthreads();
    const int N = a.N, T = c.nThreads, t0 = c.gtid;
    const bool dic = a.precond >= 2, diagp = a.precond == 1;

    // wA = A psi, sumA -> pA (lduMatrix::Amul + sumA);  gSum(psi)
    {
        double s[1] = {0.0};
        for (int r = t0; r < N; r += T) {
            const double xr = a.psi[r];
            double sa;
            a.wA[r] = small_row_amul(a, r, a.psi, xr, a.diag[r], &sa);
            a.pA[r] = sa;
            s[0] = __dadd_rn(s[0], xr);
        }
        small_reduce<1>(c, s, STEP_SUMPSI);
    }
    // rA = source - wA; normFactor; initial residual
    {
        double s[2] = {0.0, 0.0};
        const double xRef = sS.xRef;
        for (int r = t0; r < N; r += T) {
            const double w = a.wA[r], b = a.src[r];
            const double t = __dmul_rn(a.pA[r], xRef);
            s[0] = __dadd_rn(s[0], __dadd_rn(fabs(__dadd_rn(w, -t)), fabs(__dadd_rn(b, -t))));
            const double rr = __dadd_rn(b, -w);
            s[1] = __dadd_rn(s[1], fabs(rr));
            a.rA[r] = rr;
        }
        small_reduce<2>(c, s, STEP_NORM);
    }
    if (!sS.done) {
        // preconditioner set-up
        if (diagp) {
            for (int r = t0; r < N; r += T) a.rD[r] = __ddiv_rn(1.0, a.diag[r]);
        } else if (dic) {
            for (int k = 0; k < a.nColours; ++k) {      // calcReciprocalD, colour by colour
                for (int r = a.colourStart[k] + t0; r < a.colourStart[k + 1]; r += T) {
                    const int64_t base = a.sliceBase[r >> 5] + (r & 31);
                    const int nLower = (int)(a.rowLen[r] & 0xffffu);
                    double d = a.diag[r];
                    for (int j = 0; j < nLower; ++j) {
                        const int64_t e = base + 32 * (int64_t)j;
                        const double v = a.val[e];
                        d = __dadd_rn(d, -__ddiv_rn(__dmul_rn(v, v), a.rD[a.col[e]]));
                    }
                    a.rD[r] = d;
                }
                c.cluster.sync();
            }
            for (int r = t0; r < N; r += T) a.rD[r] = __ddiv_rn(1.0, a.rD[r]);
            c.cluster.sync();
        }
        // first wArA = (precondition(rA), rA)
        {
            double s[1] = {0.0};
            if (dic) s[0] = small_dic_apply(c, a);
            else
                for (int r = t0; r < N; r += T) {
                    const double rr = a.rA[r];
                    const double w = diagp ? __dmul_rn(a.rD[r], rr) : rr;
                    s[0] = __dadd_rn(s[0], __dmul_rn(w, rr));
                }
            small_reduce<1>(c, s, STEP_WARA);
        }
    }
    // PCG loop (regrouped like the large path: psi update deferred into the next p update)
    while (!sS.done) {
        {
            const bool first = (sS.nIter == 0);
            const double beta = sS.beta, alpha = sS.alpha;
            for (int r = t0; r < N; r += T) {
                double p = dic ? a.wA[r] : (diagp ? __dmul_rn(a.rD[r], a.rA[r]) : a.rA[r]);
                if (!first) {
                    const double po = a.pA[r];
                    a.psi[r] = __dadd_rn(a.psi[r], __dmul_rn(alpha, po));
                    p = __dadd_rn(p, __dmul_rn(beta, po));
                }
                a.pA[r] = p;
            }
        }
        c.cluster.sync();                    // pA complete before anyone gathers it
        {
            double s[1] = {0.0};
            for (int r = t0; r < N; r += T) {
                const double xr = a.pA[r];
                const double y = small_row_amul(a, r, a.pA, xr, a.diag[r], nullptr);
                a.wA[r] = y;
                s[0] = __dadd_rn(s[0], __dmul_rn(y, xr));
            }
            small_reduce<1>(c, s, STEP_WAPA);
        }
        if (sS.done) break;                  // singular
        {
            const double alpha = sS.alpha;
            double s[2] = {0.0, 0.0};
            for (int r = t0; r < N; r += T) {
                const double rr = __dadd_rn(a.rA[r], -__dmul_rn(alpha, a.wA[r]));
                a.rA[r] = rr;
                s[0] = __dadd_rn(s[0], fabs(rr));
                if (diagp) s[1] = __dadd_rn(s[1], __dmul_rn(__dmul_rn(a.rD[r], rr), rr));
                else if (!dic) s[1] = __dadd_rn(s[1], __dmul_rn(rr, rr));
            }
            small_reduce<2>(c, s, dic ? STEP_RES : STEP_RES_WARA);
        }
        if (dic && !sS.done) {
            double s[1];
            s[0] = small_dic_apply(c, a);
            small_reduce<1>(c, s, STEP_WARA);
        }
    }
    if (sS.pendingPsi) {
        const double alpha = sS.alpha;
        for (int r = t0; r < N; r += T) a.psi[r] = __dadd_rn(a.psi[r], __dmul_rn(alpha, a.pA[r]));
    }
    if (c.cta == 0 && threadIdx.x == 0) *a.S = sS;
}

// ---- on-chip variant of the whole-solve kernel -------------------------------------------------
// k_pcg_small keeps matrix and vectors in L2: every phase pays 1-3 dependent L2 round trips after the
// L1 flush of the cluster barrier (measured 9.3 us per PCG iteration on the 9 000-cell steckler
// system).  When the system fits the cluster's shared memory (<= RPT*1024 rows per CTA, RPT <= 2) the
// matrix rows (packed column + value), diag, rD and the two gathered vectors (pA, and wA for the
// DIC-class sweeps) live in SHARED memory for the whole solve, the thread-private vectors (rA, A*pA,
// previous pA) in registers, and neighbour values are read from the owning CTA's shared memory
// over DSMEM (cluster.map_shared_rank).  Reductions push each CTA's partial into every CTA's shared
// memory before the barrier, so after it every CTA sums local data.  Global memory is touched only
// to load the system, to update psi (off the critical path) and to store the status block.
// Row r is owned by thread (r mod T) of the cluster, slot r / T;  T = nCtas*1024.
// Same arithmetic, same row-sum order, same scalar_step() as every other path.
struct FastArgs {
    int N, precond, nColours, W;   // W = max faces per row
    const int* colourStart;
    const int64_t* sliceBase;
    const uint32_t* rowLen;
    const int* col;
    const double* val;
    const double* diag;
    const double* src;
    double* psi;
    Scalars* S;
};

__host__ __device__ inline size_t fast_smem_bytes(int RPT, int W) {
    const size_t S = (size_t)RPT * kSmallBlock;
    return S * ((size_t)W * 12 + 32) + sizeof(double) * 2 * kSmallMaxCtas * kNSums + 16;
}

template <int RPT>
__global__ void __launch_bounds__(kSmallBlock, 1) k_pcg_small_fast(FastArgs a) {
    extern __shared__ __align__(16) unsigned char fsm[];
    __shared__ Scalars sS;
    __shared__ double sh[kNSums][kSmallBlock / 32];
    cgx::cluster_group cluster = cgx::this_cluster();
    constexpr int S = RPT * kSmallBlock;
    const int W = a.W, N = a.N;
    double* sVal = reinterpret_cast<double*>(fsm);            // [W][S]
    double* sP = sVal + (size_t)W * S;                        // [S] pA   (gathered by neighbours)
    double* sW = sP + S;                                      // [S] wA / rD scratch (DIC-class sweeps)
    double* sDg = sW + S;                                     // [S] diag
    double* sRD = sDg + S;                                    // [S] rD
    double* sPart = sRD + S;                                  // [2][kSmallMaxCtas][kNSums]
    uint32_t* sCol = reinterpret_cast<uint32_t*>(sPart + 2 * kSmallMaxCtas * kNSums);   // [W][S] packed
    const int nCtas = (int)cluster.num_blocks(), cta = (int)cluster.block_rank();
    const int T = nCtas * kSmallBlock, tid = (int)threadIdx.x, gtid = cta * kSmallBlock + tid;
    const bool dic = a.precond >= 2, diagp = a.precond == 1;
    unsigned nred = 0;
    if (tid == 0) sS = *a.S;

    int nTot[RPT], nLow[RPT];
    double rA[RPT], wv[RPT], pOwn[RPT];
    // ---- load this thread's rows into shared memory ------------------------------------------------
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        const int r = k * T + gtid, slot = k * kSmallBlock + tid;
        nTot[k] = nLow[k] = 0;
        rA[k] = wv[k] = pOwn[k] = 0.0;
        sP[slot] = 0.0;
        sW[slot] = 0.0;
        sDg[slot] = 1.0;
        sRD[slot] = 0.0;
        if (r < N) {
            const int64_t base = a.sliceBase[r >> 5] + (r & 31);
            const uint32_t len = a.rowLen[r];
            nTot[k] = (int)(len >> 16);
            nLow[k] = (int)(len & 0xffffu);
            for (int j = 0; j < nTot[k]; ++j) {
                const int64_t e = base + 32 * (int64_t)j;
                const int c = a.col[e];
                const int kk = c / T, g = c - kk * T;
                sCol[(size_t)j * S + slot] = ((uint32_t)(g >> 10) << 20) | (uint32_t)(kk * kSmallBlock + (g & (kSmallBlock - 1)));
                sVal[(size_t)j * S + slot] = a.val[e];
            }
            sDg[slot] = a.diag[r];
            sP[slot] = a.psi[r];            // first Amul gathers psi
        }
    }
    cluster.sync();

    auto gather = [&](double* arr, uint32_t packed) -> double {
        return cluster.map_shared_rank(arr, packed >> 20)[packed & 0xfffffu];
    };
    // cluster-wide sums: push this CTA's partial into every CTA, barrier, sum locally in CTA order
    auto reduce = [&](double* v, int nv, int step) {
        const int lane = tid & 31, w = tid >> 5;
        for (int i = 0; i < nv; ++i) {
            const double s_ = warp_sum(v[i]);
            if (lane == 0) sh[i][w] = s_;
        }
        __syncthreads();
        double* buf = sPart + (size_t)(nred & 1u) * kSmallMaxCtas * kNSums;
        if (w == 0) {
            for (int i = 0; i < nv; ++i) {
                double s_ = sh[i][lane];
                s_ = warp_sum(s_);
                s_ = __shfl_sync(0xffffffffu, s_, 0);
                if (lane < nCtas) cluster.map_shared_rank(buf, lane)[cta * kNSums + i] = s_;
            }
        }
        cluster.sync();
        if (tid == 0) {
            double g[kNSums];
            for (int i = 0; i < kNSums; ++i) g[i] = 0.0;
            for (int i = 0; i < nv; ++i) {
                double t_ = 0.0;
                for (int b = 0; b < nCtas; ++b) t_ = __dadd_rn(t_, buf[b * kNSums + i]);
                g[i] = t_;
            }
            scalar_step(step, &sS, g);
        }
        nred++;
        __syncthreads();
    };
    // DIC-class apply: wA -> sW (own slots), returns this thread's share of (wA, rA)
    auto dic_apply = [&]() -> double {
        double dot = 0.0;
        const int C = a.nColours;
        for (int c = 0; c < C; ++c) {
            const int r0 = a.colourStart[c], r1 = a.colourStart[c + 1];
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int r = k * T + gtid, slot = k * kSmallBlock + tid;
                if (r >= r0 && r < r1) {
                    const double d = sRD[slot];
                    double w = __dmul_rn(d, rA[k]);
                    for (int j = 0; j < nLow[k]; ++j)
                        w = __dadd_rn(w, -__dmul_rn(__dmul_rn(d, sVal[(size_t)j * S + slot]),
                                                    gather(sW, sCol[(size_t)j * S + slot])));
                    sW[slot] = w;
                    if (c == C - 1) dot = __dadd_rn(dot, __dmul_rn(w, rA[k]));
                }
            }
            if (c + 1 < C) cluster.sync();
        }
        for (int c = C - 2; c >= 0; --c) {
            cluster.sync();
            const int r0 = a.colourStart[c], r1 = a.colourStart[c + 1];
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int r = k * T + gtid, slot = k * kSmallBlock + tid;
                if (r >= r0 && r < r1) {
                    const double d = sRD[slot];
                    double w = sW[slot];
                    for (int j = nTot[k] - 1; j >= nLow[k]; --j)
                        w = __dadd_rn(w, -__dmul_rn(__dmul_rn(d, sVal[(size_t)j * S + slot]),
                                                    gather(sW, sCol[(size_t)j * S + slot])));
                    sW[slot] = w;
                    dot = __dadd_rn(dot, __dmul_rn(w, rA[k]));
                }
            }
        }
        return dot;
    };

    // ---- wA = A psi, sumA; gSum(psi) ------------------------------------------------------------------
    double sa[RPT];
    {
        double s_[1] = {0.0};
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int r = k * T + gtid, slot = k * kSmallBlock + tid;
            sa[k] = 0.0;
            if (r < N) {
                const double xr = sP[slot], d = sDg[slot];
                double acc = __dmul_rn(d, xr), q = d;
                for (int j = 0; j < nTot[k]; ++j) {
                    const double v = sVal[(size_t)j * S + slot];
                    acc = __dadd_rn(acc, __dmul_rn(v, gather(sP, sCol[(size_t)j * S + slot])));
                    q = __dadd_rn(q, v);
                }
                wv[k] = acc;
                sa[k] = q;
                s_[0] = __dadd_rn(s_[0], xr);
            }
        }
        reduce(s_, 1, STEP_SUMPSI);
    }
    {
        double s_[2] = {0.0, 0.0};
        const double xRef = sS.xRef;
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int r = k * T + gtid;
            if (r < N) {
                const double w = wv[k], b = a.src[r];
                const double t_ = __dmul_rn(sa[k], xRef);
                s_[0] = __dadd_rn(s_[0], __dadd_rn(fabs(__dadd_rn(w, -t_)), fabs(__dadd_rn(b, -t_))));
                const double rr = __dadd_rn(b, -w);
                s_[1] = __dadd_rn(s_[1], fabs(rr));
                rA[k] = rr;
            }
        }
        reduce(s_, 2, STEP_NORM);
    }
    if (!sS.done) {
        if (diagp) {
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int slot = k * kSmallBlock + tid;
                if (k * T + gtid < N) sRD[slot] = __ddiv_rn(1.0, sDg[slot]);
            }
        } else if (dic) {
            for (int c = 0; c < a.nColours; ++c) {      // calcReciprocalD: un-inverted rD staged in sW
                const int r0 = a.colourStart[c], r1 = a.colourStart[c + 1];
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    const int r = k * T + gtid, slot = k * kSmallBlock + tid;
                    if (r >= r0 && r < r1) {
                        double d = sDg[slot];
                        for (int j = 0; j < nLow[k]; ++j) {
                            const double v = sVal[(size_t)j * S + slot];
                            d = __dadd_rn(d, -__ddiv_rn(__dmul_rn(v, v), gather(sW, sCol[(size_t)j * S + slot])));
                        }
                        sW[slot] = d;
                    }
                }
                cluster.sync();
            }
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int slot = k * kSmallBlock + tid;
                if (k * T + gtid < N) sRD[slot] = __ddiv_rn(1.0, sW[slot]);
            }
            cluster.sync();      // everyone has consumed sW as rD scratch before the first sweep rewrites it
        }
        double s_[1] = {0.0};
        if (dic) s_[0] = dic_apply();
        else {
#pragma unroll
            for (int k = 0; k < RPT; ++k)
                if (k * T + gtid < N) {
                    const double w = diagp ? __dmul_rn(sRD[k * kSmallBlock + tid], rA[k]) : rA[k];
                    s_[0] = __dadd_rn(s_[0], __dmul_rn(w, rA[k]));
                }
        }
        reduce(s_, 1, STEP_WARA);
    }
    // ---- PCG loop ------------------------------------------------------------------------------------
    while (!sS.done) {
        {
            const bool first = (sS.nIter == 0);
            const double beta = sS.beta, alpha = sS.alpha;
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int r = k * T + gtid, slot = k * kSmallBlock + tid;
                if (r < N) {
                    double p = dic ? sW[slot] : (diagp ? __dmul_rn(sRD[slot], rA[k]) : rA[k]);
                    if (!first) {
                        a.psi[r] = __dadd_rn(a.psi[r], __dmul_rn(alpha, pOwn[k]));
                        p = __dadd_rn(p, __dmul_rn(beta, pOwn[k]));
                    }
                    pOwn[k] = p;
                    sP[slot] = p;
                }
            }
        }
        cluster.sync();
        {
            double s_[1] = {0.0};
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int slot = k * kSmallBlock + tid;
                if (k * T + gtid < N) {
                    double acc = __dmul_rn(sDg[slot], pOwn[k]);
                    for (int j = 0; j < nTot[k]; ++j)
                        acc = __dadd_rn(acc, __dmul_rn(sVal[(size_t)j * S + slot], gather(sP, sCol[(size_t)j * S + slot])));
                    wv[k] = acc;
                    s_[0] = __dadd_rn(s_[0], __dmul_rn(acc, pOwn[k]));
                }
            }
            reduce(s_, 1, STEP_WAPA);
        }
        if (sS.done) break;
        {
            const double alpha = sS.alpha;
            double s_[2] = {0.0, 0.0};
#pragma unroll
            for (int k = 0; k < RPT; ++k)
                if (k * T + gtid < N) {
                    const double rr = __dadd_rn(rA[k], -__dmul_rn(alpha, wv[k]));
                    rA[k] = rr;
                    s_[0] = __dadd_rn(s_[0], fabs(rr));
                    if (diagp) s_[1] = __dadd_rn(s_[1], __dmul_rn(__dmul_rn(sRD[k * kSmallBlock + tid], rr), rr));
                    else if (!dic) s_[1] = __dadd_rn(s_[1], __dmul_rn(rr, rr));
                }
            reduce(s_, 2, dic ? STEP_RES : STEP_RES_WARA);
        }
        if (dic && !sS.done) {
            double s_[1];
            s_[0] = dic_apply();
            reduce(s_, 1, STEP_WARA);
        }
    }
    if (sS.pendingPsi) {
        const double alpha = sS.alpha;
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int r = k * T + gtid;
            if (r < N) a.psi[r] = __dadd_rn(a.psi[r], __dmul_rn(alpha, pOwn[k]));
        }
    }
    if (cta == 0 && tid == 0) *a.S = sS;
    cluster.sync();     // no CTA may exit while others can still read its shared memory
}

// ---- fvMatrix::flux() internal faces (OF-dev fvMatrix.C) --------------------
__global__ void __launch_bounds__(kBlock)
k_flux(int F, const int* __restrict__ l, const int* __restrict__ u,
       const double* __restrict__ upper, const double* __restrict__ psi,
       double* __restrict__ flux) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) {
        const double a = upper[f];
        flux[f] = __dadd_rn(__dmul_rn(a, __ldg(&psi[u[f]])), -__dmul_rn(a, __ldg(&psi[l[f]])));
    }
}

}  // namespace b200
