// solver.cu -- context, device-resident PCG driver and the C ABI of include/b200pcg.h.
//
// Host side of the hot path: replaces PCG::solve's control flow (OF-dev PCG.C),
// lduMatrix::solver::normFactor (lduMatrixSolver.C), the interface update protocol
// (lduMatrixUpdateMatrixInterfaces.C / processorFvPatchField.C) and Pstream's reduce()
// (UPstream.C).  The CG scalars (alpha, beta, residuals, iteration counter, convergence flag)
// live in device memory and are advanced by the kernels themselves; the host only enqueues
// batches of iterations and polls a 200-byte status block, so the loop stops at exactly the
// iteration OpenFOAM's do/while would stop at, with no per-iteration host round trip.
//
// There is NO CPU fallback: every entry point fails with B200_ENODEVICE / B200_ECUDA when the
// device is unusable.
#include "../../include/b200pcg.h"
#include "kernels.cuh"
#include "plan.hpp"

#include <dlfcn.h>
#include <nccl.h>
#include <omp.h>

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace b200;

namespace {

thread_local std::string g_createError;

// NCCL is bound lazily with dlopen, only when a multi-rank context is created.  A hard link
// against libnccl.so.2 would pin whichever copy the loader finds first (the system 2.27 here),
// and a host process that later loads a newer NCCL of the same soname (PyTorch bundles 2.28)
// would then fail to resolve its symbols.  dlopen("libnccl.so.2") returns the copy already in
// the process when there is one.
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool load() {
        if (handle) return true;
        const char* env = getenv("B200_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            error = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
            return false;
        }
#define SYM(field, name)                                                      \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, name));           \
    if (!field) { error = std::string("NCCL symbol missing: ") + name; handle = nullptr; return false; }
        SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank")
        SYM(CommDestroy, "ncclCommDestroy") SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather") SYM(Send, "ncclSend")
        SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
        SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
        return true;
    }
};
NcclApi g_nccl;

enum ProfClass : int {
    PC_FILL = 0, PC_GATHER, PC_SPMV, PC_SPMV_INIT, PC_IFACE, PC_PACK, PC_SUM, PC_NORM, PC_RECIP,
    PC_PRECOND_DOT, PC_DIC_RD, PC_DIC_FWD, PC_DIC_BWD, PC_ASM_FACE,
    PC_ASM_DIAG, PC_FLUX, PC_SCALAR, PC_KP, PC_KR, PC_PSI_FINAL, PC_SMALL, PC_ASM_PRGH,
    PC_EIS_SETUP, PC_EIS_P, PC_EIS_BWD, PC_EIS_FWD, PC_EIS_R, PC_EIS_RES, PC_EIS_ROWS, PC_GS_ROWS, PC_GS_RESID, PC_BICG_TRI, PC_BICG_VEC, PC_BICG_DOT, PC_COUNT
};
const char* kProfNames[PC_COUNT] = {
    "fill_values", "gather_scatter", "spmv_dot", "spmv_init", "iface_fix", "halo_pack", "sum",
    "norm_resid", "recip", "precond_dot", "dic_calc_rd", "dic_fwd",
    "dic_bwd", "asm_face_coeff", "asm_neg_sum_diag", "flux", "scalar_step", "p_psi_update",
    "r_update_dots", "psi_final", "pcg_small_whole_solve", "asm_p_rgh_cells",
    "eis_setup", "eis_p_psi_update", "eis_bwd", "eis_fwd_dot", "eis_r_update_rho", "eis_true_residual",
    "eis_iface_rows", "gs_sweep_rows", "gs_residual", "bicg_dilu_sweeps", "bicg_vector_updates", "bicg_dots"};

struct DevPlan {
    bool built = false;
    HostPlan h;   // big arrays are released after upload; sizes/colourStart stay
    int64_t* sliceBase = nullptr;
    uint32_t* rowLen = nullptr;
    int* col = nullptr;
    int* faceOf = nullptr;
    double* val = nullptr;
    double* valT = nullptr;       // PBiCG: entries of the TRANSPOSED matrix (the other coefficient of each face), on first use
    int* perm = nullptr;
    int *slotRow = nullptr, *bRow = nullptr, *bStart = nullptr, *bSlot = nullptr;
    // 16-bit column offsets (plan.hpp colBase / col16) for the ELL-bound kernels
    uint16_t* col16 = nullptr;
    int* colBase = nullptr;
    bool c16 = false;
    int* colourStart = nullptr;   // device copy of h.colourStart (k_pcg_small)
    int* segStart = nullptr;      // device copy of h.segStart ((tile, colour) row segments)
    int tileShift = 31;           // log2(tileRows) of a tiled multicolour plan, else 31
    int maxRowLen = 0;            // widest row (faces per cell)
    int nSlots = 0;
    // symmetric single-read layout (SymPlan)
    bool sym = false;
    int64_t symNU = 0;
    int symWU = 0, symWL = 0;
    bool symTma = false;
    size_t symStage = 0;
    uint32_t* sRowLen = nullptr;
    uint8_t* sRowLen8 = nullptr;  // nL | nU << 4 per row (bulk-copy staged kernel)
    int *sUCol = nullptr, *sUFace = nullptr;
    double* sUVal = nullptr;
    uint32_t* sLRef = nullptr;
    // interface rows of every CTA of the staged Amul's grid (kernels.cuh IfaceTail), for grid size tailGrid
    int *ctaBStart = nullptr, *ctaB = nullptr;
    int tailGrid = 0, tailUnitRows = 0;
    // slots of every CTA of the k_p launch (kernels.cuh PackTail), for grid size packGrid
    int *ctaSStart = nullptr, *ctaS = nullptr;
    int packGrid = 0;
    // single-read face-ordered layout of renumbered natural plans (plan.hpp SrPlan)
    bool sr = false;
    int64_t srNOwn = 0;
    uint32_t* srMeta = nullptr;
    int64_t* srOwnBase = nullptr;
    int* srOwnFace = nullptr;
    double* srOwnVal = nullptr;
    // Eisenstat form, nranks > 1: interface-row index of every row (-1: none) and the halo term B t
    int* rowB = nullptr;
    double* hb = nullptr;
    double* hbv = nullptr;        // smoothSolver, nranks > 1: bPrime of the interface rows (k_gs_bprime)
    int nB0 = 0;                  // interface rows of the first colour: bRow[0 .. nB0) (bRow is ascending)
    // CUDA graphs of kGraphIters loop bodies, by loop form (0 none, 1 diagonal, 2 DIC-class loop, 3 Eisenstat):
    // kernel arguments are the context's own buffers and the CG scalars live on the device, so one capture
    // serves every later solve on this plan
    cudaGraphExec_t iterGraph[4] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t iterGraphLaunches[4] = {0, 0, 0, 0};
    bool iterGraphFailed[4] = {false, false, false, false};
};

struct HostIface {
    int32_t nbrRank;
    std::vector<int32_t> faceCells;
};

}  // namespace

struct b200_ctx {
    int device = 0, rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    cudaStream_t sc = nullptr, sm = nullptr;
    cudaEvent_t evPack = nullptr, evRecv = nullptr, ev[6] = {};
    std::string err;
    int numSMs = 148;
    // mesh
    bool haveMesh = false;
    uint64_t meshKey = 0;
    uint64_t meshFingerprint = 0;   // mesh_fingerprint() of the addressing the plans were built from
    int32_t N = 0, F = 0;
    double nGlobalCells = 0;
    std::vector<int32_t> hl, hu;
    std::vector<HostIface> hif;
    int *d_l = nullptr, *d_u = nullptr;
    // natural-order face lists (lduAddressing's ownerStart / losortStart / losort), built only when the plan's
    // rows are renumbered: negSumDiag then runs in the caller's cell order (k_neg_sum_diag_nat)
    int *d_ownerStart = nullptr, *d_losortStart = nullptr, *d_losort = nullptr;
    DevPlan plans[3];
    int nSlots = 0;
    // vectors (internal order)
    double *diag = nullptr, *src = nullptr, *psi = nullptr, *r = nullptr, *p = nullptr,
           *w = nullptr, *rD = nullptr;
    double *t = nullptr, *dT = nullptr, *eD = nullptr;   // Eisenstat form: t, D~, D - 2 D~ (allocated on first use)
    double *bou = nullptr, *sendbuf = nullptr, *recvbuf = nullptr;
    double *ifaceProd = nullptr, *partials2 = nullptr;   // k_iface_pre: products per patch face (CSR order), its partials
    // staging for the host entry points (natural order)
    double *in_diag = nullptr, *in_upper = nullptr, *in_src = nullptr, *in_psi = nullptr,
           *in_bou = nullptr, *in_f1 = nullptr, *in_f2 = nullptr, *in_f3 = nullptr;
    double* in_lower = nullptr;      // asymmetric matrices (b200_smooth_solve / b200_amul_asym), allocated on first use
    // smoothSolver
    bool forceEll = false;           // an asymmetric matrix is loaded: Amul must take the full-row ELL (the single-read
                                     // layouts store every coefficient once, i.e. assume lower == upper)
    int gsSweeps = 1;                // |nSweeps| of the current smoothSolver call (Scalars::nSweeps)
    int gsCtas = 0;                  // B200PCG_GS_CTAS=3|4: register build of the narrow-row smoothSolver kernels (0: default 3)
    bool gsLagged = true;            // B200PCG_GS_LAGGED=0: two-colour plans evaluate the residual with a separate kernel (A/B switch)
    // boundary faces (b200_set_boundary_faces): CSR cell -> boundary faces in patch order
    int32_t nB = 0;
    std::vector<int32_t> hbCells;
    int *bfStart = nullptr, *bfOrder = nullptr;
    double* scratch = nullptr;      // staging arena of the host entry point b200_assemble_p_rgh
    size_t scratchElems = 0;
    // one-shot peer-memory all-reduce (peer_allreduce_step in the reducing kernels' last block)
    bool p2pReduce = false;
    PeerBuf* peerLocal = nullptr;
    std::vector<void*> peerMapped;     // cudaIpcOpenMemHandle results (to close)
    PeerBuf** d_peers = nullptr;
    // processor-patch halos over peer memory (kernels.cuh Halo): this rank's receive buffer (vals[2][nSlots] +
    // flags[2][kMaxRanks], exported with CUDA IPC), the neighbours' buffers mapped here, the device tables
    bool p2pHalo = false;
    bool forceNcclHalo = false;        // B200PCG_HALO=nccl
    double* haloLocal = nullptr;
    std::vector<void*> haloMapped;
    double** d_haloDst = nullptr;
    unsigned long long** d_haloNbrFlag = nullptr;
    int* d_haloNbrRanks = nullptr;
    Halo haloDev = {nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
    // staged copies of pageable caller memory (B200PCG_STAGED_COPY=1): two page-locked pieces + their DMA events
    bool stagedCopy = true;
    void* stageBuf[2] = {nullptr, nullptr};
    cudaEvent_t stageEv[2] = {nullptr, nullptr};
    Scalars* S = nullptr;
    Scalars* hS = nullptr;  // pinned
    double* partials = nullptr;
    uint64_t launches = 0;
    uint64_t graphLaunches = 0;     // cudaGraphLaunch calls (each replays kGraphIters loop bodies)
    int32_t forceIters = 0;
    // bulk-copy pipeline depth / CTAs per SM (B200PCG_STAGES, B200PCG_CTAS).  Measured on the 16 M hex
    // box (profiles/r01_tma_sweep.md): 2 stages x 4 CTAs/SM is the optimum -- deeper pipelines or more
    // CTAs shrink the L1 carve-out that serves the neighbour gathers.
    int symStages = 2, symPerSM = 0;
    // single-launch cluster kernel for small systems (k_pcg_small): up to this many cells
    // (B200PCG_SMALL_N, 0 disables), cluster size 8 or 16 (B200PCG_SMALL_CTAS)
    int tileRows = 0;             // B200PCG_TILE: rows per tile of large multicolour plans (0: colour-major).
                                  // Opt-in: measured slower (profiles/r01_v7_dic_tiles.md) -- in a red-black order rows are
                                  // all-upper or all-lower, so the uniform-width single-read layout is half padding
    int smallN = 150000, smallCtas = 16;
    bool usedSmall = false, usedFast = false;
    bool disableFast = false;     // B200PCG_SMALL_FAST=0: always the L2-resident k_pcg_small
    int fastMaxCtas = 16;         // B200PCG_FAST_CTAS: largest cluster the on-chip kernel may use
    int renumber = (int)Renumber::Auto;   // B200PCG_RENUMBER=0|1|auto: RCM base order (plan.hpp)
    bool disableCol16 = false;  // B200PCG_COL16=0: always 32-bit columns in the full-row ELL kernels
    bool sortCols = false;      // B200PCG_SORT_COLS=1: multicolour plans order a row's entries by column (plan.hpp)
    bool dicDefaultEis = true;  // B200PCG_DIC=multicolour: code 2 (`preconditioner DIC`) keeps the three-kernel loop
    bool eisOverlap = true;     // B200PCG_EIS_OVERLAP=0: nranks > 1: exchange t exposed between the two sweeps (A/B switch)
    int eisCtas = 0;            // B200PCG_EIS_CTAS=3|4: force the 80- / 64-register build of both 6-entry batched
                                // sweeps (default 0: backward 64, forward 80 registers)
    int eisBatch = 1;           // B200PCG_EIS_BATCH=0: plain entry loops in the Eisenstat sweeps (A/B switch)
    int sweepPerSM = 8;         // B200PCG_SWEEP_CTAS: CTAs per SM of the colour sweeps (DIC-class, Eisenstat form)
    bool sweepPerSMSet = false;
    bool noFuseFirst = false;   // B200PCG_FUSE_FIRST=0: keep the first colour's forward sweep a separate launch
    bool exactWidth = true;     // B200PCG_EXACT=0: always use the 4+4-slot generic instantiation
    bool disableTma = false;    // B200PCG_SPMV=sym: symmetric layout with direct loads (no bulk-copy staging)
    bool disableSym = false;    // B200PCG_SPMV=ell: keep the full-row sliced-ELL Amul (A/B switch)
    bool enableSr = false;      // B200PCG_SPMV=sr: renumbered natural plans use the single-read face-ordered layout
                                // (k_spmv_sr).  Opt-in: it moves 0.73x the DRAM bytes of the full-row ELL but its
                                // dependent meta -> ownBase -> value chain and the extra 48 M gather sectors make it
                                // latency-bound: 307 vs 185 us on the 5 M-cell polyhedral workload
                                // (profiles/r02_ncu_full_poly_sr.md)
    // profiling
    bool useGraph = true;       // B200PCG_GRAPH=0: enqueue every loop body kernel by kernel
    bool fuseIface = false;     // B200PCG_FUSE_IFACE=1: N > 1, peer-memory halos: k_p packs in its tail and the Amul corrects
                                // its own interface rows in its tail (three launches per iteration, one stream).
                                // Opt-in: measured EQUAL to the separate pack / fix-up kernels at 2 GPUs (437.7 vs
                                // 435.5 us per iteration): the tails add 12 us to k_p and 18 us to the Amul -- the
                                // cost is the chain of dependent loads and the system-scope fence, not the launches
    bool splitIface = false;    // B200PCG_SPLIT_IFACE=1: k_iface_pre (comm stream, concurrent with the Amul) + a short
                                // k_iface_apply instead of one k_iface_fix behind the Amul.  Opt-in: measured EQUAL at 2
                                // and 8 GPUs (467.3 vs 468.2 us per iteration at 8): the fix-up gets 17 us shorter, the
                                // Amul 16 us longer (it shares the SMs with the pre kernel)
    bool graphMulti = false;    // B200PCG_GRAPH_MULTI=1: iteration graphs with nranks > 1 on single-stream loop bodies
    bool prof = false;
    bool profOpen = false;
    struct ProfRec { int cls; int iter; cudaEvent_t a, b; };
    int profIter = 0;            // loop body being enqueued (0: set-up / tail); see prof_collect
    std::vector<ProfRec> profRecs;
    std::vector<cudaEvent_t> profPool;
    size_t profUsed = 0;
    uint64_t profSkipped = 0;    // launches of surplus iterations (returned on S->done), excluded from the averages
    double profMs[PC_COUNT] = {};
    uint64_t profN[PC_COUNT] = {};
    std::string profJson;
};

namespace {

int fail(b200_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_createError = msg;
    return code;
}

#define CU(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail(ctx, B200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define NC(call)                                                                         \
    do {                                                                                 \
        ncclResult_t e_ = (call);                                                        \
        if (e_ != ncclSuccess)                                                           \
            return fail(ctx, B200_ENCCL, std::string(#call) + ": " + g_nccl.GetErrorString(e_)); \
    } while (0)
#define RET(call)                        \
    do {                                 \
        int rc_ = (call);                \
        if (rc_ != B200_OK) return rc_;  \
    } while (0)

template <class T>
int dev_alloc(b200_ctx* ctx, T** p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    CU(cudaMalloc((void**)p, n * sizeof(T)));
    return B200_OK;
}
template <class T>
void dev_free(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}
template <class T>
int upload(b200_ctx* ctx, T** d, const std::vector<T>& h) {
    RET(dev_alloc(ctx, d, h.size()));
    if (!h.empty())
        CU(cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->sc));
    return B200_OK;
}

void prof_begin(b200_ctx* c, int cls) {
    c->profOpen = false;
    if (!c->prof || c->profUsed + 2 > c->profPool.size()) return;
    c->profOpen = true;
    b200_ctx::ProfRec r{cls, c->profIter, c->profPool[c->profUsed], c->profPool[c->profUsed + 1]};
    c->profUsed += 2;
    cudaEventRecord(r.a, c->sc);
    c->profRecs.push_back(r);
}
void prof_end(b200_ctx* c, int cls) {
    if (!c->prof || !c->profOpen || c->profRecs.empty()) return;
    c->profOpen = false;
    auto& r = c->profRecs.back();
    if (r.cls == cls && r.b) cudaEventRecord(r.b, c->sc);
}
// nIterDone: loop bodies the device actually executed.  The host enqueues iterations in batches and the
// kernels of the surplus ones return at once on S->done: those launches are NOT part of the averages.
void prof_collect(b200_ctx* c, int nIterDone = 0x7fffffff) {
    if (!c->prof) return;
    cudaStreamSynchronize(c->sc);
    for (auto& r : c->profRecs) {
        if (r.iter > nIterDone) { c->profSkipped++; continue; }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            c->profMs[r.cls] += ms;
            c->profN[r.cls]++;
        }
    }
    c->profRecs.clear();
    c->profUsed = 0;
}

#define LAUNCH(cls, kern, grid, ...)                             \
    do {                                                         \
        prof_begin(ctx, cls);                                    \
        kern<<<(grid), kBlock, 0, ctx->sc>>>(__VA_ARGS__);       \
        prof_end(ctx, cls);                                      \
        ctx->launches++;                                         \
    } while (0)

inline int grid_for(const b200_ctx* c, int64_t items, int perSM = 8) {
    int64_t g = (items + kBlock - 1) / kBlock;
    int64_t cap = (int64_t)c->numSMs * perSM;
    if (cap > kMaxGrid) cap = kMaxGrid;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

void free_plan(DevPlan& P) {
    for (int m = 0; m < 4; ++m) {
        if (P.iterGraph[m]) cudaGraphExecDestroy(P.iterGraph[m]);
        P.iterGraph[m] = nullptr;
        P.iterGraphFailed[m] = false;
    }
    dev_free(P.sliceBase); dev_free(P.rowLen); dev_free(P.col); dev_free(P.faceOf);
    dev_free(P.val); dev_free(P.valT); dev_free(P.perm); dev_free(P.slotRow); dev_free(P.bRow);
    dev_free(P.bStart); dev_free(P.bSlot); dev_free(P.colourStart); dev_free(P.segStart);
    dev_free(P.col16); dev_free(P.colBase); P.c16 = false;
    dev_free(P.sUCol); dev_free(P.sUFace);
    dev_free(P.sUVal); dev_free(P.sLRef); dev_free(P.sRowLen); dev_free(P.sRowLen8);
    dev_free(P.rowB); dev_free(P.hb); dev_free(P.hbv);
    dev_free(P.srMeta); dev_free(P.srOwnBase); dev_free(P.srOwnFace); dev_free(P.srOwnVal);
    dev_free(P.ctaBStart); dev_free(P.ctaB); P.tailGrid = 0;
    dev_free(P.ctaSStart); dev_free(P.ctaS); P.packGrid = 0;
    P.sr = false;
    P.sym = false;
    P.built = false;
    P.h = HostPlan();
}

void free_halo(b200_ctx* c) {
    for (void* p : c->haloMapped) cudaIpcCloseMemHandle(p);
    c->haloMapped.clear();
    dev_free(c->haloLocal); dev_free(c->d_haloDst); dev_free(c->d_haloNbrFlag); dev_free(c->d_haloNbrRanks);
    c->haloDev = Halo{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
    c->p2pHalo = false;
}

void free_mesh(b200_ctx* c) {
    free_halo(c);
    for (auto& P : c->plans) free_plan(P);
    dev_free(c->d_l); dev_free(c->d_u);
    dev_free(c->d_ownerStart); dev_free(c->d_losortStart); dev_free(c->d_losort);
    dev_free(c->diag); dev_free(c->src); dev_free(c->psi); dev_free(c->r); dev_free(c->p);
    dev_free(c->w); dev_free(c->rD); dev_free(c->t); dev_free(c->dT); dev_free(c->eD); dev_free(c->bou); dev_free(c->sendbuf); dev_free(c->recvbuf); dev_free(c->ifaceProd);
    dev_free(c->in_diag); dev_free(c->in_upper); dev_free(c->in_src); dev_free(c->in_psi);
    dev_free(c->in_bou); dev_free(c->in_f1); dev_free(c->in_f2); dev_free(c->in_f3); dev_free(c->in_lower);
    dev_free(c->bfStart); dev_free(c->bfOrder); dev_free(c->scratch);
    c->nB = 0; c->hbCells.clear(); c->scratchElems = 0;
    c->haveMesh = false;
    c->hl.clear(); c->hu.clear(); c->hif.clear();
}

int ensure_plan(b200_ctx* ctx, Ordering ord, DevPlan** out) {
    DevPlan& P = ctx->plans[(int)ord];
    *out = &P;
    if (P.built) return B200_OK;
    std::vector<IfaceIn> ifs(ctx->hif.size());
    for (size_t k = 0; k < ifs.size(); ++k)
        ifs[k] = IfaceIn{ctx->hif[k].nbrRank, (int32_t)ctx->hif[k].faceCells.size(),
                         ctx->hif[k].faceCells.data()};
    // large multicolour plans are tiled (plan.hpp); small ones stay colour-major (cluster kernels)
    const int32_t tile = (ord == Ordering::MultiColour && ctx->tileRows > 0 &&
                          ctx->N > std::max(ctx->nranks == 1 ? ctx->smallN : 0, 4 * ctx->tileRows)) ? ctx->tileRows : 0;
    std::string e = build_plan(ord, ctx->N, ctx->F, ctx->hl.data(), ctx->hu.data(),
                               (int32_t)ifs.size(), ifs.data(), P.h, (Renumber)ctx->renumber, tile, ctx->sortCols);
    if (!e.empty()) return fail(ctx, B200_EINVAL, "set_addressing: " + e);
    RET(upload(ctx, &P.sliceBase, P.h.sliceBase));
    RET(upload(ctx, &P.rowLen, P.h.rowLen));
    RET(upload(ctx, &P.col, P.h.col));
    RET(upload(ctx, &P.faceOf, P.h.faceOf));
    RET(dev_alloc(ctx, &P.val, (size_t)P.h.nEntries));
    if (!P.h.perm.empty()) RET(upload(ctx, &P.perm, P.h.perm));
    RET(upload(ctx, &P.slotRow, P.h.slotRow));
    RET(upload(ctx, &P.bRow, P.h.bRow));
    RET(upload(ctx, &P.bStart, P.h.bStart));
    RET(upload(ctx, &P.bSlot, P.h.bSlot));
    if (!P.h.colBase.empty() && P.h.col16Fraction == 1.0 && !ctx->disableCol16) {   // all-or-nothing (kernels.cuh)
        RET(upload(ctx, &P.col16, P.h.col16));
        RET(upload(ctx, &P.colBase, P.h.colBase));
        P.c16 = true;
    }
    RET(upload(ctx, &P.colourStart, P.h.colourStart));
    if (P.h.segStart.empty()) P.h.segStart = P.h.colourStart;   // Natural: one segment
    RET(upload(ctx, &P.segStart, P.h.segStart));
    P.tileShift = 31;
    if (P.h.nTiles > 1) {
        P.tileShift = 0;
        while ((1 << P.tileShift) < P.h.tileRows) ++P.tileShift;
    }
    P.maxRowLen = 0;
    for (size_t sl = 0; sl + 1 < P.h.sliceBase.size(); ++sl)
        P.maxRowLen = std::max(P.maxRowLen, (int)((P.h.sliceBase[sl + 1] - P.h.sliceBase[sl]) / 32));
    P.nSlots = (int)P.h.slotRow.size();
    // permuted (colour-major) orders put a row's earlier neighbours hundreds of MB upstream: the
    // re-read misses L2, so those plans keep the full-row sliced ELL for Amul
    // the symmetric single-read Amul needs a row's earlier neighbours close upstream (L2 hits): natural
    // order, or a multicolour order that is tiled (or small enough to be L2-resident anyway)
    // (renumbered natural plans take the face-ordered single-read layout below -- SrPlan -- instead; the
    // single-read layout on tiled multicolour plans, B200PCG_TILE=<rows>, is opt-in: it measured slower than the
    // full-row ELL it replaces, profiles/r01_v7_dic_tiles.md)
    const bool symOrder = (ord == Ordering::Natural && !P.h.renumbered) ||
                          (ord == Ordering::MultiColour && P.h.nTiles > 1);
    if (P.h.sym.valid && !ctx->disableSym && symOrder) {
        SymPlan& Y = P.h.sym;
        // arrays padded to whole 256-row chunks for the bulk-copy (TMA) staging
        const size_t rowsPad = (((size_t)ctx->N + kChunkRows - 1) / kChunkRows) * kChunkRows;
        Y.uCol.resize(rowsPad * Y.WU, 0);
        Y.uFace.resize(rowsPad * Y.WU, -1);
        Y.lRef.resize(rowsPad * Y.WL, 0);
        RET(upload(ctx, &P.sUCol, Y.uCol));
        RET(upload(ctx, &P.sUFace, Y.uFace));
        RET(upload(ctx, &P.sLRef, Y.lRef));
        RET(dev_alloc(ctx, &P.sUVal, rowsPad * Y.WU));
        {
            // the layout's own row lengths (split by row index), padded: the staged kernel reads whole chunks
            std::vector<uint32_t> rl(rowsPad, 0u);
            std::copy(Y.rowLen.begin(), Y.rowLen.end(), rl.begin());
            RET(upload(ctx, &P.sRowLen, rl));
            if (Y.WU <= 15 && Y.WL <= 15) {
                std::vector<uint8_t> rl8(rowsPad, 0);
                for (size_t r = 0; r < Y.rowLen.size(); ++r) {
                    const uint32_t nLo = Y.rowLen[r] & 0xffffu, nUp = (Y.rowLen[r] >> 16) - nLo;
                    rl8[r] = (uint8_t)(nLo | (nUp << 4));
                }
                RET(upload(ctx, &P.sRowLen8, rl8));
            }
            CU(cudaStreamSynchronize(ctx->sc));
        }
        P.symNU = (int64_t)(rowsPad * Y.WU);
        P.symStage = sym_stage_bytes(Y.WU, Y.WL);
        P.symTma = !ctx->disableTma && P.sRowLen8 && (128 + ctx->symStages * P.symStage) <= 48 * 1024;
        if (P.symTma) {
#define B200_TMA_ATTR(...) CU(cudaFuncSetAttribute(k_spmv_sym_tma<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024))
            B200_TMA_ATTR(true, 2, 4, 4, true); B200_TMA_ATTR(false, 2, 4, 4, true);
            B200_TMA_ATTR(true, 3, 4, 4, true); B200_TMA_ATTR(false, 3, 4, 4, true);
            B200_TMA_ATTR(true, 2, 3, 3, false); B200_TMA_ATTR(false, 2, 3, 3, false);
            B200_TMA_ATTR(true, 2, 2, 2, false); B200_TMA_ATTR(false, 2, 2, 2, false);
#undef B200_TMA_ATTR
        }
        P.symWU = Y.WU;
        P.symWL = Y.WL;
        P.sym = true;
    }
    if (P.h.sr.valid && ctx->enableSr && !P.sym) {
        RET(upload(ctx, &P.srMeta, P.h.sr.meta));
        RET(upload(ctx, &P.srOwnBase, P.h.sr.ownBase));
        RET(upload(ctx, &P.srOwnFace, P.h.sr.ownFace));
        RET(dev_alloc(ctx, &P.srOwnVal, (size_t)P.h.sr.nOwn));
        P.srNOwn = P.h.sr.nOwn;
        P.sr = true;
    }
    CU(cudaStreamSynchronize(ctx->sc));
    P.h.sym = SymPlan();
    P.h.sr = SrPlan();

    // release the big host arrays
    std::vector<int32_t>().swap(P.h.col);
    std::vector<uint16_t>().swap(P.h.col16);
    std::vector<int32_t>().swap(P.h.colBase);
    std::vector<int32_t>().swap(P.h.faceOf);
    std::vector<uint32_t>().swap(P.h.rowLen);
    std::vector<int64_t>().swap(P.h.sliceBase);
    std::vector<int32_t>().swap(P.h.perm);
    std::vector<int32_t>().swap(P.h.iperm);
    std::vector<int32_t>().swap(P.h.bSlot);
    P.built = true;
    return B200_OK;
}

int reset_scalars(b200_ctx* ctx, const b200_controls* ctl) {
    Scalars& h = *ctx->hS;
    std::memset(&h, 0, sizeof(Scalars));
    h.tol = ctl ? ctl->tolerance : 0.0;
    h.relTol = ctl ? ctl->relTol : 0.0;
    h.maxIter = ctl ? ctl->maxIter : 0;
    h.minIter = ctl ? ctl->minIter : 0;
    h.forceIters = ctx->forceIters;
    h.nranks = ctx->nranks;
    h.nSweeps = ctx->gsSweeps;
    h.nGlobalCells = ctx->nGlobalCells;
    h.wArA = 1e20;
    h.wArAold = 1e20;
    // everything but redSeq: the count of executed cross-rank reductions lives on the device for the
    // lifetime of the context (kernels.cuh Scalars)
    CU(cudaMemcpyAsync(ctx->S, ctx->hS, offsetof(Scalars, redSeq), cudaMemcpyHostToDevice, ctx->sc));
    return B200_OK;
}

// Reduction descriptor of a reducing kernel.  With nranks > 1 and peer-memory reductions, the
// kernel that completes the reduction (step != STEP_NONE) also performs the cross-rank sum and the
// scalar step in its last block: it gets the peer table and the next sequence number.
Reduce mkR(b200_ctx* ctx, int step) {
    Reduce R{ctx->S, ctx->partials, step, nullptr, ctx->rank};
    if (step != STEP_NONE && ctx->nranks > 1 && ctx->p2pReduce) R.peers = ctx->d_peers;
    return R;
}

// after a reducing kernel with a step: all-reduce + scalar step when nranks > 1 (nothing to do when
// the kernel's last block already did both over peer memory)
int reduce_post(b200_ctx* ctx, int step) {
    if (ctx->nranks == 1 || ctx->p2pReduce) return B200_OK;
    NC(g_nccl.AllReduce(ctx->S->sums, ctx->S->gsums, kNSums, ncclDouble, ncclSum, ctx->comm, ctx->sc));
    LAUNCH(PC_SCALAR, k_scalar_step, 1, ctx->S, step);
    return B200_OK;
}

// Start the exchange of x[faceCells] across the processor patches on `st` (the data is complete on `st`):
// peer-memory form = ONE small kernel that stores straight into the neighbours' receive buffers and raises
// their flags (kernels.cuh k_pack_p2p); NCCL form = pack + grouped ncclSend/ncclRecv.  The consumer kernels
// take ctx->haloDev (+ ctx->recvbuf for the NCCL form) and wait for the flags themselves.
int halo_exchange(b200_ctx* ctx, DevPlan& P, const double* x, cudaStream_t st, bool forceNccl = false) {
    if (ctx->p2pHalo && !forceNccl) {
        k_pack_p2p<<<grid_for(ctx, P.nSlots), kBlock, 0, st>>>(ctx->haloDev, P.slotRow, x, ctx->S);
        ctx->launches++;
        return B200_OK;
    }
    k_pack<<<grid_for(ctx, P.nSlots), kBlock, 0, st>>>(P.nSlots, P.slotRow, x, ctx->sendbuf, ctx->S);
    ctx->launches++;
    NC(g_nccl.GroupStart());
    for (int k = 0; k < P.h.nIfaces; ++k) {
        const int off = P.h.patchStart[k], n = P.h.patchStart[k + 1] - off;
        if (n == 0) continue;
        NC(g_nccl.Send(ctx->sendbuf + off, n, ncclDouble, P.h.nbrRank[k], ctx->comm, st));
        NC(g_nccl.Recv(ctx->recvbuf + off, n, ncclDouble, P.h.nbrRank[k], ctx->comm, st));
    }
    NC(g_nccl.GroupEnd());
    return B200_OK;
}

// interface rows of every CTA of an Amul launch.  Staged kernel: chunk c of 256 rows is computed by CTA c mod grid
// (unitRows 256, unitsPerCta 1); full-row ELL kernel: slice s of 32 rows by warp s mod (8 grid), CTA = warp / 8
// (unitRows 32, unitsPerCta 8).
int ensure_cta_lists(b200_ctx* ctx, DevPlan& P, int grid, int unitRows, int unitsPerCta) {
    if (P.ctaBStart && P.tailGrid == grid && P.tailUnitRows == unitRows) return B200_OK;
    auto ctaOf = [&](int32_t r) { return (int)(((int64_t)(r / unitRows) % ((int64_t)grid * unitsPerCta)) / unitsPerCta); };
    CU(cudaStreamSynchronize(ctx->sc));
    dev_free(P.ctaBStart);
    dev_free(P.ctaB);
    std::vector<int32_t> start((size_t)grid + 1, 0), list((size_t)P.h.nBRows);
    for (int b = 0; b < P.h.nBRows; ++b) start[(size_t)ctaOf(P.h.bRow[(size_t)b]) + 1]++;
    for (int g = 0; g < grid; ++g) start[(size_t)g + 1] += start[(size_t)g];
    std::vector<int32_t> pos(start.begin(), start.end() - 1);
    for (int b = 0; b < P.h.nBRows; ++b) list[(size_t)pos[(size_t)ctaOf(P.h.bRow[(size_t)b])]++] = b;
    RET(upload(ctx, &P.ctaBStart, start));
    RET(upload(ctx, &P.ctaB, list));
    CU(cudaStreamSynchronize(ctx->sc));
    P.tailGrid = grid;
    P.tailUnitRows = unitRows;
    return B200_OK;
}

// patch-face slots of every CTA of a k_p launch (B200_VEC_LOOP: double2 i = rows 2i, 2i+1 belongs to CTA
// (i / 256) mod grid; the odd last row to CTA 0)
int ensure_pack_lists(b200_ctx* ctx, DevPlan& P, int grid) {
    if (P.ctaSStart && P.packGrid == grid) return B200_OK;
    const int N = ctx->N;
    auto ctaOf = [&](int32_t r) { return ((N & 1) && r == N - 1) ? 0 : (int)(((r >> 1) / kBlock) % grid); };
    CU(cudaStreamSynchronize(ctx->sc));
    dev_free(P.ctaSStart);
    dev_free(P.ctaS);
    const size_t nS = P.h.slotRow.size();
    std::vector<int32_t> start((size_t)grid + 1, 0), list(nS);
    for (size_t i = 0; i < nS; ++i) start[(size_t)ctaOf(P.h.slotRow[i]) + 1]++;
    for (int g = 0; g < grid; ++g) start[(size_t)g + 1] += start[(size_t)g];
    std::vector<int32_t> pos(start.begin(), start.end() - 1);
    for (size_t i = 0; i < nS; ++i) list[(size_t)pos[(size_t)ctaOf(P.h.slotRow[i])]++] = (int32_t)i;
    RET(upload(ctx, &P.ctaSStart, start));
    RET(upload(ctx, &P.ctaS, list));
    CU(cudaStreamSynchronize(ctx->sc));
    P.packGrid = grid;
    return B200_OK;
}

// can this plan's Amul correct its own interface rows (kernels.cuh IfaceTail)?
bool amul_fuses_iface(const b200_ctx* ctx, const DevPlan& P) {
    return ctx->nranks > 1 && P.nSlots > 0 && ctx->p2pHalo && ctx->fuseIface &&
           (ctx->forceEll || (P.sym && P.symTma) || (!P.sym && !P.sr));
}

// y = A x (+ interfaces) [+ (y,x) -> step]; INIT: also sA = sumA
// packed: the producer of x has already stored its patch-face values into the neighbours (k_p's fused pack tail)
template <bool INIT, bool DOT>
int spmv_full(b200_ctx* ctx, DevPlan& P, const double* x, double* y, double* sA, int step, bool packed = false) {
    const int N = ctx->N;
    const bool halo = (ctx->nranks > 1 && P.nSlots > 0);
    // peer-memory halos + the staged Amul: the pack kernel runs on the MAIN stream ahead of the Amul (a few us: it
    // only stores; the neighbours' values arrive while the Amul computes) and the Amul corrects its own interface
    // rows in its tail -- one stream, no events, no separate fix-up kernel, one reduction instead of two
    const bool fuseTail = !INIT && amul_fuses_iface(ctx, P);
    const bool splitFix = halo && !fuseTail && !INIT && ctx->splitIface;
    if (halo && fuseTail) {
        if (!packed) RET(halo_exchange(ctx, P, x, ctx->sc));
    } else if (halo) {
        // the exchange runs on the comm stream, concurrently with the Amul
        CU(cudaEventRecord(ctx->evPack, ctx->sc));
        CU(cudaStreamWaitEvent(ctx->sm, ctx->evPack, 0));
        RET(halo_exchange(ctx, P, x, ctx->sm));
        if (splitFix) {
            // products + dot correction on the comm stream, concurrent with the Amul (kernels.cuh k_iface_pre)
            auto pre = k_iface_pre<DOT>;
            pre<<<grid_for(ctx, P.h.nBRows), kBlock, 0, ctx->sm>>>(P.h.nBRows, P.bRow, P.bStart, P.bSlot, ctx->bou,
                                                                   ctx->recvbuf, ctx->haloDev, x, ctx->ifaceProd,
                                                                   ctx->partials2, ctx->S);
            ctx->launches++;
        }
        CU(cudaEventRecord(ctx->evRecv, ctx->sm));
    }
    Reduce R = mkR(ctx, (halo && !fuseTail) ? STEP_NONE : step);
    if (P.sym && P.symTma && !INIT && !ctx->forceEll) {
        const size_t smem = 128 + ctx->symStages * P.symStage;
        int perSM = (int)((size_t)144 * 1024 / (smem + 1024));   // leave >= 80 KB of L1 for the gathers
        if (perSM > 8) perSM = 8;
        if (perSM < 1) perSM = 1;
        if (ctx->symPerSM > 0) perSM = std::min(ctx->symPerSM, (int)((size_t)224 * 1024 / (smem + 1024)));
        const int nChunks = (N + kChunkRows - 1) / kChunkRows;
        int g = std::min(nChunks, perSM * ctx->numSMs);
        if (g > kMaxGrid) g = kMaxGrid;
        IfaceTail T = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ctx->haloDev};
        if (fuseTail) {
            RET(ensure_cta_lists(ctx, P, g, kChunkRows, 1));
            T = IfaceTail{P.ctaBStart, P.ctaB, P.bRow, P.bStart, P.bSlot, ctx->bou, ctx->haloDev};
        }
        prof_begin(ctx, PC_SPMV);
#define B200_TMA_LAUNCH(...) k_spmv_sym_tma<DOT, __VA_ARGS__><<<g, kBlock, smem, ctx->sc>>>(N, P.symWU, P.symWL, P.sRowLen8, P.sUCol, \
                                                                                   P.sUVal, P.sLRef, ctx->diag, x, y, R, T)
        // exact-width instantiations (no tail loops, no idle slots) for the common narrow rows
        if (ctx->symStages == 3) B200_TMA_LAUNCH(3, 4, 4, true);
        else if (ctx->exactWidth && P.symWU <= 2 && P.symWL <= 2) B200_TMA_LAUNCH(2, 2, 2, false);
        else if (ctx->exactWidth && P.symWU <= 3 && P.symWL <= 3) B200_TMA_LAUNCH(2, 3, 3, false);
        else B200_TMA_LAUNCH(2, 4, 4, true);
#undef B200_TMA_LAUNCH
        prof_end(ctx, PC_SPMV);
        ctx->launches++;
    } else if (P.sr && !ctx->forceEll) {
        auto kern = k_spmv_sr<INIT, DOT>;
        LAUNCH(INIT ? PC_SPMV_INIT : PC_SPMV, kern, grid_for(ctx, N), N, P.sliceBase, P.rowLen, P.srMeta, P.srOwnBase,
               P.srOwnVal, ctx->diag, x, y, sA, R);
    } else if (P.sym && !ctx->forceEll) {
        auto kern = k_spmv_sym<INIT, DOT>;
        LAUNCH(INIT ? PC_SPMV_INIT : PC_SPMV, kern, grid_for(ctx, N, 4), N, P.symWU, P.symWL, P.sRowLen,
               P.sUCol, P.sUVal, P.sLRef, ctx->diag, x, y, sA, R);
    } else {
        const EllCols E{P.col, P.col16, P.colBase};
        const int g = grid_for(ctx, N);
        IfaceTail T = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ctx->haloDev};
        if (fuseTail) {
            RET(ensure_cta_lists(ctx, P, g, 32, kBlock / 32));
            T = IfaceTail{P.ctaBStart, P.ctaB, P.bRow, P.bStart, P.bSlot, ctx->bou, ctx->haloDev};
        }
        if (P.c16) {
            auto kern = k_spmv<INIT, DOT, true>;
            LAUNCH(INIT ? PC_SPMV_INIT : PC_SPMV, kern, g, N, P.sliceBase, P.rowLen, E, P.val,
                   ctx->diag, x, y, sA, R, T);
        } else {
            auto kern = k_spmv<INIT, DOT, false>;
            LAUNCH(INIT ? PC_SPMV_INIT : PC_SPMV, kern, g, N, P.sliceBase, P.rowLen, E, P.val,
                   ctx->diag, x, y, sA, R, T);
        }
    }
    if (halo && !fuseTail) {
        CU(cudaStreamWaitEvent(ctx->sc, ctx->evRecv, 0));
        Reduce R2 = mkR(ctx, step);
        if (splitFix) {
            auto app = k_iface_apply<DOT>;
            LAUNCH(PC_IFACE, app, grid_for(ctx, P.h.nBRows), P.h.nBRows, P.bRow, P.bStart, ctx->ifaceProd, y, R2);
        } else {
            auto fix = k_iface_fix<0, DOT>;
            LAUNCH(PC_IFACE, fix, grid_for(ctx, P.h.nBRows), P.h.nBRows, P.bRow, P.bStart, P.bSlot,
                   ctx->bou, ctx->recvbuf, ctx->haloDev, x, y, R2);
        }
        if (INIT) {
            auto fix1 = k_iface_fix<1, false>;
            LAUNCH(PC_IFACE, fix1, grid_for(ctx, P.h.nBRows), P.h.nBRows, P.bRow, P.bStart,
                   P.bSlot, ctx->bou, ctx->recvbuf, ctx->haloDev, x, sA, R2);
        }
    }
    if (DOT) RET(reduce_post(ctx, step));
    return B200_OK;
}

int alloc_vectors(b200_ctx* ctx) {
    // rounded up to whole 256-row chunks (+1 chunk): the bulk-copy Amul stages whole chunks
    const size_t n = (((size_t)ctx->N + kChunkRows - 1) / kChunkRows + 1) * kChunkRows;
    RET(dev_alloc(ctx, &ctx->diag, n)); RET(dev_alloc(ctx, &ctx->src, n));
    RET(dev_alloc(ctx, &ctx->psi, n));  RET(dev_alloc(ctx, &ctx->r, n));
    RET(dev_alloc(ctx, &ctx->p, n));    RET(dev_alloc(ctx, &ctx->w, n));
    RET(dev_alloc(ctx, &ctx->rD, n));
    const size_t ns = (size_t)ctx->nSlots;
    RET(dev_alloc(ctx, &ctx->bou, ns)); RET(dev_alloc(ctx, &ctx->sendbuf, ns));
    RET(dev_alloc(ctx, &ctx->recvbuf, ns));
    RET(dev_alloc(ctx, &ctx->ifaceProd, ns));
    return B200_OK;
}

int ensure_staging(b200_ctx* ctx, bool faceTemps) {
    if (!ctx->in_diag) {
        RET(dev_alloc(ctx, &ctx->in_diag, (size_t)ctx->N)); RET(dev_alloc(ctx, &ctx->in_src, (size_t)ctx->N));
        RET(dev_alloc(ctx, &ctx->in_psi, (size_t)ctx->N));  RET(dev_alloc(ctx, &ctx->in_upper, (size_t)ctx->F));
        RET(dev_alloc(ctx, &ctx->in_bou, (size_t)ctx->nSlots));
    }
    if (faceTemps && !ctx->in_f1) {
        RET(dev_alloc(ctx, &ctx->in_f1, (size_t)ctx->F)); RET(dev_alloc(ctx, &ctx->in_f2, (size_t)ctx->F));
        RET(dev_alloc(ctx, &ctx->in_f3, (size_t)ctx->F));
    }
    return B200_OK;
}

// gather the matrix/vectors of one solve into the plan's internal layout
int load_system(b200_ctx* ctx, DevPlan& P, const double* dn_diag, const double* dn_upper,
                const double* dn_src, const double* dn_psi) {
    const int N = ctx->N;
    // the full-row ELL values of a single-read (SR) plan are only read by the cluster kernels of small systems
    const bool ellUnused = P.sr && !(ctx->nranks == 1 && N <= ctx->smallN);
    if (!ellUnused)
        LAUNCH(PC_FILL, k_fill_values, grid_for(ctx, P.h.nEntries, 16), P.h.nEntries, P.faceOf, dn_upper, P.val);
    if (P.sr)
        LAUNCH(PC_FILL, k_fill_values, grid_for(ctx, P.srNOwn, 16), P.srNOwn, P.srOwnFace, dn_upper, P.srOwnVal);
    if (P.sym)
        LAUNCH(PC_FILL, k_fill_values, grid_for(ctx, P.symNU, 16), P.symNU, P.sUFace, dn_upper, P.sUVal);
    LAUNCH(PC_GATHER, k_gather, grid_for(ctx, N), N, P.perm, dn_diag, ctx->diag);
    if (dn_src) LAUNCH(PC_GATHER, k_gather, grid_for(ctx, N), N, P.perm, dn_src, ctx->src);
    if (dn_psi) LAUNCH(PC_GATHER, k_gather, grid_for(ctx, N), N, P.perm, dn_psi, ctx->psi);
    return B200_OK;
}

// launch geometry of a sweep over the rows of colour c (kernels.cuh ColourRows)
ColourRows colour_rows(const b200_ctx* ctx, const DevPlan& P, int c, int* grid) {
    ColourRows cr{P.segStart, P.h.nColours, c, P.h.nTiles, 1};
    const int rows = P.h.colourStart[c + 1] - P.h.colourStart[c];
    if (P.h.nTiles == 1) {
        *grid = grid_for(ctx, rows, ctx->sweepPerSM);
        cr.bps = *grid;
    } else {
        // ~4 rows per thread inside a segment; the grid strides over (tile, block-in-segment) items
        const int segRows = std::max(1, rows / P.h.nTiles);
        cr.bps = std::max(1, std::min(8, (segRows + 4 * kBlock - 1) / (4 * kBlock)));
        const int64_t items = (int64_t)P.h.nTiles * cr.bps;
        *grid = (int)std::min<int64_t>(items, std::min(kMaxGrid, ctx->numSMs * 8));
    }
    return cr;
}

// precondition rA -> wA and form wArA = (wA, rA) [STEP_WARA]; used once before the loop for every
// mode and, for the DIC-class modes, after every iteration (none/diagonal fuse it into k_r).
int enqueue_precondition(b200_ctx* ctx, DevPlan& P, int precond, bool firstColourDone = false) {
    const int N = ctx->N;
    const int gv = grid_for(ctx, (N + 1) / 2);
    if (precond == B200_PRECOND_NONE) {
        Reduce R = mkR(ctx, STEP_WARA);
        auto k = k_precond_dot<false>;
        LAUNCH(PC_PRECOND_DOT, k, gv, N, (const double*)nullptr, ctx->r, (double*)nullptr, R);
    } else if (precond == B200_PRECOND_DIAGONAL) {
        Reduce R = mkR(ctx, STEP_WARA);
        auto k = k_precond_dot<true>;
        LAUNCH(PC_PRECOND_DOT, k, gv, N, ctx->rD, ctx->r, ctx->w, R);
    } else {
        const int C = P.h.nColours;
        for (int k = firstColourDone ? 1 : 0; k < C; ++k) {   // (first colour: fused into k_r<2>)
            int g;
            const ColourRows cr = colour_rows(ctx, P, k, &g);
            const bool last = (k == C - 1);
            Reduce R = mkR(ctx, (last && C == 1) ? STEP_WARA : STEP_NONE);
            const EllCols E{P.col, P.col16, P.colBase};
#define B200_FWD(DOT_, C16_)                                                                             \
    do {                                                                                                 \
        auto kf = k_dic_fwd<DOT_, C16_>;                                                                 \
        LAUNCH(PC_DIC_FWD, kf, g, cr, P.sliceBase, P.rowLen, E, P.val, ctx->rD, ctx->r, ctx->w, R);       \
    } while (0)
            if (last && P.c16) B200_FWD(true, true);
            else if (last) B200_FWD(true, false);
            else if (P.c16) B200_FWD(false, true);
            else B200_FWD(false, false);
#undef B200_FWD
        }
        for (int k = C - 2; k >= 0; --k) {
            int g;
            const ColourRows cr = colour_rows(ctx, P, k, &g);
            Reduce R = mkR(ctx, k == 0 ? STEP_WARA : STEP_NONE);
            const EllCols E{P.col, P.col16, P.colBase};
            if (P.c16) {
                auto kb = k_dic_bwd<true>;
                LAUNCH(PC_DIC_BWD, kb, g, cr, P.sliceBase, P.rowLen, E, P.val, ctx->rD, ctx->r, ctx->w, R);
            } else {
                auto kb = k_dic_bwd<false>;
                LAUNCH(PC_DIC_BWD, kb, g, cr, P.sliceBase, P.rowLen, E, P.val, ctx->rD, ctx->r, ctx->w, R);
            }
        }
    }
    RET(reduce_post(ctx, STEP_WARA));
    return B200_OK;
}

// One loop body of PCG::solve, regrouped (see kernels.cuh "fused PCG vector kernels"):
//   k_p -> Amul+wApA -> k_r (+ next wArA) [-> DIC-class sweeps + next wArA]
int enqueue_iteration(b200_ctx* ctx, DevPlan& P, int precond) {
    const int N = ctx->N;
    Scalars* S = ctx->S;
    const int gv = grid_for(ctx, (N + 1) / 2);
    // N > 1, peer-memory halos, an Amul that corrects its own interface rows: k_p also PACKS -- its CTAs store the
    // patch-face values of the rows they have just written into the neighbours (kernels.cuh PackTail), so the loop
    // body is the same three launches as on one GPU
    PackTail T = {nullptr, nullptr, nullptr, ctx->haloDev};
    const bool packed = amul_fuses_iface(ctx, P);
    if (packed) {
        RET(ensure_pack_lists(ctx, P, gv));
        T = PackTail{P.ctaSStart, P.ctaS, P.slotRow, ctx->haloDev};
    }
    if (precond == B200_PRECOND_NONE) {
        auto k = k_p<0>;
        LAUNCH(PC_KP, k, gv, N, ctx->psi, ctx->p, ctx->r, ctx->rD, ctx->w, S, T);
    } else if (precond == B200_PRECOND_DIAGONAL) {
        auto k = k_p<1>;
        LAUNCH(PC_KP, k, gv, N, ctx->psi, ctx->p, ctx->r, ctx->rD, ctx->w, S, T);
    } else {
        auto k = k_p<2>;
        LAUNCH(PC_KP, k, gv, N, ctx->psi, ctx->p, ctx->r, ctx->rD, ctx->w, S, T);
    }
    RET((spmv_full<false, true>(ctx, P, ctx->p, ctx->w, nullptr, STEP_WAPA, packed)));
    if (precond == B200_PRECOND_NONE) {
        Reduce R = mkR(ctx, STEP_RES_WARA);
        auto k = k_r<0>;
        LAUNCH(PC_KR, k, gv, N, ctx->r, ctx->w, ctx->rD, FirstColour{nullptr, 1, 31}, R);
        RET(reduce_post(ctx, STEP_RES_WARA));
    } else if (precond == B200_PRECOND_DIAGONAL) {
        Reduce R = mkR(ctx, STEP_RES_WARA);
        auto k = k_r<1>;
        LAUNCH(PC_KR, k, gv, N, ctx->r, ctx->w, ctx->rD, FirstColour{nullptr, 1, 31}, R);
        RET(reduce_post(ctx, STEP_RES_WARA));
    } else {
        Reduce R = mkR(ctx, STEP_RES);
        // with >= 2 colours the first colour's forward sweep (wA = rD*rA) rides along in k_r
        const bool fuse = P.h.nColours >= 2 && !ctx->noFuseFirst;
        auto k = k_r<2>;
        LAUNCH(PC_KR, k, gv, N, ctx->r, ctx->w, ctx->rD,
               FirstColour{fuse ? P.segStart : nullptr, P.h.nColours, P.tileShift}, R);
        RET(reduce_post(ctx, STEP_RES));
        RET(enqueue_precondition(ctx, P, precond, fuse));
    }
    return B200_OK;
}

// ---- Eisenstat form of the DIC-class loop (kernels.cuh "Eisenstat form") ----------------------
// Buffers: rh = ctx->r, ph = ctx->p, y = ctx->w, t = ctx->t, sv = ctx->dT (D~ during set-up, then
// s = 1/sqrt|D~|), eb = ctx->eD, xa = ctx->rD (the reciprocal diagonal itself is not used by this form).
int ensure_eis_buffers(b200_ctx* ctx, DevPlan& P) {
    if (!ctx->t) {
        const size_t n = (((size_t)ctx->N + kChunkRows - 1) / kChunkRows + 1) * kChunkRows;
        RET(dev_alloc(ctx, &ctx->t, n)); RET(dev_alloc(ctx, &ctx->dT, n)); RET(dev_alloc(ctx, &ctx->eD, n));
        CU(cudaMemsetAsync(ctx->t, 0, n * sizeof(double), ctx->sc));
    }
    if (ctx->nranks > 1 && P.nSlots > 0 && !P.rowB) {
        std::vector<int32_t> rb((size_t)ctx->N, -1);
        for (int b = 0; b < P.h.nBRows; ++b) rb[(size_t)P.h.bRow[b]] = b;
        RET(upload(ctx, &P.rowB, rb));
        RET(dev_alloc(ctx, &P.hb, (size_t)P.h.nBRows));
        P.nB0 = 0;
        while (P.nB0 < P.h.nBRows && P.h.bRow[(size_t)P.nB0] < P.h.colourStart[1]) ++P.nB0;
        CU(cudaStreamSynchronize(ctx->sc));   // rb goes out of scope
    }
    return B200_OK;
}

// entries gathered per batch by the sweeps (kernels.cuh eis_row_sub): 0 = plain loop (B200PCG_EIS_BATCH=0)
int eis_batch(const b200_ctx* ctx, const DevPlan& P) {
    if (ctx->eisBatch == 0) return 0;
    return P.maxRowLen <= 6 ? 6 : 8;
}

// exchange of t across the processor patches on the comm stream (after everything enqueued on the compute
// stream so far); the compute stream does NOT wait here: eis_halo_wait does
int eis_halo_start(b200_ctx* ctx, DevPlan& P) {
    // (on the comm stream for either transport: the pack kernel's system-scope fence waits for the NVLink
    // acknowledgements, ~8 us that are hidden there and were exposed when the pack ran on the main stream:
    // 571 -> 589 us per iteration at 2 GPUs)
    CU(cudaEventRecord(ctx->evPack, ctx->sc));
    CU(cudaStreamWaitEvent(ctx->sm, ctx->evPack, 0));
    RET(halo_exchange(ctx, P, ctx->t, ctx->sm));
    CU(cudaEventRecord(ctx->evRecv, ctx->sm));
    return B200_OK;
}
// ... and the halo term hb = B- t once it has arrived
int eis_halo_wait(b200_ctx* ctx, DevPlan& P, bool firstColourRows = false) {
    CU(cudaStreamWaitEvent(ctx->sc, ctx->evRecv, 0));
    Reduce R = mkR(ctx, STEP_NONE);
    if (firstColourRows) {
        auto kh = k_eis_halo<true>;
        LAUNCH(PC_IFACE, kh, grid_for(ctx, P.h.nBRows), P.h.nBRows, P.bStart, P.bSlot, ctx->bou,
               ctx->recvbuf, ctx->haloDev, P.hb, ctx->S, P.nB0, P.bRow, ctx->p, ctx->t, ctx->w, R);
    } else {
        auto kh = k_eis_halo<false>;
        LAUNCH(PC_IFACE, kh, grid_for(ctx, P.h.nBRows), P.h.nBRows, P.bStart, P.bSlot, ctx->bou,
               ctx->recvbuf, ctx->haloDev, P.hb, ctx->S, 0, P.bRow, ctx->p, ctx->t, ctx->w, R);
    }
    return B200_OK;
}

// fuse0: 0 = every colour gets its own forward launch; 1 = single rank: the first colour's forward sweep
// rides in its backward sweep; 2 = nranks > 1 with the halo exchange overlapped (kernels.cuh k_eis_bwd)
template <bool C16, int B, int CTB, int CTF>
int launch_eis_sweeps(b200_ctx* ctx, DevPlan& P, bool halo, int fuse0) {
    const int C = P.h.nColours;
    Scalars* S = ctx->S;
    const EllCols E{P.col, P.col16, P.colBase};
    // batched sweeps: one resident wave (the CTAs per SM the kernels are compiled for); plain loops: 8 per SM
    const int perSMB = ctx->sweepPerSMSet ? ctx->sweepPerSM : (B == 0 ? 8 : CTB);
    const int perSMF = ctx->sweepPerSMSet ? ctx->sweepPerSM : (B == 0 ? 8 : CTF);
    for (int k = C - 2; k >= 0; --k) {
        const int r0 = P.h.colourStart[k], r1 = P.h.colourStart[k + 1];
        const int g = grid_for(ctx, r1 - r0, perSMB);
        Reduce R = mkR(ctx, STEP_NONE);
        if (k == 0 && fuse0 == 2) {
            // the first colour's interface rows first, so that the exchange of t overlaps the bulk of the colour
            if (P.nB0 > 0) {
                auto kr = k_eis_bwd_rows<C16>;
                LAUNCH(PC_EIS_ROWS, kr, grid_for(ctx, P.nB0), P.nB0, P.bRow, P.sliceBase, P.rowLen, E, P.val, ctx->p,
                       ctx->t, S);
            }
            RET(eis_halo_start(ctx, P));
            auto kb = k_eis_bwd<2, C16, B, CTB>;
            LAUNCH(PC_EIS_BWD, kb, g, r0, r1, P.sliceBase, P.rowLen, E, P.val, ctx->p, ctx->t, ctx->w, P.rowB, R);
        } else if (k == 0 && fuse0 == 1) {
            auto kb = k_eis_bwd<1, C16, B, CTB>;
            LAUNCH(PC_EIS_BWD, kb, g, r0, r1, P.sliceBase, P.rowLen, E, P.val, ctx->p, ctx->t, ctx->w, P.rowB, R);
        } else {
            auto kb = k_eis_bwd<0, C16, B, CTB>;
            LAUNCH(PC_EIS_BWD, kb, g, r0, r1, P.sliceBase, P.rowLen, E, P.val, ctx->p, ctx->t, ctx->w, P.rowB, R);
        }
    }
    if (halo && fuse0 == 2) {
        RET(eis_halo_wait(ctx, P, P.nB0 > 0));     // + the forward part of the first colour's interface rows
    } else if (halo) {
        // t is complete on every rank: pack + exchange on the comm stream, then the halo term B- t
        RET(eis_halo_start(ctx, P));
        RET(eis_halo_wait(ctx, P));
    }
    for (int k = fuse0 ? 1 : 0; k < C; ++k) {
        const int r0 = P.h.colourStart[k], r1 = P.h.colourStart[k + 1];
        const int g = grid_for(ctx, r1 - r0, perSMF);
        const bool last = (k == C - 1);
        Reduce R = mkR(ctx, last ? STEP_WAPA : STEP_NONE);
#define B200_EFWD(L_, H_)                                                                                 \
    do {                                                                                                  \
        auto kf = k_eis_fwd<L_, H_, C16, B, CTF>;                                                             \
        LAUNCH(PC_EIS_FWD, kf, g, r0, r1, P.sliceBase, P.rowLen, E, P.val, ctx->p, ctx->eD, ctx->t, ctx->w, \
               P.rowB, P.hb, R);                                                                          \
    } while (0)
        if (last && halo) B200_EFWD(true, true);
        else if (last) B200_EFWD(true, false);
        else if (halo) B200_EFWD(false, true);
        else B200_EFWD(false, false);
#undef B200_EFWD
    }
    return B200_OK;
}

// per solve: D~ (DIC recurrence, one launch per colour), its sign, the symmetric scaling (s, D- - 2, the
// coefficient copy, the interface coefficients), r^ = (I+L-)^-1 sigma S r, rho_0
int eis_setup(b200_ctx* ctx, DevPlan& P) {
    if (P.h.nTiles > 1)
        return fail(ctx, B200_EUNSUPPORTED, "dicMode eisenstat needs the colour-major plan (unset B200PCG_TILE)");
    RET(ensure_eis_buffers(ctx, P));
    const int N = ctx->N;
    Scalars* S = ctx->S;
    const int gv = grid_for(ctx, (N + 1) / 2);
    const int gn = grid_for(ctx, N);
    const EllCols E{P.col, P.col16, P.colBase};
    for (int k = 0; k < P.h.nColours; ++k) {
        int g;
        const ColourRows cr = colour_rows(ctx, P, k, &g);
        if (P.c16) {
            auto kd = k_dic_calc_rd<true>;
            LAUNCH(PC_DIC_RD, kd, g, cr, P.sliceBase, P.rowLen, E, P.val, ctx->diag, ctx->dT);
        } else {
            auto kd = k_dic_calc_rd<false>;
            LAUNCH(PC_DIC_RD, kd, g, cr, P.sliceBase, P.rowLen, E, P.val, ctx->diag, ctx->dT);
        }
    }
    {
        Reduce R = mkR(ctx, STEP_EIS_SIGN);
        LAUNCH(PC_EIS_SETUP, k_eis_sign, gn, N, ctx->dT, R);
        RET(reduce_post(ctx, STEP_EIS_SIGN));
    }
    LAUNCH(PC_EIS_SETUP, k_eis_setup, gn, N, ctx->diag, ctx->dT, ctx->eD, ctx->r, ctx->rD, S);
    if (ctx->nranks > 1 && P.nSlots > 0) {
        // the neighbours' s on the patch faces -> interface coefficients of B- (once per solve: same stream)
        if (ctx->p2pHalo) {
            RET(halo_exchange(ctx, P, ctx->dT, ctx->sc));
        } else {
            k_eis_pack_s<<<grid_for(ctx, P.nSlots), kBlock, 0, ctx->sc>>>(P.nSlots, P.slotRow, ctx->dT, ctx->sendbuf);
            ctx->launches++;
            NC(g_nccl.GroupStart());
            for (int k = 0; k < P.h.nIfaces; ++k) {
                const int off = P.h.patchStart[k], n = P.h.patchStart[k + 1] - off;
                if (n == 0) continue;
                NC(g_nccl.Send(ctx->sendbuf + off, n, ncclDouble, P.h.nbrRank[k], ctx->comm, ctx->sc));
                NC(g_nccl.Recv(ctx->recvbuf + off, n, ncclDouble, P.h.nbrRank[k], ctx->comm, ctx->sc));
            }
            NC(g_nccl.GroupEnd());
        }
        LAUNCH(PC_EIS_SETUP, k_eis_scale_bou, grid_for(ctx, P.nSlots), P.nSlots, P.slotRow, ctx->dT, ctx->recvbuf,
               ctx->haloDev, ctx->bou, S);
    }
    if (P.c16) {
        auto ks = k_eis_scale_vals<true>;
        LAUNCH(PC_EIS_SETUP, ks, gn, N, P.sliceBase, P.rowLen, E, P.val, ctx->dT, S);
    } else {
        auto ks = k_eis_scale_vals<false>;
        LAUNCH(PC_EIS_SETUP, ks, gn, N, P.sliceBase, P.rowLen, E, P.val, ctx->dT, S);
    }
    for (int k = 0; k < P.h.nColours; ++k) {
        const int r0 = P.h.colourStart[k], r1 = P.h.colourStart[k + 1];
        const int g = grid_for(ctx, r1 - r0, ctx->sweepPerSM);
        if (P.c16) {
            auto kf = k_eis_init_fwd<true>;
            LAUNCH(PC_EIS_SETUP, kf, g, r0, r1, P.sliceBase, P.rowLen, E, P.val, ctx->r, S);
        } else {
            auto kf = k_eis_init_fwd<false>;
            LAUNCH(PC_EIS_SETUP, kf, g, r0, r1, P.sliceBase, P.rowLen, E, P.val, ctx->r, S);
        }
    }
    Reduce R = mkR(ctx, STEP_EIS_RHO0);
    LAUNCH(PC_EIS_SETUP, k_eis_rho0, gv, N, ctx->r, R);
    RET(reduce_post(ctx, STEP_EIS_RHO0));
    return B200_OK;
}

// One loop body: k_eis_p -> backward sweeps -> [halo exchange of t] -> forward sweeps (+ (p^, w^)) ->
// k_eis_r (+ rho) -> k_eis_res (true residual, device-gated)
int enqueue_eis_iteration(b200_ctx* ctx, DevPlan& P) {
    const int N = ctx->N;
    Scalars* S = ctx->S;
    const int gv = grid_for(ctx, (N + 1) / 2);
    const int C = P.h.nColours;
    const int lastStart = P.h.colourStart[C - 1];
    const bool halo = (ctx->nranks > 1 && P.nSlots > 0);
    // first colour's forward sweep inside its backward sweep: always on one rank; with processor patches only
    // in the overlapped form (B200PCG_EIS_OVERLAP=1; not yet the default: validated on the GPU in round 2)
    const int fuse0 = C < 2 ? 0 : (!halo ? 1 : (ctx->eisOverlap ? 2 : 0));
    LAUNCH(PC_EIS_P, k_eis_p, gv, N, lastStart, ctx->rD, ctx->p, ctx->t, ctx->r, S);
    const int B = eis_batch(ctx, P);
    // CTB / CTF = resident CTAs per SM the backward / forward sweep kernels are compiled for (register cap
    // 65536 / (256 CT)).  Measured on the 16 M hex box (profiles/r01_v10_eisenstat_sweep_builds_ab.log): the
    // backward sweep is fastest in the 64-register build (141 vs 170 us), the forward sweep -- two more
    // per-row streams -- in the 80-register build, where all 12 column loads of a batch get their own
    // registers (141 vs 183 us).  B200PCG_EIS_CTAS=3|4 forces one build for both.
    const int ctb = ctx->eisCtas ? ctx->eisCtas : 4, ctf = ctx->eisCtas ? ctx->eisCtas : 3;
#define B200_SWEEPS(C16_)                                                                                     \
    do {                                                                                                      \
        if (B == 0) RET((launch_eis_sweeps<C16_, 0, 1, 1>(ctx, P, halo, fuse0)));                              \
        else if (B == 8) RET((launch_eis_sweeps<C16_, 8, 3, 3>(ctx, P, halo, fuse0)));                         \
        else if (ctb == 4 && ctf == 3) RET((launch_eis_sweeps<C16_, 6, 4, 3>(ctx, P, halo, fuse0)));           \
        else if (ctb == 3) RET((launch_eis_sweeps<C16_, 6, 3, 3>(ctx, P, halo, fuse0)));                       \
        else RET((launch_eis_sweeps<C16_, 6, 4, 4>(ctx, P, halo, fuse0)));                                     \
    } while (0)
    if (P.c16) B200_SWEEPS(true);
    else B200_SWEEPS(false);
#undef B200_SWEEPS
    RET(reduce_post(ctx, STEP_WAPA));
    {
        Reduce R = mkR(ctx, STEP_EIS_RHO);
        LAUNCH(PC_EIS_R, k_eis_r, gv, N, lastStart, ctx->r, ctx->w, ctx->t, R);
        RET(reduce_post(ctx, STEP_EIS_RHO));
    }
    {
        Reduce R = mkR(ctx, STEP_EIS_RES);
        const EllCols E{P.col, P.col16, P.colBase};
        // batched like the sweeps (one resident wave), or the plain loop at 8 CTAs per SM
        const int gr = grid_for(ctx, N, B == 0 ? 8 : eis_sweep_ctas(B));
#define B200_ERES(C16_, B_)                                                                                \
    do {                                                                                                   \
        auto kr = k_eis_res<C16_, B_>;                                                                     \
        LAUNCH(PC_EIS_RES, kr, gr, N, P.sliceBase, P.rowLen, E, P.val, ctx->dT, ctx->r, R);                \
    } while (0)
        if (P.c16) {
            if (B == 0) B200_ERES(true, 0);
            else if (B == 6) B200_ERES(true, 6);
            else B200_ERES(true, 8);
        } else {
            if (B == 0) B200_ERES(false, 0);
            else if (B == 6) B200_ERES(false, 6);
            else B200_ERES(false, 8);
        }
#undef B200_ERES
        RET(reduce_post(ctx, STEP_EIS_RES));
    }
    return B200_OK;
}

int copy_bou(b200_ctx* ctx, DevPlan& P, const double* const* bouPtrs, cudaMemcpyKind kind, double* dst) {
    for (int k = 0; k < P.h.nIfaces; ++k) {
        const int off = P.h.patchStart[k], n = P.h.patchStart[k + 1] - off;
        if (n == 0) continue;
        if (!bouPtrs || !bouPtrs[k]) return fail(ctx, B200_EINVAL, "null interfaceBouCoeffs entry");
        CU(cudaMemcpyAsync(dst + off, bouPtrs[k], (size_t)n * sizeof(double), kind, ctx->sc));
    }
    return B200_OK;
}

constexpr int kGraphIters = 4;   // loop bodies per iteration graph (= the first batch size of the loop)

Ordering ordering_for(int precond) {
    if (precond == B200_PRECOND_DIC_MC || precond == B200_PRECOND_DIC_MC_EIS || precond == B200_PRECOND_DIC_MC_LOOP)
        return Ordering::MultiColour;
    if (precond == B200_PRECOND_DIC_EXACT) return Ordering::Levels;
    return Ordering::Natural;
}

// fill b200_perf from the status block read back after the loop
int finish_solve(b200_ctx* ctx, DevPlan& P, b200_perf* perf) {
    const Scalars& h = *ctx->hS;
    if (perf) {
        std::memset(perf, 0, sizeof(*perf));
        perf->initialResidual = h.initRes;
        perf->finalResidual = h.finalRes;
        perf->normFactor = h.normFactor;
        perf->nIterations = h.nIter;
        perf->converged = h.converged;
        perf->singular = h.singular;
        perf->nColours = P.h.nColours;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        perf->setupMs = ms;
        cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
        perf->solveMs = ms;
    }
    if (h.nonfinite == 3)
        return fail(ctx, B200_EUNSUPPORTED, "DIC pivots are zero or of mixed sign: dicMode eisenstat needs a definite matrix "
                                            "(use dicMode multicolour)");
    if (h.nonfinite == 2) return fail(ctx, B200_ENCCL, "peer-memory all-reduce timed out (a rank never arrived)");
    if (h.nonfinite) return fail(ctx, B200_ENONFINITE, "non-finite residual in PCG");
    return B200_OK;
}

// The solve proper; all pointers are device pointers in natural order; bou already in ctx->bou.
int solve_core(b200_ctx* ctx, const double* dn_diag, const double* dn_upper, const double* dn_src,
               double* dn_psi, const b200_controls* ctl, b200_perf* perf, bool forceLoop = false) {
    if (ctl->precond < 0 || ctl->precond > 5) return fail(ctx, B200_EINVAL, "bad preconditioner code");
    // Eisenstat form of the DIC-class loop: same preconditioner as B200_PRECOND_DIC_MC.  Systems small enough
    // for the single-launch cluster kernels are latency-bound, not bandwidth-bound: they take that path
    // B200PCG_DIC=eisenstat: `preconditioner DIC` without a dicMode takes the Eisenstat form as well (the switch
    // that becomes the default once the form has run on 4 and 8 GPUs); tiled plans keep the three-kernel loop
    // code 2 (`preconditioner DIC` without a dicMode): Eisenstat's form unless B200PCG_DIC=multicolour, a tiled
    // plan, or (retry below) a matrix whose DIC pivots are not of one sign; code 5: always the three-kernel loop
    bool eis = (ctl->precond == B200_PRECOND_DIC_MC_EIS);
    const bool autoForm = (ctl->precond == B200_PRECOND_DIC_MC);
    if (autoForm && ctx->dicDefaultEis && ctx->tileRows == 0 && !forceLoop) eis = true;
    const int32_t precond = (ctl->precond == B200_PRECOND_DIC_MC_LOOP || ctl->precond == B200_PRECOND_DIC_MC_EIS)
                                ? (int32_t)B200_PRECOND_DIC_MC : ctl->precond;   // kernel-side code: 0..3
    const int32_t smallPrecond = precond;
    if (ctl->reserved != 0) return fail(ctx, B200_EINVAL, "b200_controls.reserved must be 0");
    DevPlan* Pp = nullptr;
    RET(ensure_plan(ctx, ordering_for(ctl->precond), &Pp));
    DevPlan& P = *Pp;
    const int N = ctx->N;
    const int gv = grid_for(ctx, (N + 1) / 2);
    Scalars* S = ctx->S;

    CU(cudaEventRecord(ctx->ev[0], ctx->sc));
    RET(reset_scalars(ctx, ctl));
    RET(load_system(ctx, P, dn_diag, dn_upper, dn_src, dn_psi));
    ctx->usedSmall = false;
    if (ctx->nranks == 1 && N > 0 && N <= ctx->smallN) {
        // small system: the whole solve in one launch of one thread-block cluster
        ctx->usedSmall = true;
        ctx->usedFast = false;
        // on-chip variant: matrix + gathered vectors in the cluster's shared memory (<= 2 rows per thread)
        int fastCtas = 0, fastRpt = 0;
        if (!ctx->disableFast) {
            for (int nc : {1, 2, 4, 8, 16}) {
                if (nc > ctx->fastMaxCtas) break;
                const int rpt = (N + nc * kSmallBlock - 1) / (nc * kSmallBlock);
                if (rpt <= 2 && fast_smem_bytes(rpt, P.maxRowLen) <= (size_t)225 * 1024) {
                    // prefer one row per thread when a larger cluster offers it
                    if (fastCtas == 0 || (fastRpt == 2 && rpt == 1)) { fastCtas = nc; fastRpt = rpt; }
                    if (rpt == 1) break;
                }
            }
        }
        SmallArgs a{N, smallPrecond, P.h.nColours, P.colourStart, P.sliceBase, P.rowLen, P.col, P.val,
                    ctx->diag, ctx->src, ctx->psi, ctx->r, ctx->p, ctx->w, ctx->rD, S, ctx->partials};
        FastArgs fa{N, smallPrecond, P.h.nColours, P.maxRowLen, P.colourStart, P.sliceBase, P.rowLen, P.col, P.val,
                    ctx->diag, ctx->src, ctx->psi, S};
        int nCtas = std::min(ctx->smallCtas, (N + kSmallBlock - 1) / kSmallBlock);
        if (nCtas > 8) nCtas = 16;
        else if (nCtas > 4) nCtas = 8;
        else if (nCtas > 2) nCtas = 4;
        size_t dynSmem = 0;
        if (fastCtas) {
            ctx->usedFast = true;
            nCtas = fastCtas;
            dynSmem = fast_smem_bytes(fastRpt, P.maxRowLen);
            if (fastRpt == 1) CU(cudaFuncSetAttribute(k_pcg_small_fast<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynSmem));
            else CU(cudaFuncSetAttribute(k_pcg_small_fast<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynSmem));
        }
        if (nCtas > 8) {
            CU(cudaFuncSetAttribute(k_pcg_small, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            CU(cudaFuncSetAttribute(k_pcg_small_fast<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            CU(cudaFuncSetAttribute(k_pcg_small_fast<2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)nCtas);
        cfg.blockDim = dim3(kSmallBlock);
        cfg.stream = ctx->sc;
        cfg.dynamicSmemBytes = dynSmem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)nCtas;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        CU(cudaEventRecord(ctx->ev[1], ctx->sc));
        prof_begin(ctx, PC_SMALL);
        if (fastCtas && fastRpt == 1) CU(cudaLaunchKernelEx(&cfg, k_pcg_small_fast<1>, fa));
        else if (fastCtas) CU(cudaLaunchKernelEx(&cfg, k_pcg_small_fast<2>, fa));
        else CU(cudaLaunchKernelEx(&cfg, k_pcg_small, a));
        prof_end(ctx, PC_SMALL);
        ctx->launches++;
        CU(cudaEventRecord(ctx->ev[2], ctx->sc));
        CU(cudaMemcpyAsync(ctx->hS, S, sizeof(Scalars), cudaMemcpyDeviceToHost, ctx->sc));
        LAUNCH(PC_GATHER, k_scatter, grid_for(ctx, N), N, P.perm, ctx->psi, dn_psi);
        CU(cudaStreamSynchronize(ctx->sc));
        CU(cudaGetLastError());
        prof_collect(ctx);
        return finish_solve(ctx, P, perf);
    }
    // wA = A psi, sumA -> pA (OpenFOAM also uses pA as the normFactor temporary)
    RET((spmv_full<true, false>(ctx, P, ctx->psi, ctx->w, ctx->p, STEP_NONE)));
    {
        Reduce R = mkR(ctx, STEP_SUMPSI);
        LAUNCH(PC_SUM, k_sum, gv, N, ctx->psi, R);
        RET(reduce_post(ctx, STEP_SUMPSI));
    }
    {
        Reduce R = mkR(ctx, STEP_NORM);
        LAUNCH(PC_NORM, k_norm_resid, gv, N, ctx->w, ctx->p, ctx->src, ctx->r, R);
        RET(reduce_post(ctx, STEP_NORM));
    }
    // preconditioner set-up
    if (precond == B200_PRECOND_DIAGONAL) {
        LAUNCH(PC_RECIP, k_recip, gv, N, ctx->diag, ctx->rD);
    } else if (eis) {
        RET(eis_setup(ctx, P));
    } else if (precond >= B200_PRECOND_DIC_MC) {
        for (int k = 0; k < P.h.nColours; ++k) {
            int g;
            const ColourRows cr = colour_rows(ctx, P, k, &g);
            const EllCols E{P.col, P.col16, P.colBase};
            if (P.c16) {
                auto kd = k_dic_calc_rd<true>;
                LAUNCH(PC_DIC_RD, kd, g, cr, P.sliceBase, P.rowLen, E, P.val, ctx->diag, ctx->rD);
            } else {
                auto kd = k_dic_calc_rd<false>;
                LAUNCH(PC_DIC_RD, kd, g, cr, P.sliceBase, P.rowLen, E, P.val, ctx->diag, ctx->rD);
            }
        }
        LAUNCH(PC_RECIP, k_recip, gv, N, ctx->rD, ctx->rD);
    }
    if (!eis) RET(enqueue_precondition(ctx, P, precond));   // early-exits on the device if converged
    CU(cudaEventRecord(ctx->ev[1], ctx->sc));
    CU(cudaMemcpyAsync(ctx->hS, S, sizeof(Scalars), cudaMemcpyDeviceToHost, ctx->sc));
    CU(cudaStreamSynchronize(ctx->sc));
    CU(cudaGetLastError());
    if (eis && autoForm && ctx->hS->nonfinite == 3) {
        // DIC pivots of mixed sign (the global count decided: every rank takes this branch): the Eisenstat
        // scaling does not exist; `preconditioner DIC` falls back to the three-kernel loop on the same inputs
        // (nothing has been scaled or overwritten yet: the set-up kernels return on S->done)
        prof_collect(ctx);
        return solve_core(ctx, dn_diag, dn_upper, dn_src, dn_psi, ctl, perf, true);
    }

    // PCG loop: batches of iterations, device decides when to stop.  A batch is replayed from a CUDA graph of
    // kGraphIters loop bodies (captured once per plan and loop form, NCCL halo exchange included) unless
    // per-kernel profiling is on; the remainder of a batch is enqueued kernel by kernel.
    int64_t cap = ctx->forceIters > 0 ? ctx->forceIters
                                      : std::max<int64_t>((int64_t)ctl->maxIter + 1, ctl->minIter);
    int64_t enq = 0;
    int chunk = 4;
    const int form = eis ? 3 : std::min(precond, 2);
    auto body = [&]() -> int { return eis ? enqueue_eis_iteration(ctx, P) : enqueue_iteration(ctx, P, precond); };
    while (!ctx->hS->done && enq < cap) {
        int n = (int)std::min<int64_t>(chunk, cap - enq);
        int i = 0;
        // one rank only: with processor patches the loop body forks onto the comm stream, and replaying that
        // fork/join from a graph measured SLOWER than plain launches (2 GPUs: 497 vs 438 us per iteration,
        // profiles/r02_bench_2gpu_graph_ab.md)
        // (B200PCG_GRAPH_MULTI=1: also with processor patches when the loop body stays on ONE stream -- peer-memory
        // halos with the interface fix-up fused into the staged Amul; experiment switch)
        const bool oneStream = ctx->nranks == 1 ||
                               (ctx->graphMulti && ctx->p2pHalo && ctx->fuseIface && form <= 2 &&
                                ((P.sym && P.symTma) || (!P.sym && !P.sr)));
        if (ctx->useGraph && oneStream && !ctx->prof && P.h.nColours <= 8 && !P.iterGraphFailed[form]) {
            if (!P.iterGraph[form] && n >= kGraphIters) {
                const uint64_t l0 = ctx->launches;
                cudaGraph_t g = nullptr;
                int rcB = B200_OK;
                cudaError_t e = cudaStreamBeginCapture(ctx->sc, cudaStreamCaptureModeThreadLocal);
                if (e == cudaSuccess) {
                    for (int k = 0; k < kGraphIters && rcB == B200_OK; ++k) rcB = body();
                    e = cudaStreamEndCapture(ctx->sc, &g);
                }
                if (e == cudaSuccess && rcB == B200_OK) e = cudaGraphInstantiate(&P.iterGraph[form], g, 0);
                if (g) cudaGraphDestroy(g);
                P.iterGraphLaunches[form] = ctx->launches - l0;
                ctx->launches = l0;                          // nothing ran during the capture
                if (e != cudaSuccess || rcB != B200_OK) {    // not fatal: this plan keeps the plain enqueue
                    cudaGetLastError();
                    P.iterGraph[form] = nullptr;
                    P.iterGraphFailed[form] = true;
                }
            }
            if (P.iterGraph[form]) {
                for (; i + kGraphIters <= n; i += kGraphIters) {
                    CU(cudaGraphLaunch(P.iterGraph[form], ctx->sc));
                    ctx->launches += P.iterGraphLaunches[form];
                    ctx->graphLaunches++;
                }
            }
        }
        for (; i < n; ++i) {
            ctx->profIter = (int)(enq + i + 1);
            RET(body());
        }
        ctx->profIter = 0;
        enq += n;
        CU(cudaMemcpyAsync(ctx->hS, S, sizeof(Scalars), cudaMemcpyDeviceToHost, ctx->sc));
        CU(cudaStreamSynchronize(ctx->sc));
        CU(cudaGetLastError());
        if (chunk < 64) chunk *= 2;
    }
    if (eis) LAUNCH(PC_PSI_FINAL, k_eis_final, grid_for(ctx, N), N, ctx->psi, ctx->rD, ctx->t, ctx->dT, S);
    else LAUNCH(PC_PSI_FINAL, k_psi_final, gv, N, ctx->psi, ctx->p, S);   // last deferred psi += alpha*pA
    CU(cudaEventRecord(ctx->ev[2], ctx->sc));
    LAUNCH(PC_GATHER, k_scatter, grid_for(ctx, N), N, P.perm, ctx->psi, dn_psi);
    CU(cudaStreamSynchronize(ctx->sc));
    CU(cudaGetLastError());
    prof_collect(ctx, ctx->hS->nIter);

    return finish_solve(ctx, P, perf);
}

// ---- smoothSolver (kernels.cuh "smoothSolver") ---------------------------------------------
// matrix of one smoothSolver / asymmetric Amul call -> the plan's full-row ELL (both triangles), vectors -> plan order
int load_system_asym(b200_ctx* ctx, DevPlan& P, const double* dn_diag, const double* dn_upper, const double* dn_lower,
                     const double* dn_src, const double* dn_psi, bool transposed = false) {
    const int N = ctx->N;
    if (transposed || (dn_lower && dn_lower != dn_upper))
        LAUNCH(PC_FILL, k_fill_values_asym, grid_for(ctx, N), N, P.sliceBase, P.rowLen, P.faceOf, P.perm, ctx->d_l,
               dn_upper, dn_lower ? dn_lower : dn_upper, P.val, transposed ? P.valT : (double*)nullptr);
    else
        LAUNCH(PC_FILL, k_fill_values, grid_for(ctx, P.h.nEntries, 16), P.h.nEntries, P.faceOf, dn_upper, P.val);
    LAUNCH(PC_GATHER, k_gather, grid_for(ctx, N), N, P.perm, dn_diag, ctx->diag);
    if (dn_src) LAUNCH(PC_GATHER, k_gather, grid_for(ctx, N), N, P.perm, dn_src, ctx->src);
    if (dn_psi) LAUNCH(PC_GATHER, k_gather, grid_for(ctx, N), N, P.perm, dn_psi, ctx->psi);
    return B200_OK;
}

// launch geometry of the smoothSolver kernels: like the Eisenstat sweeps, ONE resident wave of the CTAs per SM the
// kernel is compiled for.  Narrow rows (<= 6 faces): the 64-register build (4 CTAs per SM) or the 80-register one
// (3 per SM; B200PCG_GS_CTAS=3|4 forces either -- measured in profiles/r02_smooth_*); wide rows: batches of 8, 80 registers.
int gs_ctas(const b200_ctx* ctx, const DevPlan& P) {
    if (P.maxRowLen > 6) return 3;
    return ctx->gsCtas == 4 ? 4 : 3;
}
int gs_grid(const b200_ctx* ctx, int rows, int ct) {
    return grid_for(ctx, std::max(1, rows), ctx->sweepPerSMSet ? ctx->sweepPerSM : ct);
}

// nranks > 1: interface-row index of every row (shared with the Eisenstat form) + bPrime of the interface rows
int ensure_gs_halo(b200_ctx* ctx, DevPlan& P) {
    if (!(ctx->nranks > 1 && P.nSlots > 0)) return B200_OK;
    if (!P.rowB) {
        std::vector<int32_t> rb((size_t)ctx->N, -1);
        for (int b = 0; b < P.h.nBRows; ++b) rb[(size_t)P.h.bRow[b]] = b;
        RET(upload(ctx, &P.rowB, rb));
        RET(dev_alloc(ctx, &P.hb, (size_t)P.h.nBRows));
        P.nB0 = 0;
        while (P.nB0 < P.h.nBRows && P.h.bRow[(size_t)P.nB0] < P.h.colourStart[1]) ++P.nB0;
        CU(cudaStreamSynchronize(ctx->sc));   // rb goes out of scope
    }
    if (!P.hbv) RET(dev_alloc(ctx, &P.hbv, (size_t)P.h.nBRows));
    return B200_OK;
}

// The smoothSolver's exchanges run on the MAIN stream (producer, consumer and everything between them are serial
// anyway).  Peer-memory halos alternate two halves of the receive buffer, which is safe as long as at most two
// exchanges lie between two cross-rank reductions (kernels.cuh Halo): one sweep + the residual.  More sweeps per
// residual evaluation, or sweeps without residuals (nSweeps < 0), take the stream-ordered ncclSend/ncclRecv.
struct GsHalo {
    bool on = false, nccl = false;
    Halo H = {nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
};

// one group (dependency level / colour) of a sweep.  res 1: also accumulate the rows' share of sum |residual|
// (formed from the new values); res 2 (kernels.cuh k_gs_rows): the rows' residual at the END OF THE PREVIOUS
// iteration from their old values xo, the new ones written to xw, the reduction completed with STEP_GS_RES.
// xg / xw default to psi (in place).
int launch_gs_rows(b200_ctx* ctx, DevPlan& P, int k, int res, const GsHalo& gh, const double* xg = nullptr,
                   const double* xo = nullptr, double* xw = nullptr) {
    const int r0 = P.h.colourStart[(size_t)k], r1 = P.h.colourStart[(size_t)k + 1];
    if (r1 <= r0) return B200_OK;
    if (!xg) xg = ctx->psi;
    if (!xw) xw = ctx->psi;
    const EllCols E{P.col, P.col16, P.colBase};
    const bool wide = P.maxRowLen > 6;
    const int ct = res ? 3 : gs_ctas(ctx, P);        // (the residual term spills at 64 registers)
    const int g = gs_grid(ctx, r1 - r0, ct);
    Reduce R = mkR(ctx, res == 2 ? STEP_GS_RES : STEP_NONE);
#define B200_GS_ROWS(C16_, B_, CT_, RES_, HALO_)                                                           \
    do {                                                                                                   \
        auto kg = k_gs_rows<C16_, B_, CT_, RES_, HALO_>;                                                   \
        LAUNCH(PC_GS_ROWS, kg, g, r0, r1, P.sliceBase, P.rowLen, E, P.val, ctx->diag, ctx->src, P.rowB, P.hbv, \
               xg, xo, xw, R);                                                                             \
    } while (0)
#define B200_GS_ROWS_C(C16_)                                                                               \
    do {                                                                                                   \
        if (gh.on && wide) B200_GS_ROWS(C16_, 8, 3, 0, true);                                              \
        else if (gh.on && ct == 4) B200_GS_ROWS(C16_, 6, 4, 0, true);                                      \
        else if (gh.on) B200_GS_ROWS(C16_, 6, 3, 0, true);                                                 \
        else if (wide && res == 2) B200_GS_ROWS(C16_, 8, 3, 2, false);                                     \
        else if (wide && res) B200_GS_ROWS(C16_, 8, 3, 1, false);                                          \
        else if (wide) B200_GS_ROWS(C16_, 8, 3, 0, false);                                                 \
        else if (res == 2) B200_GS_ROWS(C16_, 6, 3, 2, false);                                             \
        else if (res) B200_GS_ROWS(C16_, 6, 3, 1, false);                                                  \
        else if (ct == 4) B200_GS_ROWS(C16_, 6, 4, 0, false);                                              \
        else B200_GS_ROWS(C16_, 6, 3, 0, false);                                                           \
    } while (0)
    if (P.c16) B200_GS_ROWS_C(true);
    else B200_GS_ROWS_C(false);
#undef B200_GS_ROWS_C
#undef B200_GS_ROWS
    return B200_OK;
}

// nranks > 1, start of a sweep: exchange psi across the processor patches, bPrime of the interface rows
int gs_sweep_halo(b200_ctx* ctx, DevPlan& P, const GsHalo& gh) {
    if (!gh.on) return B200_OK;
    RET(halo_exchange(ctx, P, ctx->psi, ctx->sc, gh.nccl));
    LAUNCH(PC_IFACE, k_gs_bprime, grid_for(ctx, P.h.nBRows), P.h.nBRows, P.bRow, P.bStart, P.bSlot, ctx->bou,
           ctx->recvbuf, gh.H, ctx->src, P.hbv, ctx->S);
    return B200_OK;
}

// sum |residual| over the rows [r0, r1) (an empty range still completes the reduction) + STEP_GS_RES
int launch_gs_resid(b200_ctx* ctx, DevPlan& P, int r0, int r1, const GsHalo& gh) {
    const EllCols E{P.col, P.col16, P.colBase};
    const bool wide = P.maxRowLen > 6;
    const int ct = gs_ctas(ctx, P);
    const int g = gs_grid(ctx, r1 - r0, ct);
    if (gh.on) RET(halo_exchange(ctx, P, ctx->psi, ctx->sc, gh.nccl));
    Reduce R = mkR(ctx, STEP_GS_RES);
#define B200_GS_RESID(C16_, B_, CT_, HALO_)                                                                \
    do {                                                                                                   \
        auto kg = k_gs_resid<C16_, B_, CT_, HALO_>;                                                        \
        LAUNCH(PC_GS_RESID, kg, g, r0, r1, P.sliceBase, P.rowLen, E, P.val, ctx->diag, ctx->src, ctx->psi, \
               P.rowB, P.bStart, P.bSlot, ctx->bou, ctx->recvbuf, gh.H, R);                                \
    } while (0)
#define B200_GS_RESID_C(C16_)                                                                              \
    do {                                                                                                   \
        if (gh.on && wide) B200_GS_RESID(C16_, 8, 3, true);                                                \
        else if (gh.on && ct == 4) B200_GS_RESID(C16_, 6, 4, true);                                        \
        else if (gh.on) B200_GS_RESID(C16_, 6, 3, true);                                                   \
        else if (wide) B200_GS_RESID(C16_, 8, 3, false);                                                   \
        else if (ct == 4) B200_GS_RESID(C16_, 6, 4, false);                                                \
        else B200_GS_RESID(C16_, 6, 3, false);                                                             \
    } while (0)
    if (P.c16) B200_GS_RESID_C(true);
    else B200_GS_RESID_C(false);
#undef B200_GS_RESID_C
#undef B200_GS_RESID
    RET(reduce_post(ctx, STEP_GS_RES));
    return B200_OK;
}

// smoothSolver::solve (OF-dev smoothSolver.C); all pointers are device pointers in natural order
int smooth_core(b200_ctx* ctx, const double* dn_diag, const double* dn_upper, const double* dn_lower,
                const double* dn_src, double* dn_psi, const b200_smooth_controls* ctl, b200_perf* perf) {
    if (ctl->smoother != B200_SMOOTHER_GAUSS_SEIDEL && ctl->smoother != B200_SMOOTHER_SYM_GAUSS_SEIDEL)
        return fail(ctx, B200_EINVAL, "bad smoother code");
    if (ctl->sweepMode != B200_SWEEP_MULTICOLOUR && ctl->sweepMode != B200_SWEEP_EXACT)
        return fail(ctx, B200_EINVAL, "bad sweepMode code");
    if (ctl->reserved != 0) return fail(ctx, B200_EINVAL, "b200_smooth_controls.reserved must be 0");
    if (ctl->nSweeps == 0) return fail(ctx, B200_EINVAL, "nSweeps must not be 0");
    if (ctx->tileRows > 0) return fail(ctx, B200_EUNSUPPORTED, "smoothSolver does not take tiled plans (B200PCG_TILE)");
    DevPlan* Pp = nullptr;
    RET(ensure_plan(ctx, ctl->sweepMode == B200_SWEEP_EXACT ? Ordering::Levels : Ordering::MultiColour, &Pp));
    DevPlan& P = *Pp;
    const int N = ctx->N;
    const int gv = grid_for(ctx, (N + 1) / 2);
    Scalars* S = ctx->S;
    const bool fixed = ctl->nSweeps < 0;             // exactly -nSweeps sweeps, no residual evaluation
    const int nS = fixed ? -ctl->nSweeps : ctl->nSweeps;
    const bool sym = ctl->smoother == B200_SMOOTHER_SYM_GAUSS_SEIDEL;
    const int C = P.h.nColours;
    // processor patches (nranks > 1): explicit, refreshed once per sweep, as upstream
    GsHalo gh;
    gh.on = ctx->nranks > 1 && P.nSlots > 0;
    if (gh.on) {
        RET(ensure_gs_halo(ctx, P));
        gh.nccl = !ctx->p2pHalo || fixed || nS > 1;
        if (!gh.nccl) gh.H = ctx->haloDev;
    }
    const bool multi = ctx->nranks > 1;

    CU(cudaEventRecord(ctx->ev[0], ctx->sc));
    const b200_controls c{ctl->tolerance, ctl->relTol, ctl->maxIter, ctl->minIter, 0, 0};
    ctx->gsSweeps = nS;
    RET(reset_scalars(ctx, &c));
    RET(load_system_asym(ctx, P, dn_diag, dn_upper, dn_lower, dn_src, dn_psi));
    ctx->usedSmall = false;
    if (!fixed) {
        // wA = A psi, sumA -> pA; normFactor; initial residual and the first convergence test (STEP_NORM)
        RET((spmv_full<true, false>(ctx, P, ctx->psi, ctx->w, ctx->p, STEP_NONE)));
        {
            Reduce R = mkR(ctx, STEP_SUMPSI);
            LAUNCH(PC_SUM, k_sum, gv, N, ctx->psi, R);
            RET(reduce_post(ctx, STEP_SUMPSI));
        }
        {
            Reduce R = mkR(ctx, STEP_NORM);
            LAUNCH(PC_NORM, k_norm_resid, gv, N, ctx->w, ctx->p, ctx->src, ctx->r, R);
            RET(reduce_post(ctx, STEP_NORM));
        }
    }
    CU(cudaEventRecord(ctx->ev[1], ctx->sc));
    CU(cudaMemcpyAsync(ctx->hS, S, sizeof(Scalars), cudaMemcpyDeviceToHost, ctx->sc));
    CU(cudaStreamSynchronize(ctx->sc));
    CU(cudaGetLastError());

    // One sweep.  Groups visited: forward 0 .. C-1, reverse C-2 .. 0 (the last group of the forward sweep would be
    // recomputed from unchanged inputs); after the first sweep of the solve a symmetric sweep also skips group 0 of
    // its forward half (it was the last group of the previous reverse half and nothing it reads has changed since):
    // both skips are bit-neutral.  Multicolour mode: the last group an ITERATION updates accumulates its own share
    // of sum |residual| (kernels.cuh k_gs_rows RES), and k_gs_resid covers the other rows only.
    // With processor patches neither shortcut holds: the halo values are refreshed at the start of every sweep (group
    // 0's interface rows DO change), and the last group's in-kernel residual would use the halo of the sweep's start.
    const bool fusedRes = !fixed && !multi && ctl->sweepMode == B200_SWEEP_MULTICOLOUR;
    // TWO colours, symGaussSeidel, multicolour order: the reverse half of a symmetric sweep degenerates -- it
    // recomputes one colour and updates the other, so a red-black "symmetric" sweep IS one red-black Gauss-Seidel
    // sweep, half as strong as upstream's forward + reverse pass.  It is executed as TWO red-black sweeps instead:
    // the same number of row updates as a symmetric sweep and about its strength (22 vs upstream's 24 sweeps on the
    // 16 M-cell test system), so that sweep counts and the maxIter cap keep their meaning.  The processor-patch
    // contributions stay refreshed once per COUNTED sweep, as upstream.
    const bool rb2 = sym && C == 2 && ctl->sweepMode == B200_SWEEP_MULTICOLOUR;
    const int inner = rb2 ? 2 : 1;                   // executed sweeps per counted sweep
    const bool back = sym && C >= 2 && !rb2;         // the sweep has a reverse half that ends on group 0
    const int lastGroup = back ? 0 : C - 1;
    // TWO colours (hex meshes): after the first sweep an iteration is [group F, group L] with F = the group the sweep
    // updates first; F's pass of the NEXT iteration delivers F's residual of this one for free (k_gs_rows RES == 2),
    // L's comes from its own pass (RES == 1): no residual kernel at all.  F's new values go to a shadow array
    // (ctx->w, free after the set-up) so that the old ones survive the decision; they alternate between the two.
    const bool lagged = fusedRes && C == 2 && ctx->gsLagged;
    const int gF = back ? 1 : 0, gL = 1 - gF;
    double* const arr[2] = {ctx->psi, ctx->w};
    int curF = 0;                                    // which array holds group F's current values
    bool firstSweep = true;
    auto sweep = [&](bool res, bool refresh = true) -> int {
        const int k0 = (back && !firstSweep && !multi) ? 1 : 0;
        firstSweep = false;
        if (refresh) RET(gs_sweep_halo(ctx, P, gh));
        for (int k = k0; k < C; ++k) RET(launch_gs_rows(ctx, P, k, (res && !back && k == C - 1) ? 1 : 0, gh));
        if (back) for (int k = C - 2; k >= 0; --k) RET(launch_gs_rows(ctx, P, k, (res && k == 0) ? 1 : 0, gh));
        return B200_OK;
    };
    // lagged form: one sweep of a two-colour plan.  lag: this sweep opens an iteration after the first one
    auto sweep2 = [&](int64_t body, bool lag, bool res) -> int {
        ctx->profIter = (int)(body + 1) * nS;        // (profile: loop bodies beyond the last executed one are dropped)
        if (firstSweep) {                            // groups 0, 1 [, 0]: everything in psi
            firstSweep = false;
            RET(launch_gs_rows(ctx, P, 0, 0, gh));
            RET(launch_gs_rows(ctx, P, 1, (res && !back) ? 1 : 0, gh));
            if (back) RET(launch_gs_rows(ctx, P, 0, res ? 1 : 0, gh));
            return B200_OK;
        }
        if (lag) {
            ctx->profIter = (int)body * nS;          // this pass completes the PREVIOUS iteration's residual
            RET(launch_gs_rows(ctx, P, gF, 2, gh, ctx->psi, arr[curF], arr[1 - curF]));
            ctx->profIter = (int)(body + 1) * nS;
            curF = 1 - curF;
        } else {
            RET(launch_gs_rows(ctx, P, gF, 0, gh, ctx->psi, nullptr, arr[curF]));
        }
        // group L: in place in psi, gathers F's current values
        RET(launch_gs_rows(ctx, P, gL, res ? 1 : 0, gh, arr[curF], nullptr, ctx->psi));
        return B200_OK;
    };
    if (fixed) {
        for (int i = 0; i < nS * inner; ++i) RET(sweep(false, i % inner == 0));
    } else {
        // rows whose residual is evaluated explicitly: all of them, or all but the group the iteration updated last
        int q0 = 0, q1 = N;
        if (fusedRes) {
            if (lastGroup == 0) q0 = P.h.colourStart[1];
            else q1 = P.h.colourStart[(size_t)C - 1];
            if (C == 1) q0 = q1 = 0;
        }
        // loop bodies the do/while can execute at most; the device decides when to stop (kernels return on S->done)
        const int64_t target = ctx->forceIters > 0 ? ctx->forceIters : std::max<int64_t>(ctl->maxIter, ctl->minIter);
        // (lagged: the residual of iteration k is completed by the first pass of body k + 1)
        const int64_t cap = std::max<int64_t>(1, (target + nS - 1) / nS) + (lagged ? 1 : 0);
        int64_t enq = 0;
        // a level-scheduled body is hundreds of launches: poll after every one; multicolour bodies are a few
        int chunk = (ctl->sweepMode == B200_SWEEP_EXACT && C > 16) ? 1 : 2;
        while (!ctx->hS->done && enq < cap) {
            const int n = (int)std::min<int64_t>(chunk, cap - enq);
            for (int i = 0; i < n; ++i) {
                ctx->profIter = (int)(enq + i + 1) * nS;
                const int nE = nS * inner;           // executed sweeps of this loop body
                if (lagged) {
                    for (int sw = 0; sw < nE; ++sw) RET(sweep2(enq + i, sw == 0 && enq + i > 0, sw == nE - 1));
                } else {
                    for (int sw = 0; sw < nE; ++sw) RET(sweep(fusedRes && sw == nE - 1, sw % inner == 0));
                    RET(launch_gs_resid(ctx, P, q0, q1, gh));
                }
            }
            ctx->profIter = 0;
            enq += n;
            CU(cudaMemcpyAsync(ctx->hS, S, sizeof(Scalars), cudaMemcpyDeviceToHost, ctx->sc));
            CU(cudaStreamSynchronize(ctx->sc));
            CU(cudaGetLastError());
            if (chunk < 16 && chunk > 1) chunk *= 2;
        }
        if (lagged) {
            // group F's values of the LAST COMPLETED iteration K: iteration 1 wrote them in psi, every later one in
            // the other array (the pass that found iteration K converged wrote iteration K + 1's elsewhere)
            const int K = ctx->hS->nIter / nS;
            if (K >= 1 && ((K - 1) & 1)) {
                const int f0 = P.h.colourStart[(size_t)gF], f1 = P.h.colourStart[(size_t)gF + 1];
                CU(cudaMemcpyAsync(ctx->psi + f0, ctx->w + f0, (size_t)(f1 - f0) * sizeof(double),
                                   cudaMemcpyDeviceToDevice, ctx->sc));
            }
        }
    }
    CU(cudaEventRecord(ctx->ev[2], ctx->sc));
    LAUNCH(PC_GATHER, k_scatter, grid_for(ctx, N), N, P.perm, ctx->psi, dn_psi);
    CU(cudaStreamSynchronize(ctx->sc));
    CU(cudaGetLastError());
    prof_collect(ctx, fixed ? 0x7fffffff : ctx->hS->nIter);
    const int rc = finish_solve(ctx, P, perf);
    if (fixed && perf) perf->nIterations = nS;       // solverPerf.nIterations() -= nSweeps_ (no residuals: 0, 0)
    return rc;
}

// ---- PBiCG + DILU (kernels.cuh "PBiCG + DILU") ------------------------------------------------
// y = A x (transposed: y = A^T x, i.e. lduMatrix::Tmul) on the plan's full-row ELL; INIT: also sumA
template <bool INIT>
int bicg_spmv(b200_ctx* ctx, DevPlan& P, bool transposed, const double* x, double* y, double* sA) {
    if (transposed) std::swap(P.val, P.valT);
    ctx->forceEll = true;
    const int rc = spmv_full<INIT, false>(ctx, P, x, y, sA, STEP_NONE);
    ctx->forceEll = false;
    if (transposed) std::swap(P.val, P.valT);
    return rc;
}

// w = M^-1 r with the (transposed) DILU factors: forward sweep over the groups ascending, backward descending
int bicg_dilu_apply(b200_ctx* ctx, DevPlan& P, const double* val, const double* rIn, double* w) {
    const int C = P.h.nColours;
    const EllCols E{P.col, P.col16, P.colBase};
    for (int k = 0; k < C; ++k) {
        int g;
        const ColourRows cr = colour_rows(ctx, P, k, &g);
        if (P.c16) { auto kf = k_tri_fwd<true>; LAUNCH(PC_BICG_TRI, kf, g, cr, P.sliceBase, P.rowLen, E, val, ctx->rD, rIn, w, ctx->S); }
        else { auto kf = k_tri_fwd<false>; LAUNCH(PC_BICG_TRI, kf, g, cr, P.sliceBase, P.rowLen, E, val, ctx->rD, rIn, w, ctx->S); }
    }
    for (int k = C - 2; k >= 0; --k) {
        int g;
        const ColourRows cr = colour_rows(ctx, P, k, &g);
        if (P.c16) { auto kb = k_tri_bwd<true>; LAUNCH(PC_BICG_TRI, kb, g, cr, P.sliceBase, P.rowLen, E, val, ctx->rD, w, ctx->S); }
        else { auto kb = k_tri_bwd<false>; LAUNCH(PC_BICG_TRI, kb, g, cr, P.sliceBase, P.rowLen, E, val, ctx->rD, w, ctx->S); }
    }
    return B200_OK;
}

// PBiCG::solve (OF-dev PBiCG.C); all pointers are device pointers in natural order.
// Buffers: rA = r, pA = p, wA = w, rT = t, wT = dT, pT = eD (the Eisenstat form's vectors, allocated on first use).
int bicg_core(b200_ctx* ctx, const double* dn_diag, const double* dn_upper, const double* dn_lower,
              const double* dn_src, double* dn_psi, const b200_controls* ctl, b200_perf* perf) {
    if (ctl->precond < B200_PRECOND_NONE || ctl->precond > B200_PRECOND_DILU_EXACT)
        return fail(ctx, B200_EINVAL, "bad preconditioner code for PBiCG (0 none, 1 diagonal, 2 DILU-class, 3 DILU exact)");
    if (ctl->reserved != 0) return fail(ctx, B200_EINVAL, "b200_controls.reserved must be 0");
    if (ctx->nranks > 1)
        return fail(ctx, B200_EUNSUPPORTED, "PBiCG with processor patches (nranks > 1) is not built yet");
    if (ctx->tileRows > 0) return fail(ctx, B200_EUNSUPPORTED, "PBiCG does not take tiled plans (B200PCG_TILE)");
    const bool dilu = ctl->precond >= B200_PRECOND_DILU_MC;
    DevPlan* Pp = nullptr;
    RET(ensure_plan(ctx, ctl->precond == B200_PRECOND_DILU_EXACT ? Ordering::Levels
                         : dilu ? Ordering::MultiColour : Ordering::Natural, &Pp));
    DevPlan& P = *Pp;
    const int N = ctx->N;
    const int gv = grid_for(ctx, (N + 1) / 2), gn = grid_for(ctx, N);
    Scalars* S = ctx->S;
    if (!P.valT) RET(dev_alloc(ctx, &P.valT, (size_t)P.h.nEntries));
    RET(ensure_eis_buffers(ctx, P));
    double *rA = ctx->r, *pA = ctx->p, *wA = ctx->w, *rT = ctx->t, *wT = ctx->dT, *pT = ctx->eD;

    CU(cudaEventRecord(ctx->ev[0], ctx->sc));
    RET(reset_scalars(ctx, ctl));
    RET(load_system_asym(ctx, P, dn_diag, dn_upper, dn_lower, dn_src, dn_psi, true));
    ctx->usedSmall = false;
    // wA = A psi (+ sumA -> pA), wT = A^T psi; normFactor, rA, initial residual (STEP_NORM); rT = source - wT
    RET(bicg_spmv<true>(ctx, P, false, ctx->psi, wA, pA));
    RET(bicg_spmv<false>(ctx, P, true, ctx->psi, wT, nullptr));
    {
        Reduce R = mkR(ctx, STEP_SUMPSI);
        LAUNCH(PC_SUM, k_sum, gv, N, ctx->psi, R);
    }
    {
        Reduce R = mkR(ctx, STEP_NORM);
        LAUNCH(PC_NORM, k_norm_resid, gv, N, wA, pA, ctx->src, rA, R);
    }
    LAUNCH(PC_BICG_VEC, k_sub, gn, N, ctx->src, wT, rT);
    // preconditioner set-up
    if (ctl->precond == B200_PRECOND_DIAGONAL) {
        LAUNCH(PC_RECIP, k_recip, gv, N, ctx->diag, ctx->rD);
    } else if (dilu) {
        const EllCols E{P.col, P.col16, P.colBase};
        for (int k = 0; k < P.h.nColours; ++k) {
            int g;
            const ColourRows cr = colour_rows(ctx, P, k, &g);
            if (P.c16) { auto kd = k_dilu_calc_rd<true>; LAUNCH(PC_DIC_RD, kd, g, cr, P.sliceBase, P.rowLen, E, P.val, P.valT, ctx->diag, ctx->rD); }
            else { auto kd = k_dilu_calc_rd<false>; LAUNCH(PC_DIC_RD, kd, g, cr, P.sliceBase, P.rowLen, E, P.val, P.valT, ctx->diag, ctx->rD); }
        }
        LAUNCH(PC_RECIP, k_recip, gv, N, ctx->rD, ctx->rD);
    }
    CU(cudaEventRecord(ctx->ev[1], ctx->sc));
    CU(cudaMemcpyAsync(ctx->hS, S, sizeof(Scalars), cudaMemcpyDeviceToHost, ctx->sc));
    CU(cudaStreamSynchronize(ctx->sc));
    CU(cudaGetLastError());

    // one loop body of PBiCG::solve
    auto body = [&]() -> int {
        const double *zA = rA, *zT = rT;             // `none`: the preconditioned residuals ARE the residuals
        if (ctl->precond == B200_PRECOND_DIAGONAL) {
            LAUNCH(PC_BICG_VEC, k_bicg_diag, gn, N, ctx->rD, rA, rT, wA, wT, S);
            zA = wA; zT = wT;
        } else if (dilu) {
            RET(bicg_dilu_apply(ctx, P, P.val, rA, wA));       // precondition(wA, rA)
            RET(bicg_dilu_apply(ctx, P, P.valT, rT, wT));      // preconditionT(wT, rT)
            zA = wA; zT = wT;
        }
        {
            Reduce R = mkR(ctx, STEP_WARA);                    // wArT = (wA, rT); beta
            LAUNCH(PC_BICG_DOT, k_dot2, gv, N, zA, rT, R);
        }
        LAUNCH(PC_BICG_VEC, k_bicg_p, gn, N, pA, zA, pT, zT, S);
        RET(bicg_spmv<false>(ctx, P, false, pA, wA, nullptr));
        RET(bicg_spmv<false>(ctx, P, true, pT, wT, nullptr));
        {
            Reduce R = mkR(ctx, STEP_WAPA);                    // wApT = (wA, pT); singularity; alpha
            LAUNCH(PC_BICG_DOT, k_dot2, gv, N, wA, pT, R);
        }
        {
            Reduce R = mkR(ctx, STEP_RES);                     // psi, rA, rT updates; final residual; loop condition
            LAUNCH(PC_BICG_VEC, k_bicg_r, gn, N, ctx->psi, pA, rA, wA, rT, wT, R);
        }
        return B200_OK;
    };
    const int64_t cap = ctx->forceIters > 0 ? ctx->forceIters : std::max<int64_t>((int64_t)ctl->maxIter + 1, ctl->minIter);
    int64_t enq = 0;
    int chunk = (ctl->precond == B200_PRECOND_DILU_EXACT && P.h.nColours > 16) ? 1 : 4;
    while (!ctx->hS->done && enq < cap) {
        const int n = (int)std::min<int64_t>(chunk, cap - enq);
        for (int i = 0; i < n; ++i) {
            ctx->profIter = (int)(enq + i + 1);
            RET(body());
        }
        ctx->profIter = 0;
        enq += n;
        CU(cudaMemcpyAsync(ctx->hS, S, sizeof(Scalars), cudaMemcpyDeviceToHost, ctx->sc));
        CU(cudaStreamSynchronize(ctx->sc));
        CU(cudaGetLastError());
        if (chunk > 1 && chunk < 64) chunk *= 2;
    }
    CU(cudaEventRecord(ctx->ev[2], ctx->sc));
    LAUNCH(PC_GATHER, k_scatter, grid_for(ctx, N), N, P.perm, ctx->psi, dn_psi);
    CU(cudaStreamSynchronize(ctx->sc));
    CU(cudaGetLastError());
    prof_collect(ctx, ctx->hS->nIter);
    return finish_solve(ctx, P, perf);
}

// ---- staged host <-> device copies for pageable caller memory ----------------------------------
// OpenFOAM's fields live in pageable memory: a plain cudaMemcpyAsync of them runs at ~12 GB/s host->device
// and ~5 GB/s device->host (measured through the host entry point with numpy arrays: 62 ms + 26 ms for the
// 766 MB + 128 MB of a 16 M-cell solve, profiles/r01_v12_perf_dic_eisenstat_final.log -- a fifth of the
// solve itself), against ~55 GB/s from page-locked memory.  Staged form (opt-in, B200PCG_STAGED_COPY=1; not
// yet run on a GPU): a few OpenMP threads copy 16 MB pieces into / out of two page-locked buffers while the
// DMA engine moves the previous piece.  Page-locked caller memory (bench.py e2e) always takes the direct copy.
constexpr size_t kStageBytes = (size_t)16 << 20;

bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

void par_memcpy(void* dst, const void* src, size_t n) {
    const int nt = std::max(1, std::min(8, omp_get_max_threads()));
    size_t piece = (n + (size_t)nt - 1) / (size_t)nt;
    piece = (piece + 4095) & ~(size_t)4095;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int t = 0; t < nt; ++t) {
        const size_t off = (size_t)t * piece;
        if (off < n) std::memcpy((char*)dst + off, (const char*)src + off, std::min(piece, n - off));
    }
}

int ensure_stage(b200_ctx* ctx) {
    for (int b = 0; b < 2; ++b) {
        if (!ctx->stageBuf[b]) CU(cudaHostAlloc(&ctx->stageBuf[b], kStageBytes, cudaHostAllocDefault));
        if (!ctx->stageEv[b]) CU(cudaEventCreateWithFlags(&ctx->stageEv[b], cudaEventDisableTiming));
    }
    return B200_OK;
}

int h2d(b200_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx->stagedCopy || bytes < ((size_t)4 << 20) || !is_pageable(src)) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->sc));
        return B200_OK;
    }
    RET(ensure_stage(ctx));
    size_t i = 0;
    for (size_t off = 0; off < bytes; off += kStageBytes, ++i) {
        const int b = (int)(i & 1);
        const size_t n = std::min(kStageBytes, bytes - off);
        CU(cudaEventSynchronize(ctx->stageEv[b]));   // the DMA that last read this piece is done
        par_memcpy(ctx->stageBuf[b], (const char*)src + off, n);
        CU(cudaMemcpyAsync((char*)dst + off, ctx->stageBuf[b], n, cudaMemcpyHostToDevice, ctx->sc));
        CU(cudaEventRecord(ctx->stageEv[b], ctx->sc));
    }
    return B200_OK;
}

int d2h(b200_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx->stagedCopy || bytes < ((size_t)4 << 20) || !is_pageable(dst)) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->sc));
        return B200_OK;
    }
    RET(ensure_stage(ctx));
    const size_t nPieces = (bytes + kStageBytes - 1) / kStageBytes;
    for (size_t i = 0; i <= nPieces; ++i) {
        if (i < nPieces) {   // DMA of piece i (its buffer was drained one step ago) ...
            const int b = (int)(i & 1);
            const size_t off = i * kStageBytes, n = std::min(kStageBytes, bytes - off);
            CU(cudaMemcpyAsync(ctx->stageBuf[b], (const char*)src + off, n, cudaMemcpyDeviceToHost, ctx->sc));
            CU(cudaEventRecord(ctx->stageEv[b], ctx->sc));
        }
        if (i >= 1) {        // ... overlaps the copy-out of piece i - 1
            const int b = (int)((i - 1) & 1);
            const size_t off = (i - 1) * kStageBytes, n = std::min(kStageBytes, bytes - off);
            CU(cudaEventSynchronize(ctx->stageEv[b]));
            par_memcpy((char*)dst + off, ctx->stageBuf[b], n);
        }
    }
    return B200_OK;
}

// ---- mesh fingerprint + cross-rank agreement -------------------------------------------------------
// b200_set_addressing is called before every solve (upstream constructs a solver object per solve), so the
// "same mesh?" test must cost microseconds: sizes, every interface's size and neighbour rank, and a 64-bit
// hash over a strided sample of lowerAddr / upperAddr / faceCells (every entry when the array has
// <= 65536 of them).  A renumbered, refined or re-decomposed mesh that happens to live at the same address with
// the same sizes changes practically every sampled label; mesh_key alone (the address of the lduAddressing
// object in the adapter) does not see that.
inline uint64_t mix64(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
    return h ^ (h >> 29);
}
uint64_t sample_hash(uint64_t h, const int32_t* a, int64_t n) {
    if (n <= 0 || !a) return mix64(h, 0x5bd1e995u);
    const int64_t stride = n <= 65536 ? 1 : n / 8192;
    for (int64_t i = 0; i < n; i += stride) h = mix64(h, (uint64_t)(uint32_t)a[i] ^ ((uint64_t)i << 32));
    for (int64_t i = std::max<int64_t>(0, n - 64); i < n; ++i) h = mix64(h, (uint64_t)(uint32_t)a[i]);
    return h;
}
uint64_t mesh_fingerprint(int32_t nCells, int32_t nFaces, const int32_t* l, const int32_t* u, int32_t nIfaces,
                          const b200_iface* ifaces) {
    uint64_t h = mix64(0x243F6A8885A308D3ull, ((uint64_t)(uint32_t)nCells << 32) | (uint32_t)nFaces);
    h = sample_hash(h, l, nFaces);
    h = sample_hash(h, u, nFaces);
    h = mix64(h, (uint64_t)(uint32_t)nIfaces);
    for (int k = 0; k < nIfaces; ++k) {
        h = mix64(h, ((uint64_t)(uint32_t)ifaces[k].nbrRank << 32) | (uint32_t)ifaces[k].nFaces);
        h = sample_hash(h, ifaces[k].nFaces > 0 ? ifaces[k].faceCells : nullptr, ifaces[k].nFaces);
    }
    return h;
}

// min over ranks of up to 4 ints (one tiny ncclAllReduce + host sync; set-up paths only).  Every rank must
// call it at the same point, WHATEVER its local status: rank-local failures are agreed on here so that no rank
// returns early while its peers wait inside a later collective.
int agree_min(b200_ctx* ctx, int* v, int n) {
    if (ctx->nranks == 1) return B200_OK;
    int* d = reinterpret_cast<int*>(ctx->partials);
    CU(cudaMemcpyAsync(d, v, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, ctx->sc));
    NC(g_nccl.AllReduce(d, d + 8, (size_t)n, ncclInt, ncclMin, ctx->comm, ctx->sc));
    CU(cudaMemcpyAsync(v, d + 8, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->sc));
    CU(cudaMemsetAsync(ctx->partials, 0, sizeof(double) * 16, ctx->sc));
    CU(cudaStreamSynchronize(ctx->sc));
    return B200_OK;
}

// Map every rank's PeerBuf into this process (CUDA IPC; handles exchanged with ncclAllGather).
// Collective: every rank calls ncclAllGather exactly once, also after a local failure (its record then
// carries valid = 0 and nobody maps anything), so a rank that cannot allocate or export its buffer cannot leave
// the others hanging inside the gather.  Returns the LOCAL outcome; the caller agrees on the global one.
bool setup_peer_reduce(b200_ctx* c, std::string& why) {
    const int n = c->nranks;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    struct Rec { cudaIpcMemHandle_t h; uint64_t valid; };
    Rec mine;
    std::memset(&mine, 0, sizeof(mine));
    bool ok = true;
    cudaError_t e;
    if ((e = cudaMalloc((void**)&c->peerLocal, sizeof(PeerBuf))) != cudaSuccess) { why = cudaGetErrorString(e); cudaGetLastError(); c->peerLocal = nullptr; ok = false; }
    if (ok && (e = cudaMemset(c->peerLocal, 0, sizeof(PeerBuf))) != cudaSuccess) { why = cudaGetErrorString(e); cudaGetLastError(); ok = false; }
    if (ok && (e = cudaIpcGetMemHandle(&mine.h, c->peerLocal)) != cudaSuccess) { why = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e); cudaGetLastError(); ok = false; }
    mine.valid = ok ? 1 : 0;
    // the gather's own buffers live in the context's partials arena (allocated at create): no allocation that
    // could fail on one rank only stands between here and the collective
    static_assert(sizeof(Rec) * (kMaxRanks + 1) <= sizeof(double) * kNSums * kMaxGrid, "partials arena too small");
    unsigned char* d_send = reinterpret_cast<unsigned char*>(c->partials);
    unsigned char* d_recv = d_send + sizeof(Rec);
    std::vector<Rec> all((size_t)n);
    cudaError_t e1 = cudaMemcpyAsync(d_send, &mine, sizeof(Rec), cudaMemcpyHostToDevice, c->sc);
    ncclResult_t r = g_nccl.AllGather(d_send, d_recv, sizeof(Rec), ncclUint8, c->comm, c->sc);
    cudaError_t e2 = cudaMemcpyAsync(all.data(), d_recv, sizeof(Rec) * (size_t)n, cudaMemcpyDeviceToHost, c->sc);
    cudaError_t e3 = cudaStreamSynchronize(c->sc);
    cudaMemsetAsync(c->partials, 0, sizeof(Rec) * (size_t)(n + 1), c->sc);
    cudaStreamSynchronize(c->sc);
    if (r != ncclSuccess) { why = g_nccl.GetErrorString(r); ok = false; }
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { why = "copy around ncclAllGather failed"; cudaGetLastError(); ok = false; }
    for (int k = 0; ok && k < n; ++k)
        if (!all[k].valid) { why = "rank " + std::to_string(k) + " could not export its buffer"; ok = false; }
    std::vector<PeerBuf*> ptrs((size_t)n, nullptr);
    for (int k = 0; ok && k < n; ++k) {
        if (k == c->rank) { ptrs[k] = c->peerLocal; continue; }
        void* p = nullptr;
        e = cudaIpcOpenMemHandle(&p, all[k].h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { why = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); cudaGetLastError(); ok = false; break; }
        c->peerMapped.push_back(p);
        ptrs[k] = (PeerBuf*)p;
    }
    if (ok) {
        if (cudaMalloc((void**)&c->d_peers, sizeof(PeerBuf*) * (size_t)n) != cudaSuccess) { why = "cudaMalloc"; cudaGetLastError(); ok = false; }
        else if (cudaMemcpy(c->d_peers, ptrs.data(), sizeof(PeerBuf*) * (size_t)n, cudaMemcpyHostToDevice) != cudaSuccess) { why = "cudaMemcpy"; cudaGetLastError(); ok = false; }
    }
    c->p2pReduce = ok;
    return ok;
}

// Peer-memory halo buffers for the CURRENT mesh (collective: called by every rank from b200_set_addressing after
// the ranks agreed that the plan build succeeded).  Every rank always takes part in the one ncclAllGather and
// in the agreement that follows; a local failure (allocation, export, import, a pair of ranks that shares two
// patches, mismatched patch sizes) only clears p2pHalo -- on EVERY rank, because the minimum is taken -- and
// the NCCL send/recv path is used instead.  Returns an error only when the collectives themselves fail.
int setup_peer_halo(b200_ctx* ctx, DevPlan& P) {
    const int n = ctx->nranks;
    const size_t nSlots = (size_t)ctx->nSlots;
    struct Rec {
        cudaIpcMemHandle_t h;
        int32_t valid, nSlots;
        int32_t off[kMaxRanks];   // slot offset of my patch towards rank q (-1: none)
        int32_t cnt[kMaxRanks];
    };
    static_assert(sizeof(Rec) * (kMaxRanks + 1) <= sizeof(double) * kNSums * kMaxGrid, "partials arena too small");
    Rec mine;
    std::memset(&mine, 0, sizeof(mine));
    for (int q = 0; q < kMaxRanks; ++q) mine.off[q] = -1;
    // the exchange counters restart with every mesh, on every rank: two ranks that become neighbours only now
    // may have executed different numbers of exchanges before (the fresh buffers' flags are zero as well)
    CU(cudaMemsetAsync(reinterpret_cast<char*>(ctx->S) + offsetof(Scalars, haloSeq), 0,
                       sizeof(Scalars) - offsetof(Scalars, haloSeq), ctx->sc));
    bool ok = !ctx->forceNcclHalo && ctx->p2pReduce && n <= kMaxRanks;   // p2pReduce: IPC mapping is known to work
    std::string why;
    mine.nSlots = (int32_t)nSlots;
    for (int k = 0; ok && k < P.h.nIfaces; ++k) {
        const int q = P.h.nbrRank[k];
        if (mine.off[q] >= 0) { ok = false; why = "two patches towards the same rank"; break; }
        mine.off[q] = P.h.patchStart[k];
        mine.cnt[q] = P.h.patchStart[k + 1] - P.h.patchStart[k];
    }
    const size_t bytes = sizeof(double) * 2 * nSlots + sizeof(unsigned long long) * 2 * kMaxRanks;
    if (ok) {
        cudaError_t e = cudaMalloc((void**)&ctx->haloLocal, bytes);
        if (e == cudaSuccess) e = cudaMemset(ctx->haloLocal, 0, bytes);
        if (e == cudaSuccess) e = cudaIpcGetMemHandle(&mine.h, ctx->haloLocal);
        if (e != cudaSuccess) { ok = false; why = cudaGetErrorString(e); cudaGetLastError(); }
    }
    mine.valid = ok ? 1 : 0;
    unsigned char* d_send = reinterpret_cast<unsigned char*>(ctx->partials);
    unsigned char* d_recv = d_send + sizeof(Rec);
    std::vector<Rec> all((size_t)n);
    CU(cudaMemcpyAsync(d_send, &mine, sizeof(Rec), cudaMemcpyHostToDevice, ctx->sc));
    NC(g_nccl.AllGather(d_send, d_recv, sizeof(Rec), ncclUint8, ctx->comm, ctx->sc));
    CU(cudaMemcpyAsync(all.data(), d_recv, sizeof(Rec) * (size_t)n, cudaMemcpyDeviceToHost, ctx->sc));
    CU(cudaStreamSynchronize(ctx->sc));
    CU(cudaMemsetAsync(ctx->partials, 0, sizeof(Rec) * (size_t)(n + 1), ctx->sc));
    CU(cudaStreamSynchronize(ctx->sc));
    for (int k = 0; ok && k < n; ++k)
        if (!all[k].valid) { ok = false; why = "rank " + std::to_string(k) + " has no exportable buffer"; }
    if (ok && nSlots > 0) {
        std::vector<double*> dst(2 * nSlots, nullptr);
        std::vector<unsigned long long*> nbrFlag;
        std::vector<int32_t> nbrRanks;
        for (int k = 0; ok && k < P.h.nIfaces; ++k) {
            const int q = P.h.nbrRank[k];
            const int off = P.h.patchStart[k], cnt = P.h.patchStart[k + 1] - off;
            if (all[q].off[ctx->rank] < 0 || all[q].cnt[ctx->rank] != cnt) {
                ok = false; why = "processor patch towards rank " + std::to_string(q) + " has no counterpart of the same size";
                break;
            }
            void* base = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&base, all[q].h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { ok = false; why = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); cudaGetLastError(); break; }
            ctx->haloMapped.push_back(base);
            double* vals = reinterpret_cast<double*>(base);
            const size_t qSlots = (size_t)all[q].nSlots;
            unsigned long long* flags = reinterpret_cast<unsigned long long*>(vals + 2 * qSlots);
            for (int par = 0; par < 2; ++par)
                for (int j = 0; j < cnt; ++j)
                    dst[(size_t)par * nSlots + (size_t)(off + j)] = vals + (size_t)par * qSlots + (size_t)(all[q].off[ctx->rank] + j);
            nbrRanks.push_back(q);
            nbrFlag.push_back(flags + ctx->rank);                 // flags[0][me] ...
        }
        if (ok) {
            const size_t nNbr = nbrRanks.size();
            std::vector<unsigned long long*> nf(2 * nNbr);
            for (size_t k = 0; k < nNbr; ++k) { nf[k] = nbrFlag[k]; nf[nNbr + k] = nbrFlag[k] + kMaxRanks; }   // ... flags[1][me]
            int rc = upload(ctx, &ctx->d_haloDst, dst);
            if (rc == B200_OK) rc = upload(ctx, &ctx->d_haloNbrFlag, nf);
            if (rc == B200_OK) rc = upload(ctx, &ctx->d_haloNbrRanks, nbrRanks);
            if (rc == B200_OK && cudaStreamSynchronize(ctx->sc) != cudaSuccess) rc = B200_ECUDA;
            if (rc != B200_OK) { ok = false; why = "upload of the halo tables failed"; cudaGetLastError(); }
            else
                ctx->haloDev = Halo{ctx->d_haloDst, ctx->d_haloNbrFlag,
                                    reinterpret_cast<const unsigned long long*>(ctx->haloLocal + 2 * nSlots),
                                    ctx->haloLocal, ctx->d_haloNbrRanks, (int)nSlots, (int)nNbr};
        }
    }
    int v = ok ? 1 : 0;
    RET(agree_min(ctx, &v, 1));
    if (!ok && !why.empty() && !ctx->forceNcclHalo)
        fprintf(stderr, "b200pcg[%d]: peer-memory halo exchange unavailable (%s); using NCCL send/recv\n", ctx->rank, why.c_str());
    if (v == 1) {
        ctx->p2pHalo = true;
    } else {
        // keep the export alive until every rank has closed its imports (next agreement below), then drop all
        for (void* p : ctx->haloMapped) cudaIpcCloseMemHandle(p);
        ctx->haloMapped.clear();
        int z = 1;
        RET(agree_min(ctx, &z, 1));
        free_halo(ctx);
    }
    return B200_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

int b200_abi_version(void) { return B200_ABI_VERSION; }

int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char* b200_last_error(const b200_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_createError.c_str();
}

int b200_get_unique_id(void* uid128) {
    b200_ctx* ctx = nullptr;
    if (!uid128) return fail(ctx, B200_EINVAL, "null uid buffer");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (!g_nccl.load()) return fail(ctx, B200_ENCCL, g_nccl.error);
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    std::memcpy(uid128, &id, 128);
    return B200_OK;
}

int b200_ctx_create(int device, int rank, int nranks, const void* nccl_uid, b200_ctx** out) {
    b200_ctx* ctx = nullptr;
    if (!out) return fail(ctx, B200_EINVAL, "null out pointer");
    *out = nullptr;
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, B200_EINVAL, "bad rank/nranks");
    if (nranks > 1 && !nccl_uid) return fail(ctx, B200_EINVAL, "nranks > 1 needs an NCCL unique id");
    int ndev = b200_device_count();
    if (ndev <= 0)
        return fail(ctx, B200_ENODEVICE, "no CUDA device: libb200pcg has no CPU fallback");
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= ndev) return fail(ctx, B200_EINVAL, "device index out of range");
    b200_ctx* c = new b200_ctx();
    c->device = device;
    c->rank = rank;
    c->nranks = nranks;
    auto bail = [&](int code, const std::string& m) {
        g_createError = m;
        b200_ctx_destroy(c);
        return code;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(B200_ECUDA, cudaGetErrorString(e));
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return bail(B200_ECUDA, cudaGetErrorString(e));
    if (prop.major < 10)
        return bail(B200_ENODEVICE, std::string("device is ") + prop.name +
                                        " (sm_" + std::to_string(prop.major * 10 + prop.minor) +
                                        "); libb200pcg is built for sm_100a only");
    c->numSMs = prop.multiProcessorCount;
    if (const char* e3 = getenv("B200PCG_STAGES")) c->symStages = std::max(2, std::min(3, atoi(e3)));
    if (const char* e4 = getenv("B200PCG_CTAS")) c->symPerSM = atoi(e4);
    if (const char* e2 = getenv("B200PCG_SPMV")) {
        c->disableSym = (std::string(e2) == "ell");
        c->disableTma = (std::string(e2) == "sym");
        c->enableSr = (std::string(e2) == "sr");
    }
    if (const char* e9 = getenv("B200PCG_SMALL_N")) c->smallN = std::max(0, atoi(e9));
    if (const char* e11 = getenv("B200PCG_SMALL_FAST")) c->disableFast = atoi(e11) == 0;
    if (const char* e14 = getenv("B200PCG_FAST_CTAS")) c->fastMaxCtas = std::max(1, atoi(e14));
    if (const char* e10 = getenv("B200PCG_SMALL_CTAS")) c->smallCtas = std::max(1, std::min(kSmallMaxCtas, atoi(e10)));
    if (const char* e15 = getenv("B200PCG_COL16")) c->disableCol16 = atoi(e15) == 0;
    if (const char* eg = getenv("B200PCG_GS_CTAS")) c->gsCtas = atoi(eg);
    if (const char* el = getenv("B200PCG_GS_LAGGED")) c->gsLagged = atoi(el) != 0;
    if (const char* e13 = getenv("B200PCG_TILE")) c->tileRows = std::max(0, atoi(e13));
    if (const char* e12 = getenv("B200PCG_FUSE_FIRST")) c->noFuseFirst = atoi(e12) == 0;
    if (const char* e17 = getenv("B200PCG_EIS_BATCH")) c->eisBatch = atoi(e17) != 0;
    if (const char* e22 = getenv("B200PCG_SORT_COLS")) c->sortCols = atoi(e22) != 0;
    if (const char* e21 = getenv("B200PCG_STAGED_COPY")) c->stagedCopy = atoi(e21) != 0;
    if (const char* e19 = getenv("B200PCG_EIS_OVERLAP")) c->eisOverlap = atoi(e19) != 0;
    if (const char* e20 = getenv("B200PCG_DIC")) c->dicDefaultEis = (std::string(e20) != "multicolour");
    if (const char* e18 = getenv("B200PCG_EIS_CTAS")) c->eisCtas = atoi(e18) == 3 ? 3 : (atoi(e18) == 4 ? 4 : 0);
    if (const char* e16 = getenv("B200PCG_SWEEP_CTAS")) {
        c->sweepPerSM = std::max(1, std::min(16, atoi(e16)));
        c->sweepPerSMSet = true;
    }
    if (const char* e7 = getenv("B200PCG_EXACT")) c->exactWidth = atoi(e7) != 0;
    if (const char* e23 = getenv("B200PCG_GRAPH")) c->useGraph = atoi(e23) != 0;
    if (const char* e24 = getenv("B200PCG_HALO")) c->forceNcclHalo = (std::string(e24) == "nccl");
    if (const char* e25 = getenv("B200PCG_FUSE_IFACE")) c->fuseIface = atoi(e25) != 0;
    if (const char* e26 = getenv("B200PCG_GRAPH_MULTI")) c->graphMulti = atoi(e26) != 0;
    if (const char* e27 = getenv("B200PCG_SPLIT_IFACE")) c->splitIface = atoi(e27) != 0;
    if (const char* e8 = getenv("B200PCG_RENUMBER"))
        c->renumber = (std::string(e8) == "auto") ? (int)Renumber::Auto : (atoi(e8) != 0 ? (int)Renumber::Force : (int)Renumber::Off);
    if ((e = cudaStreamCreateWithFlags(&c->sc, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->sm, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->evPack, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->evRecv, cudaEventDisableTiming)) != cudaSuccess)
        return bail(B200_ECUDA, cudaGetErrorString(e));
    for (auto& ev : c->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(B200_ECUDA, cudaGetErrorString(e));
    if ((e = cudaMalloc((void**)&c->S, sizeof(Scalars))) != cudaSuccess ||
        (e = cudaMalloc((void**)&c->partials, sizeof(double) * kNSums * kMaxGrid)) != cudaSuccess ||
        (e = cudaMalloc((void**)&c->partials2, sizeof(double) * kMaxGrid)) != cudaSuccess ||
        (e = cudaHostAlloc((void**)&c->hS, sizeof(Scalars), cudaHostAllocDefault)) != cudaSuccess)
        return bail(B200_ECUDA, cudaGetErrorString(e));
    cudaMemset(c->S, 0, sizeof(Scalars));
    cudaMemset(c->partials, 0, sizeof(double) * kNSums * kMaxGrid);
    if (nranks > 1) {
        ncclUniqueId id;
        std::memcpy(&id, nccl_uid, 128);
        if (!g_nccl.load()) return bail(B200_ENCCL, g_nccl.error);
        ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
        if (r != ncclSuccess) return bail(B200_ENCCL, g_nccl.GetErrorString(r));
        const char* ar = getenv("B200PCG_ALLREDUCE");
        if (!(ar && std::string(ar) == "nccl") && nranks <= kMaxRanks) {
            std::string why;
            if (!setup_peer_reduce(c, why)) {
                // not fatal: fall back to ncclAllReduce, but every rank must take the same path
                fprintf(stderr, "b200pcg[%d]: peer-memory all-reduce unavailable (%s); using NCCL\n", rank, why.c_str());
            }
            // agree across ranks (min over ranks of the local success flag); a failure of the agreement itself
            // is fatal: the ranks could otherwise disagree on the reduction path
            int ok = c->p2pReduce ? 1 : 0;
            b200_ctx* ctx = c;
            int rcA = [&]() -> int { return agree_min(ctx, &ok, 1); }();
            if (rcA != B200_OK) return bail(rcA, "peer-memory set-up: cross-rank agreement failed: " + c->err);
            c->p2pReduce = (ok == 1);
        }
    }
    *out = c;
    return B200_OK;
}

void b200_ctx_destroy(b200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->sc) cudaStreamSynchronize(c->sc);
    if (c->sm) cudaStreamSynchronize(c->sm);
    free_mesh(c);
    for (void* p : c->peerMapped) cudaIpcCloseMemHandle(p);
    if (c->d_peers) cudaFree(c->d_peers);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    if (c->peerLocal) cudaFree(c->peerLocal);
    for (auto ev : c->profPool) cudaEventDestroy(ev);
    dev_free(c->S);
    dev_free(c->partials);
    dev_free(c->partials2);
    if (c->hS) cudaFreeHost(c->hS);
    for (int b = 0; b < 2; ++b) {
        if (c->stageBuf[b]) cudaFreeHost(c->stageBuf[b]);
        if (c->stageEv[b]) cudaEventDestroy(c->stageEv[b]);
    }
    if (c->evPack) cudaEventDestroy(c->evPack);
    if (c->evRecv) cudaEventDestroy(c->evRecv);
    for (auto ev : c->ev) if (ev) cudaEventDestroy(ev);
    if (c->sc) cudaStreamDestroy(c->sc);
    if (c->sm) cudaStreamDestroy(c->sm);
    delete c;
}

int b200_set_addressing(b200_ctx* ctx, uint64_t mesh_key, int32_t nCells, int32_t nFaces,
                        const int32_t* lowerAddr, const int32_t* upperAddr, int32_t nIfaces,
                        const b200_iface* ifaces) {
    if (!ctx) return B200_EINVAL;
    CU(cudaSetDevice(ctx->device));
    // --- phase 1: everything rank-local (validation, "same mesh?" test); no early return when nranks > 1:
    //     a rank that fails or rebuilds alone would leave its peers inside (or outside) the collectives below
    int rcLocal = B200_OK;
    std::string msgLocal;
    auto bad = [&](int code, const char* m) { if (rcLocal == B200_OK) { rcLocal = code; msgLocal = m; } };
    if (nCells < 0 || nFaces < 0 || nIfaces < 0) bad(B200_EINVAL, "negative size");
    else if (nFaces > 0 && (!lowerAddr || !upperAddr)) bad(B200_EINVAL, "null addressing");
    else if (nIfaces > 0 && !ifaces) bad(B200_EINVAL, "null interface list");
    else if (ctx->nranks == 1 && nIfaces > 0) bad(B200_EINVAL, "processor interfaces given but the context has nranks == 1");
    for (int k = 0; rcLocal == B200_OK && k < nIfaces; ++k) {
        if (ifaces[k].nFaces < 0 || (ifaces[k].nFaces > 0 && !ifaces[k].faceCells)) bad(B200_EINVAL, "bad interface");
        else if (ifaces[k].nbrRank < 0 || ifaces[k].nbrRank >= ctx->nranks || ifaces[k].nbrRank == ctx->rank)
            bad(B200_EINVAL, "interface neighbour rank out of range");
    }
    uint64_t fp = 0;
    bool same = false;
    if (rcLocal == B200_OK) {
        fp = mesh_fingerprint(nCells, nFaces, lowerAddr, upperAddr, nIfaces, ifaces);
        same = ctx->haveMesh && ctx->meshKey == mesh_key && ctx->N == nCells && ctx->F == nFaces &&
               (int32_t)ctx->hif.size() == nIfaces && ctx->meshFingerprint == fp;
    }
    if (ctx->nranks == 1) {
        if (rcLocal != B200_OK) return fail(ctx, rcLocal, msgLocal);
        if (same) return B200_OK;
    } else {
        int v[2] = {rcLocal == B200_OK ? 1 : 0, same ? 1 : 0};
        RET(agree_min(ctx, v, 2));
        if (rcLocal != B200_OK) return fail(ctx, rcLocal, msgLocal);
        if (!v[0]) return fail(ctx, B200_EINVAL, "set_addressing failed on another rank");
        if (v[1]) return B200_OK;      // unchanged on EVERY rank; otherwise every rank rebuilds
    }
    // --- phase 2: rank-local build
    CU(cudaStreamSynchronize(ctx->sc));
    free_mesh(ctx);
    ctx->N = nCells;
    ctx->F = nFaces;
    ctx->hl.assign(lowerAddr, lowerAddr + nFaces);
    ctx->hu.assign(upperAddr, upperAddr + nFaces);
    ctx->hif.resize((size_t)nIfaces);
    ctx->nSlots = 0;
    for (int k = 0; k < nIfaces; ++k) {
        ctx->hif[k].nbrRank = ifaces[k].nbrRank;
        ctx->hif[k].faceCells.assign(ifaces[k].faceCells, ifaces[k].faceCells + ifaces[k].nFaces);
        ctx->nSlots += ifaces[k].nFaces;
    }
    auto build = [&]() -> int {
        DevPlan* P = nullptr;
        RET(ensure_plan(ctx, Ordering::Natural, &P));
        RET(upload(ctx, &ctx->d_l, ctx->hl));
        RET(upload(ctx, &ctx->d_u, ctx->hu));
        if (P->h.renumbered) {
            const size_t N = (size_t)ctx->N, F = (size_t)ctx->F;
            std::vector<int32_t> os(N + 1, 0), ls(N + 1, 0), lo(F);
            for (size_t f = 0; f < F; ++f) { os[(size_t)ctx->hl[f] + 1]++; ls[(size_t)ctx->hu[f] + 1]++; }
            for (size_t c = 0; c < N; ++c) { os[c + 1] += os[c]; ls[c + 1] += ls[c]; }
            std::vector<int32_t> pos(ls.begin(), ls.end() - 1);
            for (size_t f = 0; f < F; ++f) lo[(size_t)pos[(size_t)ctx->hu[f]]++] = (int32_t)f;   // stable: ascending faces
            RET(upload(ctx, &ctx->d_ownerStart, os));
            RET(upload(ctx, &ctx->d_losortStart, ls));
            RET(upload(ctx, &ctx->d_losort, lo));
            CU(cudaStreamSynchronize(ctx->sc));   // the host vectors go out of scope
        }
        RET(alloc_vectors(ctx));
        CU(cudaStreamSynchronize(ctx->sc));
        return B200_OK;
    };
    int rc = build();
    // --- phase 3: agree on the outcome, then the one data collective (global cell count for gAverage)
    ctx->nGlobalCells = (double)nCells;
    if (ctx->nranks > 1) {
        const std::string keep = ctx->err;
        int ok = rc == B200_OK ? 1 : 0;
        int rcA = agree_min(ctx, &ok, 1);
        if (rc != B200_OK) { ctx->err = keep; free_mesh(ctx); return rc; }
        if (rcA != B200_OK) { free_mesh(ctx); return rcA; }
        if (!ok) { free_mesh(ctx); return fail(ctx, B200_EINVAL, "set_addressing: plan build failed on another rank"); }
        double* tmp = ctx->partials;
        double h = (double)nCells;
        CU(cudaMemcpyAsync(tmp, &h, sizeof(double), cudaMemcpyHostToDevice, ctx->sc));
        NC(g_nccl.AllReduce(tmp, tmp + 1, 1, ncclDouble, ncclSum, ctx->comm, ctx->sc));
        CU(cudaMemcpyAsync(&h, tmp + 1, sizeof(double), cudaMemcpyDeviceToHost, ctx->sc));
        CU(cudaMemsetAsync(tmp, 0, 2 * sizeof(double), ctx->sc));
        CU(cudaStreamSynchronize(ctx->sc));
        ctx->nGlobalCells = h;
        RET(setup_peer_halo(ctx, ctx->plans[0]));
    } else if (rc != B200_OK) {
        free_mesh(ctx);
        return rc;
    }
    ctx->meshKey = mesh_key;
    ctx->meshFingerprint = fp;
    ctx->haveMesh = true;
    return B200_OK;
}

int b200_assemble_laplacian_device(b200_ctx* ctx, const double* g, const double* s, const double* d,
                                   double sign, double* upper_out, double* diag_inout) {
    if (!ctx) return B200_EINVAL;
    if (!ctx->haveMesh) return fail(ctx, B200_ESTATE, "assemble before set_addressing");
    if ((ctx->F > 0 && (!g || !s || !d || !upper_out)) || (ctx->N > 0 && !diag_inout))
        return fail(ctx, B200_EINVAL, "null argument");
    CU(cudaSetDevice(ctx->device));
    DevPlan& P = ctx->plans[0];
    LAUNCH(PC_ASM_FACE, k_face_coeff, grid_for(ctx, ctx->F, 16), ctx->F, g, s, d, sign, upper_out);
    if (ctx->d_losort)   // renumbered plan: in the caller's cell order (kernels.cuh)
        LAUNCH(PC_ASM_DIAG, k_neg_sum_diag_nat, grid_for(ctx, ctx->N, 16), ctx->N, ctx->d_ownerStart, ctx->d_losortStart,
               ctx->d_losort, upper_out, diag_inout);
    else
        LAUNCH(PC_ASM_DIAG, k_neg_sum_diag, grid_for(ctx, ctx->N, 16), ctx->N, P.sliceBase, P.rowLen,
               P.faceOf, P.perm, upper_out, diag_inout);
    CU(cudaStreamSynchronize(ctx->sc));
    CU(cudaGetLastError());
    prof_collect(ctx);
    return B200_OK;
}

int b200_assemble_laplacian(b200_ctx* ctx, const double* g, const double* s, const double* d,
                            double sign, double* upper_out, double* diag_inout) {
    if (!ctx) return B200_EINVAL;
    if (!ctx->haveMesh) return fail(ctx, B200_ESTATE, "assemble before set_addressing");
    if ((ctx->F > 0 && (!g || !s || !d || !upper_out)) || (ctx->N > 0 && !diag_inout))
        return fail(ctx, B200_EINVAL, "null argument");
    CU(cudaSetDevice(ctx->device));
    RET(ensure_staging(ctx, true));
    const size_t fb = (size_t)ctx->F * sizeof(double), nb = (size_t)ctx->N * sizeof(double);
    CU(cudaMemcpyAsync(ctx->in_f1, g, fb, cudaMemcpyHostToDevice, ctx->sc));
    CU(cudaMemcpyAsync(ctx->in_f2, s, fb, cudaMemcpyHostToDevice, ctx->sc));
    CU(cudaMemcpyAsync(ctx->in_f3, d, fb, cudaMemcpyHostToDevice, ctx->sc));
    CU(cudaMemcpyAsync(ctx->in_diag, diag_inout, nb, cudaMemcpyHostToDevice, ctx->sc));
    RET(b200_assemble_laplacian_device(ctx, ctx->in_f1, ctx->in_f2, ctx->in_f3, sign, ctx->in_upper,
                                       ctx->in_diag));
    CU(cudaMemcpyAsync(upper_out, ctx->in_upper, fb, cudaMemcpyDeviceToHost, ctx->sc));
    CU(cudaMemcpyAsync(diag_inout, ctx->in_diag, nb, cudaMemcpyDeviceToHost, ctx->sc));
    CU(cudaStreamSynchronize(ctx->sc));
    return B200_OK;
}

int b200_set_boundary_faces(b200_ctx* ctx, int32_t nB, const int32_t* bCells) {
    if (!ctx) return B200_EINVAL;
    if (!ctx->haveMesh) return fail(ctx, B200_ESTATE, "set_boundary_faces before set_addressing");
    if (nB < 0 || (nB > 0 && !bCells)) return fail(ctx, B200_EINVAL, "bad boundary face list");
    CU(cudaSetDevice(ctx->device));
    if (ctx->bfStart && ctx->nB == nB && (nB == 0 || std::memcmp(ctx->hbCells.data(), bCells, sizeof(int32_t) * (size_t)nB) == 0))
        return B200_OK;
    const int32_t N = ctx->N;
    std::vector<int32_t> start((size_t)N + 1, 0), order((size_t)nB);
    for (int32_t b = 0; b < nB; ++b) {
        if (bCells[b] < 0 || bCells[b] >= N) return fail(ctx, B200_EINVAL, "boundary faceCells out of range");
        start[(size_t)bCells[b] + 1]++;
    }
    for (int32_t c = 0; c < N; ++c) start[(size_t)c + 1] += start[c];
    std::vector<int32_t> pos(start.begin(), start.end() - 1);
    for (int32_t b = 0; b < nB; ++b) order[(size_t)pos[bCells[b]]++] = b;   // stable: patch order inside a cell
    CU(cudaStreamSynchronize(ctx->sc));
    dev_free(ctx->bfStart);
    dev_free(ctx->bfOrder);
    RET(upload(ctx, &ctx->bfStart, start));
    RET(upload(ctx, &ctx->bfOrder, order));
    CU(cudaStreamSynchronize(ctx->sc));
    ctx->hbCells.assign(bCells, bCells + nB);
    ctx->nB = nB;
    return B200_OK;
}

int b200_assemble_p_rgh_device(b200_ctx* ctx, const b200_prgh_terms* t, double* upper_out, double* diag_out,
                               double* source_out) {
    if (!ctx) return B200_EINVAL;
    if (!ctx->haveMesh) return fail(ctx, B200_ESTATE, "assemble before set_addressing");
    if (!t) return fail(ctx, B200_EINVAL, "null terms");
    const int N = ctx->N, F = ctx->F;
    if ((F > 0 && (!t->gamma_f || !t->magSf || !t->deltaCoeffs || !upper_out)) ||
        (N > 0 && (!t->V || !diag_out || !source_out)))
        return fail(ctx, B200_EINVAL, "null argument");
    if (t->psi && (!t->psi0 || !t->p0)) return fail(ctx, B200_EINVAL, "ddt term needs psi, psi0 and p0");
    if (t->nExplicit < 0 || t->nExplicit > kMaxExplicit || (t->nExplicit > 0 && !t->explicitFields))
        return fail(ctx, B200_EINVAL, "0 <= nExplicit <= 8");
    const bool needB = t->bPhi || t->bInternal || t->bBoundary;
    if (needB && (!ctx->bfStart || t->nB != ctx->nB))
        return fail(ctx, B200_ESTATE, "boundary arrays given but b200_set_boundary_faces was not called with the same nB");
    CU(cudaSetDevice(ctx->device));
    DevPlan& P = ctx->plans[0];
    PrghDev d;
    std::memset(&d, 0, sizeof(d));
    d.rDeltaT = t->rDeltaT; d.divSign = t->divSign;
    d.V = t->V; d.psi = t->psi; d.psi0 = t->psi0; d.p0 = t->p0; d.phi = t->phi; d.Su = t->Su;
    d.bPhi = t->bPhi; d.bInt = t->bInternal; d.bBou = t->bBoundary;
    d.nExplicit = t->nExplicit;
    for (int k = 0; k < t->nExplicit; ++k) {
        if (!t->explicitFields[k]) return fail(ctx, B200_EINVAL, "null explicit field");
        d.ex[k] = t->explicitFields[k];
    }
    d.bfStart = needB ? ctx->bfStart : nullptr;
    d.bfOrder = needB ? ctx->bfOrder : nullptr;
    LAUNCH(PC_ASM_FACE, k_face_coeff, grid_for(ctx, F, 16), F, t->gamma_f, t->magSf, t->deltaCoeffs,
           t->lapSign < 0 ? -1.0 : 1.0, upper_out);
    LAUNCH(PC_ASM_PRGH, k_prgh_cell, grid_for(ctx, N, 16), N, P.sliceBase, P.rowLen, P.faceOf, P.perm, ctx->d_l,
           upper_out, d, diag_out, source_out);
    CU(cudaStreamSynchronize(ctx->sc));
    CU(cudaGetLastError());
    prof_collect(ctx);
    return B200_OK;
}

int b200_assemble_p_rgh(b200_ctx* ctx, const b200_prgh_terms* t, double* upper_out, double* diag_out,
                        double* source_out) {
    if (!ctx) return B200_EINVAL;
    if (!ctx->haveMesh) return fail(ctx, B200_ESTATE, "assemble before set_addressing");
    if (!t) return fail(ctx, B200_EINVAL, "null terms");
    const size_t N = (size_t)ctx->N, F = (size_t)ctx->F;
    if (t->nExplicit < 0 || t->nExplicit > kMaxExplicit || (t->nExplicit > 0 && !t->explicitFields))
        return fail(ctx, B200_EINVAL, "0 <= nExplicit <= 8");
    if ((F > 0 && (!t->gamma_f || !t->magSf || !t->deltaCoeffs || !upper_out)) || (N > 0 && (!t->V || !diag_out || !source_out)))
        return fail(ctx, B200_EINVAL, "null argument");
    CU(cudaSetDevice(ctx->device));
    const bool needB = t->bPhi || t->bInternal || t->bBoundary;
    if (needB) {
        if (!t->bCells && t->nB > 0) return fail(ctx, B200_EINVAL, "boundary arrays given without bCells");
        RET(b200_set_boundary_faces(ctx, t->nB, t->bCells));
    }
    const size_t nB = needB ? (size_t)t->nB : 0;
    // staging arena: every array is padded to a multiple of 32 doubles
    auto pad = [](size_t n) { return (n + 31) & ~(size_t)31; };
    const size_t need = (7 + (size_t)t->nExplicit) * pad(N) + 5 * pad(F) + 3 * pad(nB) + 64;
    if (need > ctx->scratchElems) {
        CU(cudaStreamSynchronize(ctx->sc));
        dev_free(ctx->scratch);
        RET(dev_alloc(ctx, &ctx->scratch, need));
        ctx->scratchElems = need;
    }
    double* cur = ctx->scratch;
    cudaError_t cpErr = cudaSuccess;
    auto stage = [&](const double* h, size_t n) -> const double* {
        if (!h) return nullptr;
        double* d = cur;
        cur += pad(n);
        if (n && cpErr == cudaSuccess)
            cpErr = cudaMemcpyAsync(d, h, n * sizeof(double), cudaMemcpyHostToDevice, ctx->sc);
        return d;
    };
    b200_prgh_terms dt = *t;
    const double* exs[kMaxExplicit] = {};
    dt.V = stage(t->V, N); dt.psi = stage(t->psi, N); dt.psi0 = stage(t->psi0, N); dt.p0 = stage(t->p0, N);
    for (int k = 0; k < t->nExplicit; ++k) exs[k] = stage(t->explicitFields[k], N);
    dt.explicitFields = exs;
    dt.phi = stage(t->phi, F);
    dt.gamma_f = stage(t->gamma_f, F); dt.magSf = stage(t->magSf, F); dt.deltaCoeffs = stage(t->deltaCoeffs, F);
    dt.Su = stage(t->Su, N);
    dt.bPhi = stage(t->bPhi, nB); dt.bInternal = stage(t->bInternal, nB); dt.bBoundary = stage(t->bBoundary, nB);
    double* d_upper = cur; cur += pad(F);
    double* d_diag = cur; cur += pad(N);
    double* d_src = cur; cur += pad(N);
    if (cpErr != cudaSuccess) return fail(ctx, B200_ECUDA, std::string("H2D copy: ") + cudaGetErrorString(cpErr));
    RET(b200_assemble_p_rgh_device(ctx, &dt, d_upper, d_diag, d_src));
    CU(cudaMemcpyAsync(upper_out, d_upper, F * sizeof(double), cudaMemcpyDeviceToHost, ctx->sc));
    CU(cudaMemcpyAsync(diag_out, d_diag, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->sc));
    CU(cudaMemcpyAsync(source_out, d_src, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->sc));
    CU(cudaStreamSynchronize(ctx->sc));
    return B200_OK;
}

int b200_solve_device(b200_ctx* ctx, const double* d_diag, const double* d_upper,
                      const double* const* d_bou, const double* d_source, double* d_psi,
                      const b200_controls* ctl, b200_perf* perf) {
    if (!ctx) return B200_EINVAL;
    if (!ctx->haveMesh) return fail(ctx, B200_ESTATE, "solve before set_addressing");
    if (!ctl) return fail(ctx, B200_EINVAL, "null controls");
    if ((ctx->N > 0 && (!d_diag || !d_source || !d_psi)) || (ctx->F > 0 && !d_upper))
        return fail(ctx, B200_EINVAL, "null matrix/vector argument");
    CU(cudaSetDevice(ctx->device));
    RET(copy_bou(ctx, ctx->plans[0], d_bou, cudaMemcpyDeviceToDevice, ctx->bou));
    return solve_core(ctx, d_diag, d_upper, d_source, d_psi, ctl, perf);
}

int b200_solve(b200_ctx* ctx, const double* diag, const double* upper, const double* const* bou,
               const double* source, double* psi, const b200_controls* ctl, b200_perf* perf) {
    if (!ctx) return B200_EINVAL;
    if (!ctx->haveMesh) return fail(ctx, B200_ESTATE, "solve before set_addressing");
    if (!ctl) return fail(ctx, B200_EINVAL, "null controls");
    if ((ctx->N > 0 && (!diag || !source || !psi)) || (ctx->F > 0 && !upper))
        return fail(ctx, B200_EINVAL, "null matrix/vector argument");
    CU(cudaSetDevice(ctx->device));
    RET(ensure_staging(ctx, false));
    const size_t fb = (size_t)ctx->F * sizeof(double), nb = (size_t)ctx->N * sizeof(double);
    CU(cudaEventRecord(ctx->ev[3], ctx->sc));
    RET(h2d(ctx, ctx->in_upper, upper, fb));
    RET(h2d(ctx, ctx->in_diag, diag, nb));
    RET(h2d(ctx, ctx->in_src, source, nb));
    RET(h2d(ctx, ctx->in_psi, psi, nb));
    RET(copy_bou(ctx, ctx->plans[0], bou, cudaMemcpyHostToDevice, ctx->bou));
    CU(cudaEventRecord(ctx->ev[4], ctx->sc));
    int rc = solve_core(ctx, ctx->in_diag, ctx->in_upper, ctx->in_src, ctx->in_psi, ctl, perf);
    if (rc != B200_OK && rc != B200_ENONFINITE) return rc;
    CU(cudaEventRecord(ctx->ev[5], ctx->sc));
    RET(d2h(ctx, psi, ctx->in_psi, nb));
    CU(cudaEventRecord(ctx->ev[0], ctx->sc));
    CU(cudaStreamSynchronize(ctx->sc));
    if (perf) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]);
        perf->h2dMs = ms;
        cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[0]);
        perf->d2hMs = ms;
    }
    return rc;
}

int b200_amul(b200_ctx* ctx, const double* diag, const double* upper, const double* const* bou,
              const double* psi, double* Apsi) {
    if (!ctx) return B200_EINVAL;
    if (!ctx->haveMesh) return fail(ctx, B200_ESTATE, "amul before set_addressing");
    if ((ctx->N > 0 && (!diag || !psi || !Apsi)) || (ctx->F > 0 && !upper))
        return fail(ctx, B200_EINVAL, "null argument");
    CU(cudaSetDevice(ctx->device));
    RET(ensure_staging(ctx, false));
    DevPlan& P = ctx->plans[0];
    const size_t fb = (size_t)ctx->F * sizeof(double), nb = (size_t)ctx->N * sizeof(double);
    CU(cudaMemcpyAsync(ctx->in_upper, upper, fb, cudaMemcpyHostToDevice, ctx->sc));
    CU(cudaMemcpyAsync(ctx->in_diag, diag, nb, cudaMemcpyHostToDevice, ctx->sc));
    CU(cudaMemcpyAsync(ctx->in_psi, psi, nb, cudaMemcpyHostToDevice, ctx->sc));
    RET(copy_bou(ctx, P, bou, cudaMemcpyHostToDevice, ctx->bou));
    RET(reset_scalars(ctx, nullptr));
    RET(load_system(ctx, P, ctx->in_diag, ctx->in_upper, nullptr, ctx->in_psi));
    RET((spmv_full<false, false>(ctx, P, ctx->psi, ctx->w, nullptr, STEP_NONE)));
    const double* out = ctx->w;
    if (P.perm) {   // renumbered plan: back to the caller's cell order
        LAUNCH(PC_GATHER, k_scatter, grid_for(ctx, ctx->N), ctx->N, P.perm, ctx->w, ctx->in_src);
        out = ctx->in_src;
    }
    CU(cudaMemcpyAsync(Apsi, out, nb, cudaMemcpyDeviceToHost, ctx->sc));
    CU(cudaStreamSynchronize(ctx->sc));
    CU(cudaGetLastError());
    prof_collect(ctx);
    return B200_OK;
}

int b200_smooth_solve_device(b200_ctx* ctx, const double* d_diag, const double* d_upper, const double* d_lower,
                             const double* const* d_bou, const double* d_source, double* d_psi,
                             const b200_smooth_controls* ctl, b200_perf* perf) {
    if (!ctx) return B200_EINVAL;
    if (!ctx->haveMesh) return fail(ctx, B200_ESTATE, "smooth_solve before set_addressing");
    if (!ctl) return fail(ctx, B200_EINVAL, "null controls");
    if ((ctx->N > 0 && (!d_diag || !d_source || !d_psi)) || (ctx->F > 0 && !d_upper))
        return fail(ctx, B200_EINVAL, "null matrix/vector argument");
    CU(cudaSetDevice(ctx->device));
    RET(copy_bou(ctx, ctx->plans[0], d_bou, cudaMemcpyDeviceToDevice, ctx->bou));
    return smooth_core(ctx, d_diag, d_upper, d_lower, d_source, d_psi, ctl, perf);
}

int b200_smooth_solve(b200_ctx* ctx, const double* diag, const double* upper, const double* lower,
                      const double* const* bou, const double* source, double* psi,
                      const b200_smooth_controls* ctl, b200_perf* perf) {
    if (!ctx) return B200_EINVAL;
    if (!ctx->haveMesh) return fail(ctx, B200_ESTATE, "smooth_solve before set_addressing");
    if (!ctl) return fail(ctx, B200_EINVAL, "null controls");
    if ((ctx->N > 0 && (!diag || !source || !psi)) || (ctx->F > 0 && !upper))
        return fail(ctx, B200_EINVAL, "null matrix/vector argument");
    CU(cudaSetDevice(ctx->device));
    RET(ensure_staging(ctx, false));
    const size_t fb = (size_t)ctx->F * sizeof(double), nb = (size_t)ctx->N * sizeof(double);
    const bool asym = lower && lower != upper;
    if (asym && !ctx->in_lower) RET(dev_alloc(ctx, &ctx->in_lower, (size_t)ctx->F));
    CU(cudaEventRecord(ctx->ev[3], ctx->sc));
    RET(h2d(ctx, ctx->in_upper, upper, fb));
    if (asym) RET(h2d(ctx, ctx->in_lower, lower, fb));
    RET(h2d(ctx, ctx->in_diag, diag, nb));
    RET(h2d(ctx, ctx->in_src, source, nb));
    RET(h2d(ctx, ctx->in_psi, psi, nb));
    RET(copy_bou(ctx, ctx->plans[0], bou, cudaMemcpyHostToDevice, ctx->bou));
    CU(cudaEventRecord(ctx->ev[4], ctx->sc));
    int rc = smooth_core(ctx, ctx->in_diag, ctx->in_upper, asym ? ctx->in_lower : nullptr, ctx->in_src, ctx->in_psi,
                         ctl, perf);
    if (rc != B200_OK && rc != B200_ENONFINITE) return rc;
    CU(cudaEventRecord(ctx->ev[5], ctx->sc));
    RET(d2h(ctx, psi, ctx->in_psi, nb));
    CU(cudaEventRecord(ctx->ev[0], ctx->sc));
    CU(cudaStreamSynchronize(ctx->sc));
    if (perf) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]);
        perf->h2dMs = ms;
        cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[0]);
        perf->d2hMs = ms;
    }
    return rc;
}

int b200_bicg_solve_device(b200_ctx* ctx, const double* d_diag, const double* d_upper, const double* d_lower,
                           const double* const* d_bou, const double* const* d_int, const double* d_source,
                           double* d_psi, const b200_controls* ctl, b200_perf* perf) {
    (void)d_bou; (void)d_int;   // processor patches: not built (bicg_core rejects nranks > 1)
    if (!ctx) return B200_EINVAL;
    if (!ctx->haveMesh) return fail(ctx, B200_ESTATE, "bicg_solve before set_addressing");
    if (!ctl) return fail(ctx, B200_EINVAL, "null controls");
    if ((ctx->N > 0 && (!d_diag || !d_source || !d_psi)) || (ctx->F > 0 && !d_upper))
        return fail(ctx, B200_EINVAL, "null matrix/vector argument");
    CU(cudaSetDevice(ctx->device));
    return bicg_core(ctx, d_diag, d_upper, d_lower, d_source, d_psi, ctl, perf);
}

int b200_bicg_solve(b200_ctx* ctx, const double* diag, const double* upper, const double* lower,
                    const double* const* bou, const double* const* intc, const double* source, double* psi,
                    const b200_controls* ctl, b200_perf* perf) {
    (void)bou; (void)intc;
    if (!ctx) return B200_EINVAL;
    if (!ctx->haveMesh) return fail(ctx, B200_ESTATE, "bicg_solve before set_addressing");
    if (!ctl) return fail(ctx, B200_EINVAL, "null controls");
    if ((ctx->N > 0 && (!diag || !source || !psi)) || (ctx->F > 0 && !upper))
        return fail(ctx, B200_EINVAL, "null matrix/vector argument");
    CU(cudaSetDevice(ctx->device));
    RET(ensure_staging(ctx, false));
    const size_t fb = (size_t)ctx->F * sizeof(double), nb = (size_t)ctx->N * sizeof(double);
    const bool asym = lower && lower != upper;
    if (asym && !ctx->in_lower) RET(dev_alloc(ctx, &ctx->in_lower, (size_t)ctx->F));
    CU(cudaEventRecord(ctx->ev[3], ctx->sc));
    RET(h2d(ctx, ctx->in_upper, upper, fb));
    if (asym) RET(h2d(ctx, ctx->in_lower, lower, fb));
    RET(h2d(ctx, ctx->in_diag, diag, nb));
    RET(h2d(ctx, ctx->in_src, source, nb));
    RET(h2d(ctx, ctx->in_psi, psi, nb));
    CU(cudaEventRecord(ctx->ev[4], ctx->sc));
    int rc = bicg_core(ctx, ctx->in_diag, ctx->in_upper, asym ? ctx->in_lower : nullptr, ctx->in_src, ctx->in_psi, ctl, perf);
    if (rc != B200_OK && rc != B200_ENONFINITE) return rc;
    CU(cudaEventRecord(ctx->ev[5], ctx->sc));
    RET(d2h(ctx, psi, ctx->in_psi, nb));
    CU(cudaEventRecord(ctx->ev[0], ctx->sc));
    CU(cudaStreamSynchronize(ctx->sc));
    if (perf) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]);
        perf->h2dMs = ms;
        cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[0]);
        perf->d2hMs = ms;
    }
    return rc;
}

int b200_amul_asym(b200_ctx* ctx, const double* diag, const double* upper, const double* lower,
                   const double* const* bou, const double* psi, double* Apsi) {
    if (!ctx) return B200_EINVAL;
    if (!ctx->haveMesh) return fail(ctx, B200_ESTATE, "amul before set_addressing");
    if ((ctx->N > 0 && (!diag || !psi || !Apsi)) || (ctx->F > 0 && !upper))
        return fail(ctx, B200_EINVAL, "null argument");
    CU(cudaSetDevice(ctx->device));
    RET(ensure_staging(ctx, false));
    DevPlan& P = ctx->plans[0];
    const size_t fb = (size_t)ctx->F * sizeof(double), nb = (size_t)ctx->N * sizeof(double);
    const bool asym = lower && lower != upper;
    if (asym && !ctx->in_lower) RET(dev_alloc(ctx, &ctx->in_lower, (size_t)ctx->F));
    CU(cudaMemcpyAsync(ctx->in_upper, upper, fb, cudaMemcpyHostToDevice, ctx->sc));
    if (asym) CU(cudaMemcpyAsync(ctx->in_lower, lower, fb, cudaMemcpyHostToDevice, ctx->sc));
    CU(cudaMemcpyAsync(ctx->in_diag, diag, nb, cudaMemcpyHostToDevice, ctx->sc));
    CU(cudaMemcpyAsync(ctx->in_psi, psi, nb, cudaMemcpyHostToDevice, ctx->sc));
    RET(copy_bou(ctx, P, bou, cudaMemcpyHostToDevice, ctx->bou));
    RET(reset_scalars(ctx, nullptr));
    RET(load_system_asym(ctx, P, ctx->in_diag, ctx->in_upper, asym ? ctx->in_lower : nullptr, nullptr, ctx->in_psi));
    ctx->forceEll = true;       // the single-read layouts assume lower == upper
    const int rcA = spmv_full<false, false>(ctx, P, ctx->psi, ctx->w, nullptr, STEP_NONE);
    ctx->forceEll = false;
    RET(rcA);
    const double* out = ctx->w;
    if (P.perm) {   // renumbered plan: back to the caller's cell order
        LAUNCH(PC_GATHER, k_scatter, grid_for(ctx, ctx->N), ctx->N, P.perm, ctx->w, ctx->in_src);
        out = ctx->in_src;
    }
    CU(cudaMemcpyAsync(Apsi, out, nb, cudaMemcpyDeviceToHost, ctx->sc));
    CU(cudaStreamSynchronize(ctx->sc));
    CU(cudaGetLastError());
    prof_collect(ctx);
    return B200_OK;
}

int b200_flux(b200_ctx* ctx, const double* upper, const double* psi, double* flux_out) {
    if (!ctx) return B200_EINVAL;
    if (!ctx->haveMesh) return fail(ctx, B200_ESTATE, "flux before set_addressing");
    if ((ctx->F > 0 && (!upper || !flux_out)) || (ctx->N > 0 && !psi))
        return fail(ctx, B200_EINVAL, "null argument");
    CU(cudaSetDevice(ctx->device));
    RET(ensure_staging(ctx, true));
    const size_t fb = (size_t)ctx->F * sizeof(double), nb = (size_t)ctx->N * sizeof(double);
    CU(cudaMemcpyAsync(ctx->in_upper, upper, fb, cudaMemcpyHostToDevice, ctx->sc));
    CU(cudaMemcpyAsync(ctx->in_psi, psi, nb, cudaMemcpyHostToDevice, ctx->sc));
    LAUNCH(PC_FLUX, k_flux, grid_for(ctx, ctx->F, 16), ctx->F, ctx->d_l, ctx->d_u, ctx->in_upper,
           ctx->in_psi, ctx->in_f1);
    CU(cudaMemcpyAsync(flux_out, ctx->in_f1, fb, cudaMemcpyDeviceToHost, ctx->sc));
    CU(cudaStreamSynchronize(ctx->sc));
    CU(cudaGetLastError());
    prof_collect(ctx);
    return B200_OK;
}

int b200_host_alloc(void** p, size_t bytes) {
    b200_ctx* ctx = nullptr;
    if (!p) return fail(ctx, B200_EINVAL, "null pointer");
    *p = nullptr;
    if (b200_device_count() <= 0) return fail(ctx, B200_ENODEVICE, "no CUDA device");
    CU(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault));
    return B200_OK;
}
void b200_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

uint64_t b200_launch_count(const b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

int b200_debug_force_iterations(b200_ctx* ctx, int32_t n) {
    if (!ctx || n < 0) return B200_EINVAL;
    ctx->forceIters = n;
    return B200_OK;
}

int b200_profile_enable(b200_ctx* ctx, int on) {
    if (!ctx) return B200_EINVAL;
    CU(cudaSetDevice(ctx->device));
    if (on && ctx->profPool.empty()) {
        ctx->profPool.resize(2 * 16384);
        for (auto& ev : ctx->profPool) CU(cudaEventCreate(&ev));
    }
    ctx->prof = on != 0;
    if (on) {
        std::memset(ctx->profMs, 0, sizeof(ctx->profMs));
        std::memset(ctx->profN, 0, sizeof(ctx->profN));
        ctx->profSkipped = 0;
        ctx->profRecs.clear();
        ctx->profUsed = 0;
    }
    return B200_OK;
}

const char* b200_profile_json(b200_ctx* ctx) {
    if (!ctx) return "{}";
    std::string s = "{";
    bool first = true;
    char buf[640];
    for (int i = 0; i < PC_COUNT; ++i) {
        if (!ctx->profN[i]) continue;
        std::snprintf(buf, sizeof(buf), "%s\"%s\": {\"launches\": %llu, \"total_ms\": %.6f, \"avg_us\": %.3f}",
                      first ? "" : ", ", kProfNames[i], (unsigned long long)ctx->profN[i],
                      ctx->profMs[i], 1e3 * ctx->profMs[i] / (double)ctx->profN[i]);
        s += buf;
        first = false;
    }
    if (ctx->nranks > 1 && ctx->hS) {
        // wait accounting of the LAST solve (kernels.cuh Scalars): per-wait averages in us
        const Scalars& h = *ctx->hS;
        std::snprintf(buf, sizeof(buf), "%s\"_wait_halo_flags\": {\"launches\": %u, \"total_ms\": %.6f, \"avg_us\": %.3f}, "
                      "\"_wait_peer_reduction\": {\"launches\": %u, \"total_ms\": %.6f, \"avg_us\": %.3f}",
                      first ? "" : ", ", h.nHaloWaits, h.waitHaloNs * 1e-6, h.nHaloWaits ? h.waitHaloNs * 1e-3 / h.nHaloWaits : 0.0,
                      h.nRedWaits, h.waitRedNs * 1e-6, h.nRedWaits ? h.waitRedNs * 1e-3 / h.nRedWaits : 0.0);
        s += buf;
        first = false;
    }
    if (ctx->profSkipped) {
        // launches of surplus loop bodies (enqueued in batches, returned at once on S->done): not in the averages
        std::snprintf(buf, sizeof(buf), "%s\"_skipped_noop_launches\": {\"launches\": %llu, \"total_ms\": 0.0, \"avg_us\": 0.0}",
                      first ? "" : ", ", (unsigned long long)ctx->profSkipped);
        s += buf;
    }
    s += "}";
    ctx->profJson = s;
    return ctx->profJson.c_str();
}

const char* b200_describe(b200_ctx* ctx) {
    if (!ctx) return "{}";
    char buf[2048];
    const DevPlan& P = ctx->plans[0];
    const char* amul = !P.built ? "none"
                       : (P.sym && P.symTma) ? "k_spmv_sym_tma<DOT,STAGES>"
                       : P.sr ? "k_spmv_sr<INIT,DOT>"
                       : P.sym ? "k_spmv_sym<INIT,DOT>" : "k_spmv<INIT,DOT>";
    std::snprintf(buf, sizeof(buf),
                  "{\"amul_natural\": \"%s\", \"amul_permuted\": \"k_spmv<INIT,DOT>\", \"symWU\": %d, \"symWL\": %d, "
                  "\"chunk_rows\": %d, \"tma_stages\": %d, "
                  "\"nranks\": %d, \"peer_allreduce\": %s, \"nCells\": %d, \"nFaces\": %d, \"nSlots\": %d, \"sms\": %d, "
                  "\"renumbered_rcm\": %s, \"mean_face_span_natural\": %.1f, \"mean_face_span_used\": %.1f, "
                  "\"sectors_per_gather_natural\": %.2f, \"sectors_per_gather_used\": %.2f, "
                  "\"small_system_cluster_kernel\": %s, \"small_on_chip\": %s, \"small_n_max\": %d, "
                  "\"multicolour_tiles\": %d, \"multicolour_tile_rows\": %d, \"multicolour_amul\": \"%s\", "
                  "\"ell_col16_fraction_natural\": %.3f, \"ell_col16_fraction_multicolour\": %.3f, "
                  "\"iteration_graphs\": %s, \"graph_iterations\": %d, \"graph_launches\": %llu, "
                  "\"halo_exchange\": \"%s\"}",
                  amul, P.symWU, P.symWL, kChunkRows, ctx->symStages, ctx->nranks,
                  ctx->p2pReduce ? "true" : "false", ctx->N, ctx->F, ctx->nSlots, ctx->numSMs,
                  P.h.renumbered ? "true" : "false", P.h.spanNatural, P.h.spanUsed, P.h.sectorsNatural,
                  P.h.sectorsUsed, ctx->usedSmall ? "true" : "false", ctx->usedFast ? "true" : "false", ctx->smallN,
                  ctx->plans[1].built ? ctx->plans[1].h.nTiles : 0, ctx->plans[1].built ? ctx->plans[1].h.tileRows : 0,
                  !ctx->plans[1].built ? "n/a" : (ctx->plans[1].sym ? (ctx->plans[1].symTma ? "k_spmv_sym_tma" : "k_spmv_sym") : "k_spmv"),
                  P.c16 ? P.h.col16Fraction : 0.0, ctx->plans[1].c16 ? ctx->plans[1].h.col16Fraction : 0.0,
                  ctx->useGraph ? "true" : "false", kGraphIters, (unsigned long long)ctx->graphLaunches,
                  ctx->nranks == 1 ? "none" : (ctx->p2pHalo ? "peer-memory stores (k_pack_p2p)" : "ncclSend/ncclRecv"));
    ctx->profJson = buf;
    return ctx->profJson.c_str();
}

}  // extern "C"
