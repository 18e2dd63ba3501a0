// plan_debug.cpp -- host-only inspection ABI for the row structure built by plan.cpp, so that
// CPU-only tests can validate ordering, colouring, the sliced-ELL layout and the interface CSR
// without a GPU.  Not part of the OpenFOAM-facing contract (not declared in include/b200pcg.h).
#include "plan.hpp"

#include <cstdlib>
#include <cstring>
#include <string>

using namespace b200;

namespace {
thread_local std::string g_err;
}

extern "C" {

struct b200_dbg_iface {
    int32_t nbrRank;
    int32_t nFaces;
    const int32_t* faceCells;
};

void* b200_debug_plan_build(int ordering, int32_t N, int32_t F, const int32_t* l, const int32_t* u,
                            int32_t nIfaces, const b200_dbg_iface* ifaces) {
    auto* P = new HostPlan();
    static_assert(sizeof(b200_dbg_iface) == sizeof(IfaceIn), "layout");
    g_err = build_plan((Ordering)ordering, N, F, l, u, nIfaces, (const IfaceIn*)ifaces, *P);
    if (!g_err.empty()) {
        delete P;
        return nullptr;
    }
    return P;
}
void* b200_debug_plan_build2(int ordering, int renumber, int32_t N, int32_t F, const int32_t* l,
                             const int32_t* u, int32_t nIfaces, const b200_dbg_iface* ifaces, int32_t tileRows) {
    auto* P = new HostPlan();
    const char* sc = std::getenv("B200PCG_SORT_COLS");   // same switch as the context (solver.cu)
    g_err = build_plan((Ordering)ordering, N, F, l, u, nIfaces, (const IfaceIn*)ifaces, *P, (Renumber)renumber,
                       tileRows, sc && std::atoi(sc) != 0);
    if (!g_err.empty()) {
        delete P;
        return nullptr;
    }
    return P;
}
int32_t b200_debug_plan_renumbered(void* h) { return ((HostPlan*)h)->renumbered ? 1 : 0; }
double b200_debug_plan_span(void* h, int which) {
    return which ? ((HostPlan*)h)->spanUsed : ((HostPlan*)h)->spanNatural;
}
const char* b200_debug_plan_error(void) { return g_err.c_str(); }
void b200_debug_plan_free(void* h) { delete (HostPlan*)h; }

// returns element count, sets *ptr and *elemBytes; -1 for an unknown name
int64_t b200_debug_plan_get(void* h, const char* name, const void** ptr, int32_t* elemBytes) {
    HostPlan& P = *(HostPlan*)h;
#define V(nm, vec)                                  \
    if (!std::strcmp(name, nm)) {                   \
        *ptr = (vec).data();                        \
        *elemBytes = (int32_t)sizeof((vec)[0]);     \
        return (int64_t)(vec).size();               \
    }
    V("perm", P.perm) V("iperm", P.iperm) V("colourStart", P.colourStart)
    V("sliceBase", P.sliceBase) V("rowLen", P.rowLen) V("col", P.col) V("faceOf", P.faceOf)
    V("nbrRank", P.nbrRank) V("patchStart", P.patchStart) V("slotRow", P.slotRow)
    V("bRow", P.bRow) V("bStart", P.bStart) V("bSlot", P.bSlot)
    V("segStart", P.segStart) V("rowColour", P.rowColour) V("sym.rowLen", P.sym.rowLen)
    V("colBase", P.colBase) V("col16", P.col16)
    V("sym.uCol", P.sym.uCol)
    V("sym.uFace", P.sym.uFace) V("sym.lRef", P.sym.lRef)
    V("sr.meta", P.sr.meta) V("sr.ownBase", P.sr.ownBase) V("sr.ownFace", P.sr.ownFace)
#undef V
    return -1;
}
int32_t b200_debug_plan_sr_valid(void* h) { return ((HostPlan*)h)->sr.valid ? 1 : 0; }
int32_t b200_debug_plan_sym_valid(void* h) { return ((HostPlan*)h)->sym.valid ? 1 : 0; }
int32_t b200_debug_plan_sym_wu(void* h) { return ((HostPlan*)h)->sym.WU; }
int32_t b200_debug_plan_sym_wl(void* h) { return ((HostPlan*)h)->sym.WL; }
double b200_debug_plan_col16_fraction(void* h) { return ((HostPlan*)h)->col16Fraction; }
int32_t b200_debug_plan_ntiles(void* h) { return ((HostPlan*)h)->nTiles; }
int32_t b200_debug_plan_ncolours(void* h) { return ((HostPlan*)h)->nColours; }
int64_t b200_debug_plan_nentries(void* h) { return ((HostPlan*)h)->nEntries; }
}
