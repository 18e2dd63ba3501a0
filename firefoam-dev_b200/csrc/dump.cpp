// dump.cpp -- matrix dump / replay format, host-only.
//
// The reference ships no per-time-step linear systems (only a solver log), and its own
// implementation of the path cannot be built here, so real steckler / singleBox p_rgh systems can
// only come from a machine that has OpenFOAM: the adapter (adapter/B200PCG.C) writes what
// lduMatrix::solver::solve receives -- lduAddressing, diag, upper, per-interface faceCells +
// interfaceBouCoeffs + neighbour rank, totalSource, the initial psi, the solver controls -- and,
// after the solve, the SolverPerformance it reported, one file per rank and solve.  The same file
// is the regression corpus of this repository (tests/golden/*.b200sys) and the input of
// tools/b200replay.
//
// File layout (little-endian):
//     0   char[8]   magic "B200LDU\1"
//     8   uint64    headerBytes
//     16  char[headerBytes]  JSON header (UTF-8), then zero padding to a multiple of 64
//     ... raw arrays, each starting at a multiple of 64 bytes from the start of the file
// The header lists every array: {"name", "dtype" ("i4" | "f8"), "count", "offset"}.
#include "../../include/b200pcg.h"

#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_dumpError;
const char kMagic[8] = {'B', '2', '0', '0', 'L', 'D', 'U', '\1'};

int dfail(const std::string& m) {
    g_dumpError = m;
    return B200_EINVAL;
}

struct ArrayRec {
    std::string name, dtype;
    uint64_t count, offset;
    const void* data;
};

uint64_t align64(uint64_t x) { return (x + 63u) & ~(uint64_t)63u; }

std::string json_escape(const char* s) {
    std::string o;
    for (; s && *s; ++s) {
        if (*s == '"' || *s == '\\') { o += '\\'; o += *s; }
        else if ((unsigned char)*s < 0x20) o += ' ';
        else o += *s;
    }
    return o;
}

std::string fmt_double(double v) {
    char b[64];
    std::snprintf(b, sizeof(b), "%.17g", v);
    // JSON has no inf/nan
    if (std::strstr(b, "inf") || std::strstr(b, "nan")) return "null";
    return b;
}

const char* precond_name(int p) {
    switch (p) {
        case B200_PRECOND_NONE: return "none";
        case B200_PRECOND_DIAGONAL: return "diagonal";
        case B200_PRECOND_DIC_MC: return "DIC";
        case B200_PRECOND_DIC_EXACT: return "DIC";
        case B200_PRECOND_DIC_MC_EIS: return "DIC";
        default: return "unknown";
    }
}

// ---- minimal JSON field extraction for the reader (the writer above is the only producer of the
// header besides the Python twin in firefoam-dev_b200/replay.py, which emits the same flat shape)
// A key is a string token FOLLOWED BY ':' -- string VALUES that spell a key name (a field called "nCells",
// a solver called "reference") are skipped, because the scan walks the header token by token.
bool find_key(const std::string& js, const std::string& key, size_t from, size_t& valuePos) {
    size_t p = from;
    while (p < js.size()) {
        if (js[p] != '"') { ++p; continue; }
        const size_t b = ++p;                       // string token [b, e)
        while (p < js.size() && js[p] != '"') p += (js[p] == '\\' && p + 1 < js.size()) ? 2 : 1;
        const size_t e = p;
        if (p < js.size()) ++p;
        size_t q = p;
        while (q < js.size() && (js[q] == ' ' || js[q] == '\n' || js[q] == '\t' || js[q] == '\r')) ++q;
        if (q >= js.size() || js[q] != ':') continue;          // a value, not a key
        ++q;
        while (q < js.size() && (js[q] == ' ' || js[q] == '\n' || js[q] == '\t' || js[q] == '\r')) ++q;
        if (e - b == key.size() && js.compare(b, e - b, key) == 0) {
            valuePos = q;
            return true;
        }
        p = q;
    }
    return false;
}
// a size read from the header: finite, integral, 0 <= v <= limit
bool as_count(double v, uint64_t limit, uint64_t& out) {
    if (!(v >= 0.0) || v > (double)limit || v != (double)(uint64_t)v) return false;
    out = (uint64_t)v;
    return true;
}
bool get_number(const std::string& js, const std::string& key, size_t from, double& out) {
    size_t p;
    if (!find_key(js, key, from, p)) return false;
    if (js.compare(p, 4, "null") == 0) { out = 0.0; return true; }
    out = std::strtod(js.c_str() + p, nullptr);
    return true;
}
bool get_string(const std::string& js, const std::string& key, size_t from, std::string& out) {
    size_t p;
    if (!find_key(js, key, from, p) || p >= js.size() || js[p] != '"') return false;
    out.clear();
    for (++p; p < js.size() && js[p] != '"'; ++p) {
        if (js[p] == '\\' && p + 1 < js.size()) ++p;
        out += js[p];
    }
    return true;
}

}  // namespace

extern "C" {

const char* b200_dump_last_error(void) { return g_dumpError.c_str(); }

int b200_dump_write(const char* path, const b200_dump* d) {
    if (!path || !d) return dfail("null argument");
    if (d->nCells < 0 || d->nFaces < 0 || d->nIfaces < 0) return dfail("negative size");
    if ((d->nFaces > 0 && (!d->lowerAddr || !d->upperAddr || !d->upper)) ||
        (d->nCells > 0 && (!d->diag || !d->source || !d->psi0)))
        return dfail("null matrix/vector array");
    if (d->nIfaces > 0 && (!d->ifaces || !d->ifaceBouCoeffs)) return dfail("null interface arrays");
    std::vector<ArrayRec> arrs;
    auto add = [&](const std::string& name, const char* dtype, uint64_t count, const void* data) {
        arrs.push_back(ArrayRec{name, dtype, count, 0, data});
    };
    add("lowerAddr", "i4", (uint64_t)d->nFaces, d->lowerAddr);
    add("upperAddr", "i4", (uint64_t)d->nFaces, d->upperAddr);
    add("diag", "f8", (uint64_t)d->nCells, d->diag);
    add("upper", "f8", (uint64_t)d->nFaces, d->upper);
    add("source", "f8", (uint64_t)d->nCells, d->source);
    add("psi0", "f8", (uint64_t)d->nCells, d->psi0);
    if (d->psiSolution) add("psi", "f8", (uint64_t)d->nCells, d->psiSolution);
    if (d->lower && d->nFaces > 0) add("lower", "f8", (uint64_t)d->nFaces, d->lower);
    for (int k = 0; k < d->nIfaces; ++k) {
        if (d->ifaces[k].nFaces < 0 || (d->ifaces[k].nFaces > 0 && (!d->ifaces[k].faceCells || !d->ifaceBouCoeffs[k])))
            return dfail("bad interface " + std::to_string(k));
        add("iface" + std::to_string(k) + ".faceCells", "i4", (uint64_t)d->ifaces[k].nFaces, d->ifaces[k].faceCells);
        add("iface" + std::to_string(k) + ".bouCoeffs", "f8", (uint64_t)d->ifaces[k].nFaces, d->ifaceBouCoeffs[k]);
    }
    // the header length depends on the offsets and vice versa: lay out with a fixed-width offset field
    auto header = [&](bool final) {
        std::string h = "{\"format\": \"b200-ldu-system\", \"version\": 1, \"fieldName\": \"" +
                        json_escape(d->fieldName ? d->fieldName : "") + "\", \"rank\": " + std::to_string(d->rank) +
                        ", \"nranks\": " + std::to_string(d->nranks) + ", \"nCells\": " + std::to_string(d->nCells) +
                        ", \"nFaces\": " + std::to_string(d->nFaces) + ", \"symmetric\": " +
                        ((d->lower && d->nFaces > 0) ? "false" : "true") + ", \"solveIndex\": " +
                        std::to_string(d->solveIndex) + ", \"time\": " + fmt_double(d->time);
        if (d->haveSmooth) {
            // a smoothSolver solve: lduMatrix::solver's controls + nSweeps, smoother, sweep order
            const b200_smooth_controls& c = d->smooth;
            h += std::string(", \"controls\": {\"solver\": \"smoothSolver\", \"smoother\": \"") +
                 (c.smoother == B200_SMOOTHER_SYM_GAUSS_SEIDEL ? "symGaussSeidel" : "GaussSeidel") +
                 "\", \"smootherCode\": " + std::to_string(c.smoother) + ", \"sweepMode\": \"" +
                 (c.sweepMode == B200_SWEEP_EXACT ? "exact" : "multicolour") + "\", \"sweepModeCode\": " +
                 std::to_string(c.sweepMode) + ", \"nSweeps\": " + std::to_string(c.nSweeps) +
                 ", \"tolerance\": " + fmt_double(c.tolerance) + ", \"relTol\": " + fmt_double(c.relTol) +
                 ", \"maxIter\": " + std::to_string(c.maxIter) + ", \"minIter\": " + std::to_string(c.minIter) + "}";
        } else if (d->havePBiCG) {
            // a PBiCG solve: the asymmetric preconditioner names
            const int pc = d->controls.precond;
            h += std::string(", \"controls\": {\"solver\": \"PBiCG\", \"preconditioner\": \"") +
                 (pc == B200_PRECOND_NONE ? "none" : pc == B200_PRECOND_DIAGONAL ? "diagonal" : "DILU") +
                 "\", \"precondCode\": " + std::to_string(pc) +
                 ", \"tolerance\": " + fmt_double(d->controls.tolerance) + ", \"relTol\": " +
                 fmt_double(d->controls.relTol) + ", \"maxIter\": " + std::to_string(d->controls.maxIter) +
                 ", \"minIter\": " + std::to_string(d->controls.minIter) + "}";
        } else {
            h += std::string(", \"controls\": {\"preconditioner\": \"") + precond_name(d->controls.precond) +
                 "\", \"precondCode\": " + std::to_string(d->controls.precond) +
                 ", \"tolerance\": " + fmt_double(d->controls.tolerance) + ", \"relTol\": " +
                 fmt_double(d->controls.relTol) + ", \"maxIter\": " + std::to_string(d->controls.maxIter) +
                 ", \"minIter\": " + std::to_string(d->controls.minIter) + "}";
        }
        if (d->havePerf)
            h += ", \"reference\": {\"solverName\": \"" + json_escape(d->solverName ? d->solverName : "") +
                 "\", \"initialResidual\": " + fmt_double(d->perf.initialResidual) + ", \"finalResidual\": " +
                 fmt_double(d->perf.finalResidual) + ", \"nIterations\": " + std::to_string(d->perf.nIterations) +
                 ", \"converged\": " + std::to_string(d->perf.converged) + ", \"singular\": " +
                 std::to_string(d->perf.singular) + "}";
        h += ", \"interfaces\": [";
        for (int k = 0; k < d->nIfaces; ++k) {
            if (k) h += ", ";
            h += "{\"nbrRank\": " + std::to_string(d->ifaces[k].nbrRank) + ", \"nFaces\": " +
                 std::to_string(d->ifaces[k].nFaces) + ", \"tag\": " + std::to_string(d->ifaces[k].tag) + "}";
        }
        h += "], \"arrays\": [";
        for (size_t i = 0; i < arrs.size(); ++i) {
            char off[32];
            std::snprintf(off, sizeof(off), "%020" PRIu64, final ? arrs[i].offset : (uint64_t)0);
            if (i) h += ", ";
            // offsets are zero-padded decimal strings of fixed width (JSON numbers may not have leading zeros)
            h += "{\"name\": \"" + arrs[i].name + "\", \"dtype\": \"" + arrs[i].dtype + "\", \"count\": " +
                 std::to_string(arrs[i].count) + ", \"offset\": \"" + off + "\"}";
        }
        h += "]}";
        return h;
    };
    const uint64_t hlen = header(false).size();
    uint64_t pos = align64(16 + hlen);
    for (auto& a : arrs) {
        a.offset = pos;
        pos = align64(pos + a.count * (a.dtype == "i4" ? 4u : 8u));
    }
    const std::string h = header(true);
    if (h.size() != hlen) return dfail("internal: header length changed");
    FILE* f = std::fopen(path, "wb");
    if (!f) return dfail(std::string("cannot open for writing: ") + path);
    bool ok = std::fwrite(kMagic, 1, 8, f) == 8 && std::fwrite(&hlen, 8, 1, f) == 1 &&
              std::fwrite(h.data(), 1, h.size(), f) == h.size();
    uint64_t at = 16 + hlen;
    static const char zeros[64] = {0};
    for (auto& a : arrs) {
        if (!ok) break;
        ok = std::fwrite(zeros, 1, (size_t)(a.offset - at), f) == (size_t)(a.offset - at);
        const size_t bytes = (size_t)(a.count * (a.dtype == "i4" ? 4u : 8u));
        if (ok && bytes) ok = std::fwrite(a.data, 1, bytes, f) == bytes;
        at = a.offset + bytes;
    }
    if (ok) {
        const uint64_t end = align64(at);
        ok = std::fwrite(zeros, 1, (size_t)(end - at), f) == (size_t)(end - at);
    }
    if (std::fclose(f) != 0) ok = false;
    if (!ok) return dfail(std::string("write error: ") + path);
    return B200_OK;
}

struct b200_dump_file {
    std::string json, fieldName, solverName;
    std::vector<char> blob;          // whole file
    std::vector<b200_iface> ifaces;
    std::vector<const double*> bou;
    b200_dump d;
};

int b200_dump_read(const char* path, b200_dump_file** out) {
    if (!path || !out) return dfail("null argument");
    *out = nullptr;
    FILE* f = std::fopen(path, "rb");
    if (!f) return dfail(std::string("cannot open: ") + path);
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    auto* F = new b200_dump_file();
    F->blob.resize(sz > 0 ? (size_t)sz : 0);
    const bool rd = sz >= 16 && std::fread(F->blob.data(), 1, (size_t)sz, f) == (size_t)sz;
    std::fclose(f);
    auto bail = [&](const std::string& m) {
        delete F;
        return dfail(m);
    };
    if (!rd || std::memcmp(F->blob.data(), kMagic, 8) != 0) return bail(std::string("not a b200 system dump: ") + path);
    uint64_t hlen;
    std::memcpy(&hlen, F->blob.data() + 8, 8);
    if (hlen > (uint64_t)sz - 16) return bail("truncated header");   // (sz >= 16 checked above; no wrap-around)
    F->json.assign(F->blob.data() + 16, (size_t)hlen);
    const std::string& js = F->json;
    std::memset(&F->d, 0, sizeof(F->d));
    double v;
    if (!get_number(js, "version", 0, v) || (int)v != 1) return bail("unsupported dump version");
    auto num = [&](const char* key, size_t from, double& dst) { return get_number(js, key, from, dst); };
    double nC = 0, nF = 0, rk = 0, nr = 1, si = 0, tm = 0;
    if (!num("nCells", 0, nC) || !num("nFaces", 0, nF)) return bail("header lacks nCells/nFaces");
    num("rank", 0, rk); num("nranks", 0, nr); num("solveIndex", 0, si); num("time", 0, tm);
    get_string(js, "fieldName", 0, F->fieldName);
    F->d.fieldName = F->fieldName.c_str();
    uint64_t uC, uF, uRk, uNr, uSi;
    if (!as_count(nC, INT32_MAX, uC) || !as_count(nF, INT32_MAX, uF) || !as_count(rk, INT32_MAX, uRk) ||
        !as_count(nr, INT32_MAX, uNr) || uNr < 1 || uRk >= uNr || !as_count(si, INT32_MAX, uSi))
        return bail("header sizes must be non-negative integers (nCells, nFaces, rank < nranks, solveIndex)");
    F->d.nCells = (int32_t)uC; F->d.nFaces = (int32_t)uF; F->d.rank = (int32_t)uRk; F->d.nranks = (int32_t)uNr;
    F->d.solveIndex = (int32_t)uSi; F->d.time = tm;
    size_t cpos;
    if (find_key(js, "controls", 0, cpos)) {
        double t;
        if (num("precondCode", cpos, t)) F->d.controls.precond = (int32_t)t;
        if (num("tolerance", cpos, t)) F->d.controls.tolerance = t;
        if (num("relTol", cpos, t)) F->d.controls.relTol = t;
        if (num("maxIter", cpos, t)) F->d.controls.maxIter = (int32_t)t;
        if (num("minIter", cpos, t)) F->d.controls.minIter = (int32_t)t;
        // (the keys of the controls block come before "reference" / "interfaces" / "arrays": bound the search)
        size_t cend = js.find('}', cpos);
        size_t spos;
        std::string solverKind;
        if (find_key(js, "solver", cpos, spos) && spos < cend && get_string(js, "solver", cpos, solverKind) &&
            solverKind == "PBiCG")
            F->d.havePBiCG = 1;
        if (find_key(js, "smootherCode", cpos, spos) && spos < cend) {
            F->d.haveSmooth = 1;
            b200_smooth_controls& c = F->d.smooth;
            c.tolerance = F->d.controls.tolerance; c.relTol = F->d.controls.relTol;
            c.maxIter = F->d.controls.maxIter; c.minIter = F->d.controls.minIter;
            c.nSweeps = 1;
            if (num("smootherCode", cpos, t)) c.smoother = (int32_t)t;
            if (num("sweepModeCode", cpos, t)) c.sweepMode = (int32_t)t;
            if (num("nSweeps", cpos, t)) c.nSweeps = (int32_t)t;
            c.reserved = 0;
        }
    }
    size_t rpos;
    if (find_key(js, "reference", 0, rpos)) {
        double t;
        F->d.havePerf = 1;
        get_string(js, "solverName", rpos, F->solverName);
        F->d.solverName = F->solverName.c_str();
        if (num("initialResidual", rpos, t)) F->d.perf.initialResidual = t;
        if (num("finalResidual", rpos, t)) F->d.perf.finalResidual = t;
        if (num("nIterations", rpos, t)) F->d.perf.nIterations = (int32_t)t;
        if (num("converged", rpos, t)) F->d.perf.converged = (int32_t)t;
        if (num("singular", rpos, t)) F->d.perf.singular = (int32_t)t;
    }
    // interfaces
    size_t ipos;
    if (find_key(js, "interfaces", 0, ipos)) {
        const size_t iend = js.find(']', ipos);
        size_t p = ipos;
        while (true) {
            size_t q = js.find("\"nbrRank\"", p);
            if (q == std::string::npos || q > iend) break;
            double a = 0, b = 0, c = 0;
            num("nbrRank", q, a); num("nFaces", q, b); num("tag", q, c);
            uint64_t ua, ub;
            if (!as_count(a, INT32_MAX, ua) || !as_count(b, INT32_MAX, ub) || !(c >= INT32_MIN && c <= INT32_MAX))
                return bail("bad interface record in the header");
            b200_iface it;
            it.nbrRank = (int32_t)ua; it.nFaces = (int32_t)ub; it.faceCells = nullptr; it.tag = (int32_t)c;
            F->ifaces.push_back(it);
            p = q + 9;
        }
    }
    F->bou.assign(F->ifaces.size(), nullptr);
    // arrays
    size_t apos;
    if (!find_key(js, "arrays", 0, apos)) return bail("header lacks arrays");
    size_t p = apos;
    while (true) {
        size_t q = js.find("\"name\"", p);
        if (q == std::string::npos) break;
        std::string name, dtype, offs;
        double cnt = 0;
        get_string(js, "name", q, name); get_string(js, "dtype", q, dtype); num("count", q, cnt);
        uint64_t off = 0;
        size_t op;
        if (find_key(js, "offset", q, op)) {
            if (js[op] == '"') { get_string(js, "offset", q, offs); off = std::strtoull(offs.c_str(), nullptr, 10); }
            else off = (uint64_t)std::strtod(js.c_str() + op, nullptr);
        }
        if (dtype != "i4" && dtype != "f8") return bail("unknown dtype of array: " + name);
        const uint64_t elem = dtype == "i4" ? 4u : 8u;
        uint64_t ucnt;
        // overflow-safe: count <= sz / elem, then offset <= sz - bytes
        if (!as_count(cnt, (uint64_t)sz / elem, ucnt)) return bail("bad array count: " + name);
        const uint64_t bytes = ucnt * elem;
        if (off % 8 != 0 || off > (uint64_t)sz || bytes > (uint64_t)sz - off) return bail("array out of bounds: " + name);
        const char* ptr = F->blob.data() + off;
        const uint64_t N = (uint64_t)F->d.nCells, Fc = (uint64_t)F->d.nFaces;
        auto want = [&](uint64_t n, const char* dt) { return (uint64_t)cnt == n && dtype == dt; };
        if (name == "lowerAddr" && want(Fc, "i4")) F->d.lowerAddr = (const int32_t*)ptr;
        else if (name == "upperAddr" && want(Fc, "i4")) F->d.upperAddr = (const int32_t*)ptr;
        else if (name == "diag" && want(N, "f8")) F->d.diag = (const double*)ptr;
        else if (name == "upper" && want(Fc, "f8")) F->d.upper = (const double*)ptr;
        else if (name == "source" && want(N, "f8")) F->d.source = (const double*)ptr;
        else if (name == "psi0" && want(N, "f8")) F->d.psi0 = (const double*)ptr;
        else if (name == "psi" && want(N, "f8")) F->d.psiSolution = (const double*)ptr;
        else if (name == "lower" && want(Fc, "f8")) F->d.lower = (const double*)ptr;
        else if (name.compare(0, 5, "iface") == 0) {
            const size_t dot = name.find('.');
            const size_t k = (size_t)std::atoi(name.c_str() + 5);
            if (dot == std::string::npos || k >= F->ifaces.size() || (uint64_t)cnt != (uint64_t)F->ifaces[k].nFaces)
                return bail("bad interface array: " + name);
            if (name.substr(dot) == ".faceCells" && dtype == "i4") F->ifaces[k].faceCells = (const int32_t*)ptr;
            else if (name.substr(dot) == ".bouCoeffs" && dtype == "f8") F->bou[k] = (const double*)ptr;
        }
        p = q + 6;
    }
    const bool needF = F->d.nFaces > 0, needN = F->d.nCells > 0;
    if ((needF && (!F->d.lowerAddr || !F->d.upperAddr || !F->d.upper)) ||
        (needN && (!F->d.diag || !F->d.source || !F->d.psi0)))
        return bail("dump lacks a mandatory array (or its size does not match nCells/nFaces)");
    for (size_t k = 0; k < F->ifaces.size(); ++k)
        if (F->ifaces[k].nFaces > 0 && (!F->ifaces[k].faceCells || !F->bou[k])) return bail("dump lacks interface arrays");
    F->d.nIfaces = (int32_t)F->ifaces.size();
    F->d.ifaces = F->ifaces.data();
    F->d.ifaceBouCoeffs = F->bou.data();
    *out = F;
    return B200_OK;
}

const b200_dump* b200_dump_get(const b200_dump_file* f) { return f ? &f->d : nullptr; }
const char* b200_dump_header_json(const b200_dump_file* f) { return f ? f->json.c_str() : ""; }
void b200_dump_free(b200_dump_file* f) { delete f; }

}  // extern "C"
