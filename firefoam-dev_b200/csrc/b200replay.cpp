// b200replay -- re-solve dumped p_rgh systems (.b200sys, csrc/dump.cpp) through libb200pcg's C ABI.
//
//   b200replay [--precond none|diagonal|DIC|DIC-exact|DIC-eisenstat] [--repeat N] [--device D] file.b200sys ...
//
// Prints, per file, the OpenFOAM log line of the replayed solve (format of
// cases/steckler/original/linux64/log.fireFoam:92), the reference line stored in the dump (what the
// solver inside the host application reported) and the device / copy times.  Exit code 1 if a
// dump that carries a reference differs in iteration count when replayed with its own controls.
// Single-rank dumps only (nranks == 1): processor dumps need one process per rank.
#include "../../include/b200pcg.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static int precond_code(const std::string& s) {
    if (s == "none") return B200_PRECOND_NONE;
    if (s == "diagonal") return B200_PRECOND_DIAGONAL;
    if (s == "DIC") return B200_PRECOND_DIC_MC;
    if (s == "DIC-exact") return B200_PRECOND_DIC_EXACT;
    if (s == "DIC-eisenstat") return B200_PRECOND_DIC_MC_EIS;
    if (s == "DIC-multicolour") return B200_PRECOND_DIC_MC_LOOP;
    return -1;
}
static const char* precond_word(int c) {
    // only code 3 is OpenFOAM's DIC; the multicolour IC0 stand-in prints as DIC(mc) (adapter/B200PCG.C)
    return c == B200_PRECOND_NONE ? "none" : c == B200_PRECOND_DIAGONAL ? "diagonal"
         : c == B200_PRECOND_DIC_EXACT ? "DIC" : "DIC(mc)";
}

int main(int argc, char** argv) {
    int precond = -1, repeat = 1, device = -1;
    std::vector<std::string> files;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--precond" && i + 1 < argc) {
            precond = precond_code(argv[++i]);
            if (precond < 0) { std::fprintf(stderr, "unknown preconditioner\n"); return 2; }
        } else if (a == "--repeat" && i + 1 < argc) repeat = std::max(1, std::atoi(argv[++i]));
        else if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (a == "-h" || a == "--help") {
            std::printf("usage: b200replay [--precond none|diagonal|DIC|DIC-exact|DIC-eisenstat] [--repeat N] [--device D] file.b200sys ...\n");
            return 0;
        } else files.push_back(a);
    }
    if (files.empty()) { std::fprintf(stderr, "b200replay: no input files (see --help)\n"); return 2; }
    b200_ctx* ctx = nullptr;
    if (b200_ctx_create(device, 0, 1, nullptr, &ctx) != B200_OK) {
        std::fprintf(stderr, "b200replay: %s\n", b200_last_error(nullptr));
        return 3;
    }
    int mismatches = 0;
    uint64_t key = 1;
    for (const std::string& path : files) {
        b200_dump_file* f = nullptr;
        if (b200_dump_read(path.c_str(), &f) != B200_OK) {
            std::fprintf(stderr, "b200replay: %s\n", b200_dump_last_error());
            mismatches++;
            continue;
        }
        const b200_dump* d = b200_dump_get(f);
        if (d->nranks != 1 || d->nIfaces != 0) {
            std::fprintf(stderr, "b200replay: %s is rank %d of %d: multi-rank dumps need one process per rank\n",
                         path.c_str(), d->rank, d->nranks);
            b200_dump_free(f);
            mismatches++;
            continue;
        }
        int rc = b200_set_addressing(ctx, key++, d->nCells, d->nFaces, d->lowerAddr, d->upperAddr, 0, nullptr);
        if (rc != B200_OK) { std::fprintf(stderr, "b200replay: %s\n", b200_last_error(ctx)); return 3; }
        if (d->haveSmooth) {
            // a smoothSolver solve: asymmetric matrix, its own controls (--precond does not apply)
            b200_smooth_controls sc = d->smooth;
            sc.reserved = 0;
            std::vector<double> psiS((size_t)d->nCells);
            b200_perf pf;
            double bestS = 1e300;
            for (int r = 0; r < repeat; ++r) {
                std::memcpy(psiS.data(), d->psi0, sizeof(double) * (size_t)d->nCells);
                rc = b200_smooth_solve(ctx, d->diag, d->upper, d->lower, nullptr, d->source, psiS.data(), &sc, &pf);
                if (rc != B200_OK && rc != B200_ENONFINITE) { std::fprintf(stderr, "b200replay: %s\n", b200_last_error(ctx)); return 3; }
                bestS = std::min(bestS, pf.setupMs + pf.solveMs);
            }
            const bool exactS = sc.sweepMode == B200_SWEEP_EXACT;
            std::printf("%s  [%d cells, %d faces, %s, solve %d, t = %g]\n", path.c_str(), d->nCells, d->nFaces,
                        d->lower ? "asymmetric" : "symmetric", d->solveIndex, d->time);
            std::printf("  B200smoothSolver%s:  Solving for %s, Initial residual = %.8g, Final residual = %.8g, No Iterations %d\n",
                        exactS ? "" : "(mc)", d->fieldName, pf.initialResidual, pf.finalResidual, pf.nIterations);
            if (d->havePerf)
                std::printf("  %s:  Solving for %s, Initial residual = %.8g, Final residual = %.8g, No Iterations %d   (dumped reference)\n",
                            d->solverName, d->fieldName, d->perf.initialResidual, d->perf.finalResidual, d->perf.nIterations);
            std::printf("  device: set-up + solve %.3f ms (best of %d), H2D %.3f ms, D2H %.3f ms\n", bestS, repeat, pf.h2dMs, pf.d2hMs);
            if (exactS && d->havePerf && pf.nIterations != d->perf.nIterations) {
                std::printf("  MISMATCH: sweep count differs from the dumped reference\n");
                mismatches++;
            }
            b200_dump_free(f);
            continue;
        }
        if (d->havePBiCG) {
            // a PBiCG solve: its own controls (--precond does not apply)
            b200_controls bc = d->controls;
            bc.reserved = 0;
            std::vector<double> psiB((size_t)d->nCells);
            b200_perf pf;
            double bestB = 1e300;
            for (int r = 0; r < repeat; ++r) {
                std::memcpy(psiB.data(), d->psi0, sizeof(double) * (size_t)d->nCells);
                rc = b200_bicg_solve(ctx, d->diag, d->upper, d->lower, nullptr, nullptr, d->source, psiB.data(), &bc, &pf);
                if (rc != B200_OK && rc != B200_ENONFINITE) { std::fprintf(stderr, "b200replay: %s\n", b200_last_error(ctx)); return 3; }
                bestB = std::min(bestB, pf.setupMs + pf.solveMs);
            }
            const char* pw = bc.precond == B200_PRECOND_NONE ? "none" : bc.precond == B200_PRECOND_DIAGONAL ? "diagonal"
                           : bc.precond == B200_PRECOND_DILU_EXACT ? "DILU" : "DILU(mc)";
            std::printf("%s  [%d cells, %d faces, %s, solve %d, t = %g]\n", path.c_str(), d->nCells, d->nFaces,
                        d->lower ? "asymmetric" : "symmetric", d->solveIndex, d->time);
            std::printf("  %sB200PBiCG:  Solving for %s, Initial residual = %.8g, Final residual = %.8g, No Iterations %d\n",
                        pw, d->fieldName, pf.initialResidual, pf.finalResidual, pf.nIterations);
            if (d->havePerf)
                std::printf("  %s:  Solving for %s, Initial residual = %.8g, Final residual = %.8g, No Iterations %d   (dumped reference)\n",
                            d->solverName, d->fieldName, d->perf.initialResidual, d->perf.finalResidual, d->perf.nIterations);
            std::printf("  device: set-up + solve %.3f ms (best of %d), H2D %.3f ms, D2H %.3f ms\n", bestB, repeat, pf.h2dMs, pf.d2hMs);
            if (bc.precond != B200_PRECOND_DILU_MC && bc.precond != B200_PRECOND_NONE && d->havePerf && pf.nIterations != d->perf.nIterations) {
                std::printf("  MISMATCH: iteration count differs from the dumped reference\n");
                mismatches++;
            }
            b200_dump_free(f);
            continue;
        }
        b200_controls ctl = d->controls;
        const bool own = precond < 0;
        if (!own) ctl.precond = precond;
        ctl.reserved = 0;
        std::vector<double> psi((size_t)d->nCells);
        b200_perf perf;
        double best = 1e300;
        for (int r = 0; r < repeat; ++r) {
            std::memcpy(psi.data(), d->psi0, sizeof(double) * (size_t)d->nCells);
            rc = b200_solve(ctx, d->diag, d->upper, nullptr, d->source, psi.data(), &ctl, &perf);
            if (rc != B200_OK && rc != B200_ENONFINITE) { std::fprintf(stderr, "b200replay: %s\n", b200_last_error(ctx)); return 3; }
            best = std::min(best, perf.setupMs + perf.solveMs);
        }
        std::printf("%s  [%d cells, %d faces, solve %d, t = %g]\n", path.c_str(), d->nCells, d->nFaces, d->solveIndex, d->time);
        std::printf("  %sB200PCG:  Solving for %s, Initial residual = %.8g, Final residual = %.8g, No Iterations %d\n",
                    precond_word(ctl.precond), d->fieldName, perf.initialResidual, perf.finalResidual, perf.nIterations);
        if (d->havePerf)
            std::printf("  %s:  Solving for %s, Initial residual = %.8g, Final residual = %.8g, No Iterations %d   (dumped reference)\n",
                        d->solverName, d->fieldName, d->perf.initialResidual, d->perf.finalResidual, d->perf.nIterations);
        if (d->psiSolution) {
            double num = 0, den = 0;
            for (int i = 0; i < d->nCells; ++i) {
                num = std::max(num, std::fabs(psi[i] - d->psiSolution[i]));
                den = std::max(den, std::fabs(d->psiSolution[i]));
            }
            std::printf("  max |psi - dumped psi| / max |dumped psi| = %.3e\n", den > 0 ? num / den : num);
        }
        std::printf("  device: set-up + solve %.3f ms (best of %d), H2D %.3f ms, D2H %.3f ms\n", best, repeat, perf.h2dMs, perf.d2hMs);
        if (own && d->havePerf && ctl.precond != B200_PRECOND_DIC_MC && ctl.precond != B200_PRECOND_DIC_MC_EIS && ctl.precond != B200_PRECOND_DIC_MC_LOOP && perf.nIterations != d->perf.nIterations) {
            std::printf("  MISMATCH: iteration count differs from the dumped reference\n");
            mismatches++;
        }
        b200_dump_free(f);
    }
    b200_ctx_destroy(ctx);
    return mismatches ? 1 : 0;
}
