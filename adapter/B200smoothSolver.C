/*---------------------------------------------------------------------------*\
  B200smoothSolver.C -- see B200smoothSolver.H.  Thin glue: OpenFOAM objects ->
  b200_smooth_solve (include/b200pcg.h) -> solverPerformance.  No arithmetic here.
\*---------------------------------------------------------------------------*/

#include "B200smoothSolver.H"
#include "B200Context.H"
#include "processorLduInterface.H"
#include "Pstream.H"
#include "DynamicList.H"

#include "b200pcg.h"

#include <cstdlib>
#include <cstdio>
#include <string>

// * * * * * * * * * * * * * * Static Data Members * * * * * * * * * * * * * //

namespace Foam
{
    defineTypeNameAndDebug(B200smoothSolver, 0);

    lduMatrix::solver::addsymMatrixConstructorToTable<B200smoothSolver>
        addB200smoothSolverSymMatrixConstructorToTable_;

    lduMatrix::solver::addasymMatrixConstructorToTable<B200smoothSolver>
        addB200smoothSolverAsymMatrixConstructorToTable_;
}


// * * * * * * * * * * * * * * * * Constructors  * * * * * * * * * * * * * * //

Foam::B200smoothSolver::B200smoothSolver
(
    const word& fieldName,
    const lduMatrix& matrix,
    const FieldField<Field, scalar>& interfaceBouCoeffs,
    const FieldField<Field, scalar>& interfaceIntCoeffs,
    const lduInterfaceFieldPtrsList& interfaces,
    const dictionary& solverControls
)
:
    lduMatrix::solver
    (
        fieldName,
        matrix,
        interfaceBouCoeffs,
        interfaceIntCoeffs,
        interfaces,
        solverControls
    ),
    nSweeps_(1)
{
    readControls();
}


// * * * * * * * * * * * * * * * Member Functions  * * * * * * * * * * * * * //

void Foam::B200smoothSolver::readControls()
{
    lduMatrix::solver::readControls();
    nSweeps_ = controlDict_.lookupOrDefault<label>("nSweeps", 1);
}


Foam::solverPerformance Foam::B200smoothSolver::solve
(
    scalarField& psi,
    const scalarField& source,
    const direction cmpt
) const
{
    b200_smooth_controls ctl;
    ctl.tolerance = tolerance_;
    ctl.relTol = relTol_;
    ctl.maxIter = maxIter_;
    ctl.minIter = minIter_;
    ctl.nSweeps = nSweeps_;
    ctl.reserved = 0;

    const word smootherName(controlDict_.lookup("smoother"));
    if (smootherName == "symGaussSeidel")
    {
        ctl.smoother = B200_SMOOTHER_SYM_GAUSS_SEIDEL;
    }
    else if (smootherName == "GaussSeidel")
    {
        ctl.smoother = B200_SMOOTHER_GAUSS_SEIDEL;
    }
    else
    {
        FatalErrorInFunction
            << "B200smoothSolver: unsupported smoother " << smootherName
            << "; valid: GaussSeidel symGaussSeidel" << exit(FatalError);
    }

    const word mode
    (
        controlDict_.subOrEmptyDict("B200").lookupOrDefault<word>
        (
            "sweepMode", "multicolour"
        )
    );
    if (mode != "multicolour" && mode != "exact")
    {
        FatalErrorInFunction
            << "B200smoothSolver: unknown sweepMode " << mode
            << "; valid: multicolour exact" << exit(FatalError);
    }
    ctl.sweepMode = (mode == "exact") ? B200_SWEEP_EXACT : B200_SWEEP_MULTICOLOUR;

    // the log line names what ran: only `sweepMode exact` visits the cells in smoothSolver's order
    solverPerformance solverPerf
    (
        (mode == "exact") ? word(typeName) : word(typeName + "(mc)"),
        fieldName_
    );

    // --- mesh addressing + coupled (processor) interfaces (idempotent per mesh)
    const lduAddressing& addr = matrix_.lduAddr();

    lduInterfacePtrsList lduInterfaces(interfaces_.size());
    b200LduInterfaces(interfaces_, lduInterfaces);

    b200_ctx* ctx = b200Context();
    labelList coupledPatches;
    b200SetAddressing(ctx, addr, lduInterfaces, coupledPatches);

    DynamicList<const double*> bou(coupledPatches.size());
    forAll(coupledPatches, i)
    {
        bou.append(interfaceBouCoeffs_[coupledPatches[i]].begin());
    }

    // --- B200PCG_DUMP=<dir>: keep the initial guess so that the system can be written out after the solve together
    //     with what the solver reported (include/b200pcg.h b200_dump: `lower` and the smoothSolver controls travel)
    const char* dumpDir = std::getenv("B200PCG_DUMP");
    scalarField psi0;
    if (dumpDir) psi0 = psi;

    b200_perf perf;

    // a diagonal matrix has no off-diagonals at all; a symmetric one has no lower()
    const bool faces = !matrix_.diagonal();

    const int rc = b200_smooth_solve
    (
        ctx,
        matrix_.diag().begin(),
        faces ? matrix_.upper().begin() : nullptr,
        (faces && matrix_.asymmetric()) ? matrix_.lower().begin() : nullptr,
        bou.begin(),
        source.begin(),
        psi.begin(),
        &ctl,
        &perf
    );

    if (rc != B200_OK)
    {
        FatalErrorInFunction
            << "B200smoothSolver: " << b200_last_error(ctx) << exit(FatalError);
    }

    if (dumpDir)
    {
        static int solveIndex = 0;
        DynamicList<b200_iface> ifaces(coupledPatches.size());
        forAll(coupledPatches, i)
        {
            const label patchi = coupledPatches[i];
            const processorLduInterface& pi =
                refCast<const processorLduInterface>(lduInterfaces[patchi]);
            const labelUList& faceCells = addr.patchAddr(patchi);
            b200_iface itf;
            itf.nbrRank = pi.neighbProcNo();
            itf.nFaces = faceCells.size();
            itf.faceCells = faceCells.begin();
            itf.tag = pi.tag();
            ifaces.append(itf);
        }
        b200_dump d = b200_dump();
        d.fieldName = fieldName_.c_str();
        d.rank = Pstream::parRun() ? Pstream::myProcNo() : 0;
        d.nranks = Pstream::parRun() ? Pstream::nProcs() : 1;
        d.nCells = addr.size();
        d.nFaces = addr.lowerAddr().size();
        d.lowerAddr = addr.lowerAddr().begin();
        d.upperAddr = addr.upperAddr().begin();
        d.diag = matrix_.diag().begin();
        d.upper = faces ? matrix_.upper().begin() : nullptr;
        d.lower = (faces && matrix_.asymmetric()) ? matrix_.lower().begin() : nullptr;
        d.source = source.begin();
        d.psi0 = psi0.begin();
        d.psiSolution = psi.begin();
        d.nIfaces = ifaces.size();
        d.ifaces = ifaces.begin();
        d.ifaceBouCoeffs = bou.begin();
        d.haveSmooth = 1;
        d.smooth = ctl;
        d.havePerf = 1;
        d.perf = perf;
        const std::string name((mode == "exact") ? word(typeName) : word(typeName + "(mc)"));
        d.solverName = name.c_str();
        d.solveIndex = solveIndex;
        d.time = matrix_.mesh().thisDb().time().value();

        char file[64];
        std::snprintf(file, sizeof(file), "_%06d_p%d.b200sys", solveIndex++, d.rank);
        const std::string path(std::string(dumpDir) + "/" + fieldName_ + file);
        if (b200_dump_write(path.c_str(), &d) != B200_OK)
        {
            WarningInFunction
                << "B200smoothSolver: " << b200_dump_last_error() << endl;
        }
    }

    solverPerf.initialResidual() = perf.initialResidual;
    solverPerf.finalResidual() = perf.finalResidual;
    solverPerf.nIterations() = perf.nIterations;
    if (nSweeps_ > 0)
    {
        solverPerf.checkConvergence(tolerance_, relTol_);
    }

    return solverPerf;
}
