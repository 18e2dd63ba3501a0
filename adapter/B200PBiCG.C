/*---------------------------------------------------------------------------*\
  B200PBiCG.C -- see B200PBiCG.H.  Thin glue: OpenFOAM objects ->
  b200_bicg_solve (include/b200pcg.h) -> solverPerformance.  No arithmetic here.
\*---------------------------------------------------------------------------*/

#include "B200PBiCG.H"
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
    defineTypeNameAndDebug(B200PBiCG, 0);

    lduMatrix::solver::addasymMatrixConstructorToTable<B200PBiCG>
        addB200PBiCGAsymMatrixConstructorToTable_;
}


// * * * * * * * * * * * * * * * * Constructors  * * * * * * * * * * * * * * //

Foam::B200PBiCG::B200PBiCG
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
    )
{}


// * * * * * * * * * * * * * * * Member Functions  * * * * * * * * * * * * * //

Foam::solverPerformance Foam::B200PBiCG::solve
(
    scalarField& psi,
    const scalarField& source,
    const direction cmpt
) const
{
    const word preconditionerName(lduMatrix::preconditioner::getName(controlDict_));
    word logPreconditionerName(preconditionerName);

    b200_controls ctl;
    ctl.tolerance = tolerance_;
    ctl.relTol = relTol_;
    ctl.maxIter = maxIter_;
    ctl.minIter = minIter_;
    ctl.reserved = 0;

    if (preconditionerName == "none")
    {
        ctl.precond = B200_PRECOND_NONE;
    }
    else if (preconditionerName == "diagonal")
    {
        ctl.precond = B200_PRECOND_DIAGONAL;
    }
    else if (preconditionerName == "DILU")
    {
        const word mode
        (
            controlDict_.subOrEmptyDict("B200").lookupOrDefault<word>
            (
                "diluMode", "multicolour"
            )
        );
        if (mode != "multicolour" && mode != "exact")
        {
            FatalErrorInFunction
                << "B200PBiCG: unknown diluMode " << mode
                << "; valid: multicolour exact" << exit(FatalError);
        }
        ctl.precond = (mode == "exact") ? B200_PRECOND_DILU_EXACT : B200_PRECOND_DILU_MC;
        if (mode != "exact")
        {
            // the log line names what ran: DILU(mc)B200PBiCG is not DILUPBiCG (different iteration counts)
            logPreconditionerName = "DILU(mc)";
        }
    }
    else
    {
        FatalErrorInFunction
            << "B200PBiCG: unsupported preconditioner " << preconditionerName
            << "; valid: none diagonal DILU" << exit(FatalError);
    }

    solverPerformance solverPerf(logPreconditionerName + typeName, fieldName_);

    // --- mesh addressing + coupled (processor) interfaces (idempotent per mesh)
    const lduAddressing& addr = matrix_.lduAddr();

    lduInterfacePtrsList lduInterfaces(interfaces_.size());
    b200LduInterfaces(interfaces_, lduInterfaces);

    b200_ctx* ctx = b200Context();
    labelList coupledPatches;
    b200SetAddressing(ctx, addr, lduInterfaces, coupledPatches);

    // Amul takes interfaceBouCoeffs, Tmul interfaceIntCoeffs (PBiCG.C)
    DynamicList<const double*> bou(coupledPatches.size());
    DynamicList<const double*> intc(coupledPatches.size());
    forAll(coupledPatches, i)
    {
        bou.append(interfaceBouCoeffs_[coupledPatches[i]].begin());
        intc.append(interfaceIntCoeffs_[coupledPatches[i]].begin());
    }

    // --- B200PCG_DUMP=<dir>: the system + what the solver reported (include/b200pcg.h b200_dump, havePBiCG)
    const char* dumpDir = std::getenv("B200PCG_DUMP");
    scalarField psi0;
    if (dumpDir) psi0 = psi;

    b200_perf perf;

    // a diagonal matrix has no off-diagonals at all; a symmetric one has no lower()
    const bool faces = !matrix_.diagonal();

    const int rc = b200_bicg_solve
    (
        ctx,
        matrix_.diag().begin(),
        faces ? matrix_.upper().begin() : nullptr,
        (faces && matrix_.asymmetric()) ? matrix_.lower().begin() : nullptr,
        bou.begin(),
        intc.begin(),
        source.begin(),
        psi.begin(),
        &ctl,
        &perf
    );

    if (rc != B200_OK)
    {
        FatalErrorInFunction
            << "B200PBiCG: " << b200_last_error(ctx) << exit(FatalError);
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
        d.controls = ctl;
        d.havePBiCG = 1;
        d.havePerf = 1;
        d.perf = perf;
        const std::string name(logPreconditionerName + typeName);
        d.solverName = name.c_str();
        d.solveIndex = solveIndex;
        d.time = matrix_.mesh().thisDb().time().value();

        char file[64];
        std::snprintf(file, sizeof(file), "_%06d_p%d.b200sys", solveIndex++, d.rank);
        const std::string path(std::string(dumpDir) + "/" + fieldName_ + file);
        if (b200_dump_write(path.c_str(), &d) != B200_OK)
        {
            WarningInFunction
                << "B200PBiCG: " << b200_dump_last_error() << endl;
        }
    }

    solverPerf.initialResidual() = perf.initialResidual;
    solverPerf.finalResidual() = perf.finalResidual;
    solverPerf.nIterations() = perf.nIterations;
    solverPerf.checkConvergence(tolerance_, relTol_);
    if (perf.singular)
    {
        solverPerf.checkSingularity(0);   // flags the component as singular
    }

    return solverPerf;
}
