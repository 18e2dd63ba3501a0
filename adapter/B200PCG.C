/*---------------------------------------------------------------------------*\
  B200PCG.C -- see B200PCG.H.  Thin glue: OpenFOAM objects -> C ABI of
  include/b200pcg.h -> solverPerformance.  No arithmetic here.
\*---------------------------------------------------------------------------*/

#include "B200PCG.H"
#include "B200Context.H"
#include "processorLduInterface.H"
#include "processorLduInterfaceField.H"
#include "Pstream.H"
#include "DynamicList.H"

#include "b200pcg.h"

#include <cstdlib>
#include <cstdint>
#include <cstdio>
#include <string>

// * * * * * * * * * * * * * * Static Data Members * * * * * * * * * * * * * //

namespace Foam
{
    defineTypeNameAndDebug(B200PCG, 0);

    lduMatrix::solver::addsymMatrixConstructorToTable<B200PCG>
        addB200PCGSymMatrixConstructorToTable_;

    // labels are handed to the library as int32, scalars as double
    static_assert(sizeof(label) == 4, "B200PCG needs WM_LABEL_SIZE=32");
    static_assert(sizeof(scalar) == 8, "B200PCG needs WM_PRECISION_OPTION=DP");
}


// * * * * * * * * * * * * * * * Shared Functions * * * * * * * * * * * * * * //

//- One library context per MPI rank, created on first use.
//  rank r drives GPU (r mod deviceCount); the NCCL unique id is made on the
//  master and scattered with OpenFOAM's own Pstream.
b200_ctx* Foam::b200Context()
{
    static b200_ctx* ctx = nullptr;
    if (ctx) return ctx;

    const int nDev = b200_device_count();
    if (nDev < 1)
    {
        FatalErrorInFunction
            << "B200PCG: no CUDA device visible; this solver has no CPU fallback"
            << exit(FatalError);
    }

    List<char> uid(128, '\0');
    if (Pstream::parRun())
    {
        if (Pstream::master())
        {
            if (b200_get_unique_id(uid.begin()) != B200_OK)
            {
                FatalErrorInFunction
                    << "B200PCG: " << b200_last_error(nullptr) << exit(FatalError);
            }
        }
        Pstream::scatter(uid);
    }

    const int rank = Pstream::parRun() ? Pstream::myProcNo() : 0;
    const int nRanks = Pstream::parRun() ? Pstream::nProcs() : 1;

    if
    (
        b200_ctx_create
        (
            rank % nDev, rank, nRanks,
            nRanks > 1 ? uid.begin() : nullptr,
            &ctx
        ) != B200_OK
    )
    {
        FatalErrorInFunction
            << "B200PCG: " << b200_last_error(nullptr) << exit(FatalError);
    }

    if (std::getenv("B200PCG_PROFILE")) b200_profile_enable(ctx, 1);

    return ctx;
}


void Foam::b200SetAddressing
(
    b200_ctx* ctx,
    const lduAddressing& addr,
    const lduInterfacePtrsList& interfaces,
    labelList& coupledPatches
)
{
    // --- coupled (processor) interfaces; the UPtrList has null slots for
    //     non-coupled patches
    DynamicList<b200_iface> ifaces(interfaces.size());
    DynamicList<label> patches(interfaces.size());

    forAll(interfaces, patchi)
    {
        if (!interfaces.set(patchi)) continue;

        // processorCyclic patches ARE processorLduInterfaces (processorCyclicFvPatch derives from
        // processorFvPatch) but apply a transformation to the neighbour values: rejected by type name
        const word ifType(interfaces[patchi].type());
        if
        (
            !isA<processorLduInterface>(interfaces[patchi])
         || ifType.find("Cyclic") != std::string::npos
        )
        {
            FatalErrorInFunction
                << "B200PCG: unsupported coupled interface "
                << ifType << " on patch " << patchi
                << " (only plain processor interfaces are supported)"
                << exit(FatalError);
        }

        const processorLduInterface& pi =
            refCast<const processorLduInterface>(interfaces[patchi]);

        const labelUList& faceCells = addr.patchAddr(patchi);

        b200_iface itf;
        itf.nbrRank = pi.neighbProcNo();
        itf.nFaces = faceCells.size();
        itf.faceCells = faceCells.begin();
        itf.tag = pi.tag();
        ifaces.append(itf);
        patches.append(patchi);
    }

    // idempotent per mesh: keyed by the address of the lduAddressing object
    if
    (
        b200_set_addressing
        (
            ctx,
            uint64_t(uintptr_t(&addr)),
            addr.size(),
            addr.lowerAddr().size(),
            addr.lowerAddr().begin(),
            addr.upperAddr().begin(),
            ifaces.size(),
            ifaces.begin()
        ) != B200_OK
    )
    {
        FatalErrorInFunction
            << "B200PCG: " << b200_last_error(ctx) << exit(FatalError);
    }

    coupledPatches.transfer(patches);
}


void Foam::b200LduInterfaces
(
    const lduInterfaceFieldPtrsList& interfaceFields,
    lduInterfacePtrsList& lduInterfaces
)
{
    forAll(interfaceFields, patchi)
    {
        if (interfaceFields.set(patchi))
        {
            const word fieldType(interfaceFields[patchi].type());
            if
            (
                !isA<processorLduInterfaceField>(interfaceFields[patchi])
             || fieldType.find("Cyclic") != std::string::npos
            )
            {
                FatalErrorInFunction
                    << "B200PCG: unsupported coupled interface "
                    << interfaceFields[patchi].type() << " on patch " << patchi
                    << " (only processor interfaces are supported)"
                    << exit(FatalError);
            }
            lduInterfaces.set(patchi, &interfaceFields[patchi].interface());
        }
    }
}


// * * * * * * * * * * * * * * * * Constructors  * * * * * * * * * * * * * * //

Foam::B200PCG::B200PCG
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

Foam::solverPerformance Foam::B200PCG::solve
(
    scalarField& psi,
    const scalarField& source,
    const direction cmpt
) const
{
    const word preconditionerName(lduMatrix::preconditioner::getName(controlDict_));

    // --- Setup class containing solver performance data
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
    else if (preconditionerName == "DIC" || preconditionerName == "FDIC")
    {
        const word mode
        (
            controlDict_.subOrEmptyDict("B200").lookupOrDefault<word>
            (
                "dicMode", "auto"
            )
        );
        // `preconditioner DIC` is the DIC-CLASS multicolour IC0 (the library picks its form); OpenFOAM's own
        // DIC -- same elimination order, same iteration counts -- is `B200 { dicMode exact; }`
        ctl.precond =
            (mode == "exact") ? B200_PRECOND_DIC_EXACT
          : (mode == "eisenstat") ? B200_PRECOND_DIC_MC_EIS
          : (mode == "multicolour") ? B200_PRECOND_DIC_MC_LOOP
          : B200_PRECOND_DIC_MC;
        if (mode != "exact")
        {
            // the log line names what ran: DIC(mc)B200PCG is not DICPCG (different iteration counts)
            logPreconditionerName = "DIC(mc)";
        }
    }
    else
    {
        FatalErrorInFunction
            << "B200PCG: unsupported preconditioner " << preconditionerName
            << "; valid: none diagonal DIC" << exit(FatalError);
    }

    solverPerformance solverPerf(logPreconditionerName + typeName, fieldName_);

    if (!matrix_.symmetric() && !matrix_.diagonal())
    {
        FatalErrorInFunction
            << "B200PCG: matrix is not symmetric" << exit(FatalError);
    }

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

    // the dump (below) wants the interface list once more
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

    // --- B200PCG_DUMP=<dir>: keep the initial guess so that the system can be written out after the
    //     solve together with what the solver reported (include/b200pcg.h b200_dump; replayed with
    //     b200replay / firefoam-dev_b200/replay.py)
    const char* dumpDir = std::getenv("B200PCG_DUMP");
    scalarField psi0;
    if (dumpDir) psi0 = psi;

    b200_perf perf;

    const int rc = b200_solve
    (
        ctx,
        matrix_.diag().begin(),
        matrix_.upper().begin(),
        bou.begin(),
        source.begin(),
        psi.begin(),
        &ctl,
        &perf
    );

    if (rc != B200_OK)
    {
        FatalErrorInFunction
            << "B200PCG: " << b200_last_error(ctx) << exit(FatalError);
    }

    if (dumpDir)
    {
        static int solveIndex = 0;
        b200_dump d = b200_dump();   // (lower / haveSmooth of ABI version 2 stay zero: a symmetric PCG solve)
        d.fieldName = fieldName_.c_str();
        d.rank = Pstream::parRun() ? Pstream::myProcNo() : 0;
        d.nranks = Pstream::parRun() ? Pstream::nProcs() : 1;
        d.nCells = addr.size();
        d.nFaces = addr.lowerAddr().size();
        d.lowerAddr = addr.lowerAddr().begin();
        d.upperAddr = addr.upperAddr().begin();
        d.diag = matrix_.diag().begin();
        d.upper = matrix_.upper().begin();
        d.source = source.begin();
        d.psi0 = psi0.begin();
        d.psiSolution = psi.begin();
        d.nIfaces = ifaces.size();
        d.ifaces = ifaces.begin();
        d.ifaceBouCoeffs = bou.begin();
        d.controls = ctl;
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
                << "B200PCG: " << b200_dump_last_error() << endl;
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
