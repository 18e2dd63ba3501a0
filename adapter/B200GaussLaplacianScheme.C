/*---------------------------------------------------------------------------*\
  B200GaussLaplacianScheme.C -- see B200GaussLaplacianScheme.H.  Thin glue:
  OpenFOAM fields -> b200_assemble_laplacian (include/b200pcg.h) -> fvScalarMatrix.
\*---------------------------------------------------------------------------*/

#include "B200GaussLaplacianScheme.H"
#include "B200Context.H"
#include "fvMesh.H"
#include "addToRunTimeSelectionTable.H"

// * * * * * * * * * * * * * * Static Data Members * * * * * * * * * * * * * //

namespace Foam
{
namespace fv
{
    defineTypeNameAndDebug(B200GaussLaplacianScheme, 0);

    laplacianScheme<scalar, scalar>::
        addIstreamConstructorToTable<B200GaussLaplacianScheme>
        addB200GaussScalarScalarIstreamConstructorToTable_;
}
}


// * * * * * * * * * * * * * * * Member Functions  * * * * * * * * * * * * * //

Foam::tmp<Foam::fvScalarMatrix>
Foam::fv::B200GaussLaplacianScheme::fvmLaplacian
(
    const surfaceScalarField& gamma,
    const volScalarField& vf
)
{
    const fvMesh& mesh = this->mesh();

    // the explicit non-orthogonal correction is not part of the hot path
    // (cases/steckler/system/fvSchemes:63-76 use `uncorrected`): stock scheme
    if (this->tsnGradScheme_().corrected())
    {
        return gaussLaplacianScheme<scalar, scalar>::fvmLaplacian(gamma, vf);
    }

    tmp<surfaceScalarField> tdeltaCoeffs
    (
        this->tsnGradScheme_().deltaCoeffs(vf)
    );
    const surfaceScalarField& deltaCoeffs = tdeltaCoeffs();
    const surfaceScalarField& magSf = mesh.magSf();

    tmp<fvScalarMatrix> tfvm
    (
        new fvScalarMatrix
        (
            vf,
            deltaCoeffs.dimensions()*gamma.dimensions()*magSf.dimensions()
           *vf.dimensions()
        )
    );
    fvScalarMatrix& fvm = tfvm.ref();

    // --- internal faces: upper = deltaCoeffs*(gamma*magSf); negSumDiag.
    //     lduMatrix::upper()/diag() allocate zero-initialised fields, which is
    //     what b200_assemble_laplacian expects in diag_inout for a pure laplacian
    scalarField& upper = fvm.upper();
    scalarField& diag = fvm.diag();

    b200_ctx* ctx = b200Context();
    labelList coupledPatches;
    b200SetAddressing(ctx, mesh.lduAddr(), mesh.interfaces(), coupledPatches);

    if
    (
        b200_assemble_laplacian
        (
            ctx,
            gamma.primitiveField().begin(),
            magSf.primitiveField().begin(),
            deltaCoeffs.primitiveField().begin(),
            1.0,                 // `- fvm::laplacian` negates afterwards (fvMatrix::operator-)
            upper.begin(),
            diag.begin()
        ) != B200_OK
    )
    {
        FatalErrorInFunction
            << "B200Gauss: " << b200_last_error(ctx) << exit(FatalError);
    }

    // --- patches: as gaussLaplacianScheme::fvmLaplacianUncorrected
    forAll(vf.boundaryField(), patchi)
    {
        const fvPatchScalarField& pvf = vf.boundaryField()[patchi];
        const scalarField pGamma
        (
            gamma.boundaryField()[patchi]*magSf.boundaryField()[patchi]
        );
        const fvsPatchScalarField& pDeltaCoeffs =
            deltaCoeffs.boundaryField()[patchi];

        if (pvf.coupled())
        {
            fvm.internalCoeffs()[patchi] =
                pGamma*pvf.gradientInternalCoeffs(pDeltaCoeffs);
            fvm.boundaryCoeffs()[patchi] =
               -pGamma*pvf.gradientBoundaryCoeffs(pDeltaCoeffs);
        }
        else
        {
            fvm.internalCoeffs()[patchi] = pGamma*pvf.gradientInternalCoeffs();
            fvm.boundaryCoeffs()[patchi] = -pGamma*pvf.gradientBoundaryCoeffs();
        }
    }

    return tfvm;
}

// ************************************************************************* //
