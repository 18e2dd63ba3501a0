"""firefoam-dev_b200: B200-native p_rgh pressure-correction hot path (assemble + PCG solve).

Only what the path needs: csrc/ (sm_100a CUDA kernels + C ABI + harness generators) and the
host-side mirror of the OpenFOAM interfaces it sits behind (ldu.py, fvm.py, cases.py)."""
from ._lib import B200Error, build, load_pcg, load_mesh, ABI_SYMBOLS, PRECOND  # noqa: F401
from .ldu import (B200PBiCG, B200PCG, B200smoothSolver, Context, LduAddressing, LduMatrix, ProcessorLduInterface,  # noqa: F401
                  SolverPerformance, default_context, make_bicg_controls, make_controls, make_smooth_controls)
