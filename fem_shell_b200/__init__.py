"""fem_shell_b200 -- B200-native hot path of precice/fem-shell (assembly -> PCG solve).

The product is the C-ABI shared library ``libfemshell_b200.so`` (sources in ``csrc/``, interface in
``include/femshell_b200.h``).  This package is only the ctypes binding used by the tests and the
benchmark; it contains no numerical fallback: if the library is missing or no CUDA device is
present, construction fails loudly.
"""
from .api import (  # noqa: F401
    FemShell, FemShellError, SolveInfo, load_library, meshgen, partition_plan, gather_plan, read_xda, read_mesh, read_forces, write_xda, write_xdr,
    TRI3, QUAD4, DOF_FIRST_ENCOUNTER, DOF_NODE_ID, PC_NONE, PC_JACOBI, PC_BJACOBI6, PC_MLRBM,
    NORM_UNPRECONDITIONED, NORM_PRECONDITIONED, QUIRKS_REFERENCE, ASM_COLORED, ASM_GATHER, SPMV_AUTO, SPMV_FULL, COMM_AUTO, COMM_NCCL, COMM_PEER,
    FS_OK, FS_ERR_NOT_CONVERGED, FS_ERR_BREAKDOWN, LIB_PATH, EXPORTED_SYMBOLS,
)
