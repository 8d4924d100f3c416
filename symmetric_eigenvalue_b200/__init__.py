"""B200-native Cuppen divide-and-conquer eigensolver for symmetric tridiagonal matrices.

Python front end of ``libcuppen_b200.so`` (hand-written sm_100a CUDA behind the C ABI declared in
``include/cuppen_b200.h``).  The names mirror the reference's C interface for this path
(chrhenning/symmetric_eigenvalue ``src/helper.h``, ``src/filehandling.h``, ``src/eigenvalues.h``).
There is no CPU implementation in this package: without the CUDA library every call fails loudly.
"""
from .api import (  # noqa: F401
    CuppenError, CuppenSolver, MergeStat, createMatrixScheme1, createMatrixScheme2,
    readSymmTriadiagonalMatrixFromSparseMTX, determineEigenvectorsToCompute, writeResults,
    nccl_unique_id, load_library, library_path, cuppens, read_eigenvector_file,
)

__all__ = [
    "CuppenError", "CuppenSolver", "MergeStat", "createMatrixScheme1", "createMatrixScheme2",
    "readSymmTriadiagonalMatrixFromSparseMTX", "determineEigenvectorsToCompute", "writeResults",
    "nccl_unique_id", "load_library", "library_path", "cuppens", "read_eigenvector_file",
]
