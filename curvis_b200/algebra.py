"""Host mirror of reference src/algebra.rs: ``Orientation`` (a forward/up pair turned into a
rotation).  The arithmetic runs in libcurvis_b200's host helper ``curvis_orientation`` (C++,
nalgebra operation order) so Python, C++ and Rust hosts hand identical doubles to the kernel."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi


class Orientation:
    """``Orientation::new(forward, up)`` (src/algebra.rs:16-38).

    Raises :class:`CurvisError` (``ERR_PARALLEL_VECTORS``) where the reference panics with
    "Forward and up vectors must not be parallel" (:19-21).
    """

    def __init__(self, forward, up):
        lib = _abi.load_library()
        f, u = _abi.dvec(forward, 3), _abi.dvec(up, 3)
        rot, inv, up_o = (C.c_double * 9)(), (C.c_double * 9)(), (C.c_double * 3)()
        _abi.check(lib.curvis_orientation(f, u, rot, inv, up_o))
        self._forward = np.array(forward, dtype=np.float64)       # kept as given (:33)
        self._up = np.array(up_o, dtype=np.float64)               # orthogonalised (:30)
        self._rotation = np.array(rot, dtype=np.float64).reshape(3, 3)
        self._inverse = np.array(inv, dtype=np.float64).reshape(3, 3)

    def forward(self) -> np.ndarray:
        return self._forward

    def up(self) -> np.ndarray:
        return self._up

    def rotation_matrix(self) -> np.ndarray:
        return self._rotation

    def inverse_rotation_matrix(self) -> np.ndarray:
        return self._inverse
