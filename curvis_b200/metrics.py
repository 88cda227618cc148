"""Host mirror of the metric structs of reference src/metrics.rs.  A metric crosses the C ABI
as a closed enum + parameters (include/curvis_gpu.h ``curvis_metric``); the shape functions
r(l), r_squared(l), r_derivative(l) (src/metrics.rs:42-44) live in the CUDA kernels, one
template instantiation per kind."""
from __future__ import annotations

import ctypes as C

from . import _abi


class _Metric:
    kind = -1

    def __init__(self, rho=0.0, m=0.0, a=0.0):
        self._c = _abi.CurvisMetric(kind=self.kind, rho=float(rho), m=float(m), a=float(a))
        # constructors panic on non-positive parameters (src/metrics.rs:407-409, :443-456)
        _abi.check(_abi.load_library().curvis_metric_validate(C.byref(self._c)))

    def as_c(self) -> _abi.CurvisMetric:
        return self._c


class EllisMetric(_Metric):
    """``EllisMetric::new(rho)`` (src/metrics.rs:404-414)."""
    kind = _abi.METRIC_ELLIS

    def __init__(self, rho: float):
        super().__init__(rho=rho)
        self.rho = float(rho)


class InterstellarMetric(_Metric):
    """``InterstellarMetric::new(m, a, rho)`` (src/metrics.rs:441-459)."""
    kind = _abi.METRIC_INTERSTELLAR

    def __init__(self, m: float, a: float, rho: float):
        super().__init__(rho=rho, m=m, a=a)
        self.m, self.a, self.rho = float(m), float(a), float(rho)


class FlatSphericalMetric(_Metric):
    """``FlatSphericalMetric::new()`` (src/metrics.rs:496-498)."""
    kind = _abi.METRIC_FLAT

    def __init__(self):
        super().__init__()
