"""curvis_b200 — the per-pixel null-geodesic renderer of fragarriss/CurVis
(``RelativisticSystem::render_image``, reference src/systems.rs:307-330) as hand-written
sm_100a CUDA kernels behind the C ABI of include/curvis_gpu.h.  This package is the thin host
mirror of the reference's Metric / Camera / SphericalImage / RelativisticSystem surface."""
from ._abi import CurvisError, load_library  # noqa: F401
from .algebra import Orientation  # noqa: F401
from .cameras import Camera  # noqa: F401
from .images import SphericalImage  # noqa: F401
from .metrics import EllisMetric, FlatSphericalMetric, InterstellarMetric  # noqa: F401
from .systems import Context, PeerBuffer, RelativisticSystem  # noqa: F401
