"""ctypes mirror of include/curvis_gpu.h and the loader of libcurvis_b200.so.

The library is the product: there is no Python/CPU compute path behind it.  Loading fails
loudly when the shared object has not been built (``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C curvis_b200/csrc``), and creating a context fails loudly when no
sm_100 device is visible.
"""
from __future__ import annotations

import ctypes as C
import os

ABI_VERSION = 4

# curvis_status
OK = 0
ERR_INVALID_ARGUMENT = 1
ERR_CAMERA_OUTSIDE_RADIUS = 2
ERR_PARALLEL_VECTORS = 3
ERR_INVALID_METRIC = 4
ERR_NO_BACKGROUND = 5
ERR_CUDA = 6
ERR_OUT_OF_MEMORY = 7
ERR_NO_DEVICE = 8
ERR_UNSUPPORTED = 9

METRIC_ELLIS, METRIC_INTERSTELLAR, METRIC_FLAT = 0, 1, 2
PRECISION_F64, PRECISION_F32, PRECISION_F64_FAST = 0, 1, 2
SAMPLING_NEAREST, SAMPLING_BILINEAR = 0, 1
INTEGRATOR_EULER, INTEGRATOR_RK4, INTEGRATOR_EULER_ADAPTIVE = 0, 1, 2
FRAME_LOCAL, FRAME_WORLD, FRAME_WORLD_QUIRK = 0, 1, 2
COORDINATES_SPHERICAL, COORDINATES_CARTESIAN = 0, 1


class CurvisMetric(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("rho", C.c_double), ("m", C.c_double), ("a", C.c_double)]


class CurvisCamera(C.Structure):
    _fields_ = [
        ("position", C.c_double * 4),
        ("cam_to_world", C.c_double * 9),
        ("focal_length", C.c_double),
        ("sensor_width", C.c_double),
        ("sensor_height", C.c_double),
        ("resolution_width", C.c_uint32),
        ("resolution_height", C.c_uint32),
    ]


class CurvisSim(C.Structure):
    _fields_ = [
        ("max_iterations", C.c_uint32),
        ("frame", C.c_int32),
        ("max_radius", C.c_double),
        ("delta", C.c_double),
        ("precision", C.c_int32),
        ("sampling", C.c_int32),
        ("integrator", C.c_int32),
        ("coordinates", C.c_int32),
        ("step_tolerance", C.c_double),
        ("_reserved", C.c_double),
    ]


class CurvisStats(C.Structure):
    _fields_ = [
        ("total_steps", C.c_uint64),
        ("n_rays", C.c_uint64),
        ("n_positive", C.c_uint64),
        ("n_negative", C.c_uint64),
        ("n_not_escaped", C.c_uint64),
        ("n_clamped", C.c_uint64),
        ("n_reintegrated", C.c_uint64),
        ("n_kicked", C.c_uint64),
        ("kernel_ms", C.c_double),
        ("total_ms", C.c_double),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class CurvisSamplingSettings(C.Structure):
    _fields_ = [("alphas_num", C.c_uint32), ("max_iterations_sampling", C.c_uint32),
                ("threshold_1", C.c_double), ("threshold_2", C.c_double)]


class CurvisEfficientInfo(C.Structure):
    _fields_ = [("table_points", C.c_uint32), ("table_passes", C.c_uint32), ("table_evaluations", C.c_uint64),
                ("table_steps", C.c_uint64), ("table_ms", C.c_double), ("pixels_ms", C.c_double)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class CurvisRayRecord(C.Structure):
    _fields_ = [
        ("l", C.c_double), ("theta", C.c_double), ("phi", C.c_double),
        ("p_l", C.c_double), ("p_theta", C.c_double), ("p_phi", C.c_double),
        ("steps", C.c_uint32), ("side", C.c_int32), ("texel_x", C.c_uint32), ("texel_y", C.c_uint32),
        ("min_abs_sin_theta", C.c_double), ("stiffness", C.c_double),
    ]


# numpy view of curvis_ray_record (same layout, 80 bytes)
RAY_RECORD_DTYPE = [
    ("l", "<f8"), ("theta", "<f8"), ("phi", "<f8"), ("p_l", "<f8"), ("p_theta", "<f8"), ("p_phi", "<f8"),
    ("steps", "<u4"), ("side", "<i4"), ("texel_x", "<u4"), ("texel_y", "<u4"),
    ("min_abs_sin_theta", "<f8"), ("stiffness", "<f8"),
]

MAX_PEERS = 8               # CURVIS_MAX_PEERS
IPC_HANDLE_BYTES = 64       # CURVIS_IPC_HANDLE_BYTES

# Every symbol include/curvis_gpu.h declares (tests assert the .so exports exactly these).
EXPORTED_SYMBOLS = (
    "curvis_ctx_create", "curvis_ctx_destroy", "curvis_last_error", "curvis_abi_version",
    "curvis_ctx_device_count", "curvis_orientation", "curvis_camera_init", "curvis_metric_validate",
    "curvis_set_background", "curvis_render_image", "curvis_render_rows", "curvis_render_rows_device",
    "curvis_measure_fma_peak", "curvis_kernel_launch_count", "curvis_ctx_set_option", "curvis_debug_eval",
    "curvis_render_frames_device", "curvis_render_image_efficient", "curvis_render_rows_rgba32f", "curvis_debug_bilinear",
    "curvis_debug_shape_table_host", "curvis_host_register", "curvis_host_unregister",
    "curvis_peer_buffer_create", "curvis_peer_buffer_open", "curvis_peer_buffer_close", "curvis_peer_buffer_destroy",
    "curvis_render_frames_peers", "curvis_render_frames_peers_blocks", "curvis_debug_last_step_shares", "curvis_debug_rhs_check", "curvis_debug_inverse_table_host", "curvis_debug_inverse_shape",
    "curvis_debug_fn_table_host",
)

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libcurvis_b200.so")
_lib = None


class CurvisError(RuntimeError):
    """A non-zero curvis_status; ``code`` holds it (the reference panics or returns Err(String))."""

    def __init__(self, code: int, message: str):
        super().__init__(f"curvis status {code}: {message}")
        self.code = code
        self.message = message


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C curvis_b200/csrc` (or __graft_entry__.build()). "
            "curvis_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    dp, vp = C.POINTER(C.c_double), C.c_void_p
    lib.curvis_ctx_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    lib.curvis_ctx_destroy.argtypes = [vp]
    lib.curvis_ctx_destroy.restype = None
    lib.curvis_last_error.argtypes = [vp]
    lib.curvis_last_error.restype = C.c_char_p
    lib.curvis_abi_version.argtypes = []
    lib.curvis_ctx_device_count.argtypes = [vp]
    lib.curvis_orientation.argtypes = [dp, dp, dp, dp, dp]
    lib.curvis_camera_init.argtypes = [C.POINTER(CurvisCamera), dp, dp, dp, C.c_double, C.c_double, C.c_uint32, C.c_uint32]
    lib.curvis_metric_validate.argtypes = [C.POINTER(CurvisMetric)]
    lib.curvis_set_background.argtypes = [vp, C.c_int, vp, C.c_uint32, C.c_uint32, dp]
    lib.curvis_render_image.argtypes = [vp, C.POINTER(CurvisMetric), C.POINTER(CurvisCamera), C.POINTER(CurvisSim), vp, C.POINTER(CurvisStats)]
    lib.curvis_render_rows.argtypes = [vp, C.POINTER(CurvisMetric), C.POINTER(CurvisCamera), C.POINTER(CurvisSim),
                                       C.c_uint32, C.c_uint32, vp, vp, C.POINTER(CurvisStats)]
    lib.curvis_render_rows_device.argtypes = [vp, C.POINTER(CurvisMetric), C.POINTER(CurvisCamera), C.POINTER(CurvisSim),
                                              C.c_uint32, C.c_uint32, vp, vp, vp, C.POINTER(CurvisStats)]
    lib.curvis_render_frames_device.argtypes = [vp, C.POINTER(CurvisMetric), C.POINTER(CurvisCamera), C.c_uint32, C.POINTER(CurvisSim),
                                                C.c_uint32, C.c_uint32, vp, vp, C.POINTER(CurvisStats)]
    lib.curvis_render_image_efficient.argtypes = [vp, C.POINTER(CurvisMetric), C.POINTER(CurvisCamera), C.POINTER(CurvisSim),
                                                  C.POINTER(CurvisSamplingSettings), vp, dp, C.POINTER(CurvisStats),
                                                  C.POINTER(CurvisEfficientInfo)]
    lib.curvis_render_rows_rgba32f.argtypes = [vp, C.POINTER(CurvisMetric), C.POINTER(CurvisCamera), C.POINTER(CurvisSim),
                                               C.c_uint32, C.c_uint32, vp, C.POINTER(CurvisStats)]
    lib.curvis_debug_bilinear.argtypes = [vp, C.c_int, dp, dp, vp, C.c_size_t]
    lib.curvis_measure_fma_peak.argtypes = [vp, dp, dp]
    lib.curvis_ctx_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    lib.curvis_debug_eval.argtypes = [vp, C.c_int, dp, dp, dp, C.c_size_t]
    lib.curvis_peer_buffer_create.argtypes = [vp, C.c_size_t, C.POINTER(vp), C.c_char_p]
    lib.curvis_peer_buffer_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    lib.curvis_peer_buffer_close.argtypes = [vp, vp]
    lib.curvis_peer_buffer_destroy.argtypes = [vp, vp]
    lib.curvis_render_frames_peers.argtypes = [vp, C.POINTER(CurvisMetric), C.POINTER(CurvisCamera), C.c_uint32, C.POINTER(CurvisSim),
                                               C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp), C.c_uint32, vp, C.POINTER(CurvisStats)]
    lib.curvis_render_frames_peers_blocks.argtypes = [vp, C.POINTER(CurvisMetric), C.POINTER(CurvisCamera), C.c_uint32, C.POINTER(CurvisSim),
                                                      C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp), C.c_uint32, vp,
                                                      C.POINTER(CurvisStats)]
    lib.curvis_debug_last_step_shares.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.curvis_host_register.argtypes = [vp, vp, C.c_size_t]
    lib.curvis_host_unregister.argtypes = [vp, vp]
    lib.curvis_debug_shape_table_host.argtypes = [dp, dp, dp, C.c_size_t]
    lib.curvis_debug_inverse_table_host.argtypes = [C.c_double, C.c_double, dp, dp, dp, C.c_size_t]
    lib.curvis_debug_fn_table_host.argtypes = [C.c_int, dp, dp, C.c_size_t]
    lib.curvis_debug_inverse_shape.argtypes = [vp, C.POINTER(CurvisMetric), dp, dp, dp, C.c_size_t]
    lib.curvis_debug_rhs_check.argtypes = [vp, C.POINTER(CurvisMetric), C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("curvis_ctx_destroy", "curvis_last_error", "curvis_kernel_launch_count"):
            fn.restype = C.c_int
    lib.curvis_kernel_launch_count.argtypes = []
    lib.curvis_kernel_launch_count.restype = C.c_uint64
    if lib.curvis_abi_version() != ABI_VERSION:
        raise ImportError(f"libcurvis_b200.so ABI {lib.curvis_abi_version()} != expected {ABI_VERSION}")
    _lib = lib
    return lib


def check(code: int, ctx=None) -> None:
    if code != OK:
        msg = load_library().curvis_last_error(ctx)
        raise CurvisError(code, msg.decode("utf-8", "replace") if msg else "")


def dvec(values, n):
    arr = (C.c_double * n)(*[float(v) for v in values])
    return arr
