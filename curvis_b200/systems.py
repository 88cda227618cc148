"""Host mirror of reference src/systems.rs ``RelativisticSystem``: the scene (metric, two
backgrounds, camera) and its ``render_image(max_iterations, max_radius, delta)`` entry point
(src/systems.rs:307-330), executed by the sm_100a kernels of libcurvis_b200.so through the C
ABI of include/curvis_gpu.h.  Nothing here computes a pixel on the CPU."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _abi
from .cameras import Camera
from .images import SphericalImage


class Context:
    """Owner of a ``curvis_ctx``: the CUDA devices a frame is row-tiled over."""

    def __init__(self, devices: Optional[Sequence[int]] = None):
        lib = _abi.load_library()
        self._lib = lib
        self._ptr = C.c_void_p()
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            _abi.check(lib.curvis_ctx_create(arr, len(devices), C.byref(self._ptr)))
        else:
            _abi.check(lib.curvis_ctx_create(None, 0, C.byref(self._ptr)))

    @property
    def ptr(self):
        return self._ptr

    def device_count(self) -> int:
        return int(self._lib.curvis_ctx_device_count(self._ptr))

    def close(self):
        if self._ptr:
            self._lib.curvis_ctx_destroy(self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def register_host_buffer(self, array) -> None:
        """curvis_host_register: page-lock a frame buffer (C-contiguous uint8 numpy array) that will be
        passed as `out=` repeatedly; frames are then DMA'd straight into it.  The context keeps a
        reference to the array until unregister_host_buffer / close."""
        _abi.check(self._lib.curvis_host_register(self._ptr, C.c_void_p(array.ctypes.data), array.nbytes), self._ptr)
        self._registered = getattr(self, "_registered", {})
        self._registered[array.ctypes.data] = array

    def unregister_host_buffer(self, array) -> None:
        _abi.check(self._lib.curvis_host_unregister(self._ptr, C.c_void_p(array.ctypes.data)), self._ptr)
        getattr(self, "_registered", {}).pop(array.ctypes.data, None)

    def set_option(self, key: str, value: int) -> None:
        """curvis_ctx_set_option: tuning knobs ("kernel_variant", "blocks_per_sm", "window", "guard", "fast_regs", ...)."""
        _abi.check(self._lib.curvis_ctx_set_option(self._ptr, key.encode(), int(value)), self._ptr)

    def debug_eval(self, op: int, a, b=None):
        """curvis_debug_eval: one device math primitive evaluated elementwise (test hook)."""
        import numpy as np
        a = np.ascontiguousarray(a, dtype=np.float64)
        out = np.empty_like(a)
        dp = C.POINTER(C.c_double)
        bp = None
        if b is not None:
            b = np.ascontiguousarray(b, dtype=np.float64)
            bp = b.ctypes.data_as(dp)
        _abi.check(self._lib.curvis_debug_eval(self._ptr, op, a.ctypes.data_as(dp), bp, out.ctypes.data_as(dp), a.size), self._ptr)
        return out

    def debug_rhs_check(self, metric, n_samples: int, seed: int = 1):
        """curvis_debug_rhs_check: mismatching bits of kernel_variant 4's right-hand side vs the plain operators."""
        bad = (C.c_uint64 * 4)()
        m = metric.as_c()
        _abi.check(self._lib.curvis_debug_rhs_check(self._ptr, C.byref(m), int(n_samples), int(seed), bad), self._ptr)
        return list(bad)

    def last_step_shares(self):
        """curvis_debug_last_step_shares: Euler steps per hardware warp slot (64) and per SM (192) of the last launch with stats."""
        slots, sms = (C.c_uint64 * 64)(), (C.c_uint64 * 192)()
        _abi.check(self._lib.curvis_debug_last_step_shares(self._ptr, slots, sms), self._ptr)
        return list(slots), list(sms)

    def measure_fma_peak(self):
        f64, f32 = C.c_double(), C.c_double()
        _abi.check(self._lib.curvis_measure_fma_peak(self._ptr, C.byref(f64), C.byref(f32)), self._ptr)
        return f64.value, f32.value


class PeerBuffer:
    """A device buffer other ranks' GPUs write into (curvis_peer_buffer_*): ``PeerBuffer.create`` allocates it on
    the context's first device and exposes ``handle`` (64 bytes, CUDA IPC) to send to the peers;
    ``PeerBuffer.open`` maps a peer's buffer from its handle.  ``ptr`` is the device pointer;
    ``as_tensor()`` views it as a torch uint8 tensor (no copy)."""

    def __init__(self, context: "Context", ptr: int, nbytes: int, handle: Optional[bytes], owned: bool):
        self.context, self.ptr, self.nbytes, self.handle, self._owned = context, ptr, nbytes, handle, owned

    @classmethod
    def create(cls, context: "Context", nbytes: int) -> "PeerBuffer":
        ptr, handle = C.c_void_p(), C.create_string_buffer(_abi.IPC_HANDLE_BYTES)
        _abi.check(context._lib.curvis_peer_buffer_create(context.ptr, int(nbytes), C.byref(ptr), handle), context.ptr)
        return cls(context, ptr.value, int(nbytes), handle.raw, True)

    @classmethod
    def open(cls, context: "Context", handle: bytes, nbytes: int) -> "PeerBuffer":
        ptr = C.c_void_p()
        _abi.check(context._lib.curvis_peer_buffer_open(context.ptr, C.create_string_buffer(handle, _abi.IPC_HANDLE_BYTES), C.byref(ptr)), context.ptr)
        return cls(context, ptr.value, int(nbytes), None, False)

    def as_tensor(self, device_index: int = 0):
        import torch

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 3}
        return torch.as_tensor(raw, device=torch.device("cuda", device_index))

    def close(self) -> None:
        if self.ptr:
            fn = self.context._lib.curvis_peer_buffer_destroy if self._owned else self.context._lib.curvis_peer_buffer_close
            _abi.check(fn(self.context.ptr, C.c_void_p(self.ptr)), self.context.ptr)
            self.ptr = 0


class RelativisticSystem:
    """``RelativisticSystem::new(metric, background_positive, background_negative, camera)``
    (src/systems.rs:283-285).  The backgrounds are uploaded to every device of the context once,
    here — like the reference, which decodes them once per system (src/rendering.rs:36-39)."""

    def __init__(self, metric, background_positive: SphericalImage, background_negative: SphericalImage,
                 camera: Camera, context: Optional[Context] = None, devices: Optional[Sequence[int]] = None):
        self.metric = metric
        self.background_positive = background_positive
        self.background_negative = background_negative
        self.camera = camera
        self.context = context if context is not None else Context(devices)
        self._lib = _abi.load_library()
        self.last_stats: Optional[dict] = None
        self._upload(+1, background_positive)
        self._upload(-1, background_negative)

    def _upload(self, side: int, image: SphericalImage) -> None:
        inv = np.ascontiguousarray(image.orientation().inverse_rotation_matrix(), dtype=np.float64)
        _abi.check(self._lib.curvis_set_background(
            self.context.ptr, side, image.rgba8.ctypes.data_as(C.c_void_p), image.width_pixels, image.height_pixels,
            inv.ctypes.data_as(C.POINTER(C.c_double))), self.context.ptr)

    @staticmethod
    def _sim(max_iterations, max_radius, delta, precision=_abi.PRECISION_F64, sampling=_abi.SAMPLING_NEAREST,
             integrator=_abi.INTEGRATOR_EULER, frame=_abi.FRAME_LOCAL, coordinates=_abi.COORDINATES_SPHERICAL,
             step_tolerance=0.0):
        if max_iterations < 0 or max_iterations > 0xFFFFFFFF:
            raise _abi.CurvisError(_abi.ERR_INVALID_ARGUMENT, "max_iterations must fit u32")
        return _abi.CurvisSim(max_iterations=int(max_iterations), max_radius=float(max_radius), delta=float(delta),
                              precision=precision, sampling=sampling, integrator=integrator, frame=frame,
                              coordinates=coordinates, step_tolerance=float(step_tolerance))

    def render_image(self, max_iterations: int, max_radius: float, delta: float, out: Optional[np.ndarray] = None,
                     **options) -> np.ndarray:
        """The whole frame, row-tiled over the context's devices; returns uint8 (H, W, 3) —
        the layout of the ``DynamicImage::ImageRgb8`` the reference returns.  ``out`` reuses a
        caller-owned frame buffer (a fresh 25 MB array per 4K frame costs milliseconds of page
        faults)."""
        cam = self.camera.as_c()
        shape = (cam.resolution_height, cam.resolution_width, 3)
        if out is None:
            out = np.empty(shape, dtype=np.uint8)
        elif out.shape != shape or out.dtype != np.uint8 or not out.flags["C_CONTIGUOUS"]:
            raise ValueError(f"out must be a C-contiguous uint8 array of shape {shape}")
        sim = self._sim(max_iterations, max_radius, delta, **options)
        stats = _abi.CurvisStats()
        m = self.metric.as_c()
        _abi.check(self._lib.curvis_render_image(self.context.ptr, C.byref(m), C.byref(cam), C.byref(sim),
                                                 out.ctypes.data_as(C.c_void_p), C.byref(stats)), self.context.ptr)
        self.last_stats = stats.as_dict()
        return out

    def render_rows(self, max_iterations: int, max_radius: float, delta: float, row_begin: int, row_end: int,
                    with_records: bool = False, out=None, **options):
        """Rows [row_begin, row_end) on the context's first device (one rank's tile)."""
        cam = self.camera.as_c()
        n_rows = max(0, row_end - row_begin)
        if out is None:
            out = np.empty((n_rows, cam.resolution_width, 3), dtype=np.uint8)
        elif out.dtype != np.uint8 or out.shape != (n_rows, cam.resolution_width, 3) or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous uint8 array of shape (rows, W, 3)")
        rec = np.zeros((n_rows, cam.resolution_width), dtype=_abi.RAY_RECORD_DTYPE) if with_records else None
        sim = self._sim(max_iterations, max_radius, delta, **options)
        stats = _abi.CurvisStats()
        m = self.metric.as_c()
        _abi.check(self._lib.curvis_render_rows(
            self.context.ptr, C.byref(m), C.byref(cam), C.byref(sim), int(row_begin), int(row_end),
            out.ctypes.data_as(C.c_void_p), rec.ctypes.data_as(C.c_void_p) if rec is not None else None,
            C.byref(stats)), self.context.ptr)
        self.last_stats = stats.as_dict()
        return (out, rec) if with_records else out

    def render_rows_device(self, max_iterations: int, max_radius: float, delta: float, row_begin: int, row_end: int,
                           out_ptr: int, stream_ptr: int = 0, records_ptr: int = 0, want_stats: bool = False,
                           **options):
        """Device-resident tile: ``out_ptr`` is a device pointer ((row_end-row_begin)*W*3 bytes)
        on the context's first device, ``stream_ptr`` a cudaStream_t.  Asynchronous unless
        ``want_stats``."""
        cam = self.camera.as_c()
        sim = self._sim(max_iterations, max_radius, delta, **options)
        stats = _abi.CurvisStats() if want_stats else None
        m = self.metric.as_c()
        _abi.check(self._lib.curvis_render_rows_device(
            self.context.ptr, C.byref(m), C.byref(cam), C.byref(sim), int(row_begin), int(row_end),
            C.c_void_p(out_ptr), C.c_void_p(records_ptr) if records_ptr else None,
            C.c_void_p(stream_ptr) if stream_ptr else None, C.byref(stats) if stats is not None else None),
            self.context.ptr)
        if stats is not None:
            self.last_stats = stats.as_dict()
            return self.last_stats
        return None

    def render_frames_device(self, cameras, max_iterations: int, max_radius: float, delta: float, row_begin: int, row_end: int,
                             out_ptr: int, stream_ptr: int = 0, want_stats: bool = False, **options):
        """Batched video form (curvis_render_frames_device): rows [row_begin,row_end) of one frame
        per camera in ONE launch; ``out_ptr`` receives len(cameras) tiles, frame-major."""
        arr = (_abi.CurvisCamera * len(cameras))(*[c.as_c() if hasattr(c, "as_c") else c for c in cameras])
        sim = self._sim(max_iterations, max_radius, delta, **options)
        stats = _abi.CurvisStats() if want_stats else None
        m = self.metric.as_c()
        _abi.check(self._lib.curvis_render_frames_device(
            self.context.ptr, C.byref(m), arr, len(cameras), C.byref(sim), int(row_begin), int(row_end),
            C.c_void_p(out_ptr), C.c_void_p(stream_ptr) if stream_ptr else None,
            C.byref(stats) if stats is not None else None), self.context.ptr)
        if stats is not None:
            self.last_stats = stats.as_dict()
            return self.last_stats
        return None

    def render_frames_peers(self, cameras, max_iterations: int, max_radius: float, delta: float, row_begin: int, row_end: int,
                            frame_ptrs, stream_ptr: int = 0, want_stats: bool = False, row_stride: int = 1,
                            block_width: int = 0, **options):
        """Fused render + all-gather (curvis_render_frames_peers): rows row_begin, row_begin + row_stride, ...
        (< row_end) of one frame per camera in ONE launch, every pixel stored into the complete-frames buffer of every peer
        (``frame_ptrs``: device pointers — this rank's PeerBuffer and the opened ones of its peers).  ``block_width`` > 0
        (curvis_render_frames_peers_blocks): the three row arguments count blocks of that many pixels of a row, numbered
        row-major over the frame (distributed.interleaved_blocks)."""
        arr = (_abi.CurvisCamera * len(cameras))(*[c.as_c() if hasattr(c, "as_c") else c for c in cameras])
        ptrs = (C.c_void_p * len(frame_ptrs))(*[C.c_void_p(int(x)) for x in frame_ptrs])
        sim = self._sim(max_iterations, max_radius, delta, **options)
        stats = _abi.CurvisStats() if want_stats else None
        m = self.metric.as_c()
        if block_width:
            _abi.check(self._lib.curvis_render_frames_peers_blocks(
                self.context.ptr, C.byref(m), arr, len(cameras), C.byref(sim), int(row_begin), int(row_end), int(row_stride), int(block_width),
                ptrs, len(frame_ptrs), C.c_void_p(stream_ptr) if stream_ptr else None, C.byref(stats) if stats is not None else None),
                self.context.ptr)
        else:
            _abi.check(self._lib.curvis_render_frames_peers(
                self.context.ptr, C.byref(m), arr, len(cameras), C.byref(sim), int(row_begin), int(row_end), int(row_stride),
                ptrs, len(frame_ptrs), C.c_void_p(stream_ptr) if stream_ptr else None, C.byref(stats) if stats is not None else None),
                self.context.ptr)
        if stats is not None:
            self.last_stats = stats.as_dict()
            return self.last_stats
        return None

    def render_image_efficient(self, max_iterations_propagation: int, max_radius: float, delta: float, alpha_nums: int,
                               max_iterations_sampling: int, sampling_convergence_threshold_1: float,
                               sampling_convergence_threshold_2: float, debug: bool = False, out=None, **options):
        """``render_image_efficient`` (src/systems.rs:333-527), same argument order: the
        table-based renderer the ``curvis`` binary uses.  Returns uint8 (H, W, 3); with
        ``debug`` also a float64 (H, W, 3) array of (alpha, escape angle, escape space).
        ``out``: a caller frame (e.g. one registered with Context.register_host_buffer)."""
        cam = self.camera.as_c()
        if out is None:
            out = np.empty((cam.resolution_height, cam.resolution_width, 3), dtype=np.uint8)
        elif out.dtype != np.uint8 or out.shape != (cam.resolution_height, cam.resolution_width, 3) or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous uint8 array of shape (H, W, 3)")
        dbg = np.empty((cam.resolution_height, cam.resolution_width, 3), dtype=np.float64) if debug else None
        sim = self._sim(max_iterations_propagation, max_radius, delta, **options)
        smp = _abi.CurvisSamplingSettings(alphas_num=int(alpha_nums), max_iterations_sampling=int(max_iterations_sampling),
                                          threshold_1=float(sampling_convergence_threshold_1),
                                          threshold_2=float(sampling_convergence_threshold_2))
        stats, info = _abi.CurvisStats(), _abi.CurvisEfficientInfo()
        m = self.metric.as_c()
        _abi.check(self._lib.curvis_render_image_efficient(
            self.context.ptr, C.byref(m), C.byref(cam), C.byref(sim), C.byref(smp), out.ctypes.data_as(C.c_void_p),
            dbg.ctypes.data_as(C.POINTER(C.c_double)) if dbg is not None else None, C.byref(stats), C.byref(info)), self.context.ptr)
        self.last_stats = stats.as_dict()
        self.last_efficient_info = info.as_dict()
        return (out, dbg) if debug else out

    def render_rows_rgba32f(self, max_iterations: int, max_radius: float, delta: float, row_begin: int, row_end: int, **options):
        """Unrounded colours (curvis_render_rows_rgba32f): float32 (rows, W, 4) on the 0..255 scale."""
        cam = self.camera.as_c()
        out = np.empty((max(0, row_end - row_begin), cam.resolution_width, 4), dtype=np.float32)
        sim = self._sim(max_iterations, max_radius, delta, **options)
        stats = _abi.CurvisStats()
        m = self.metric.as_c()
        _abi.check(self._lib.curvis_render_rows_rgba32f(self.context.ptr, C.byref(m), C.byref(cam), C.byref(sim), int(row_begin),
                                                        int(row_end), out.ctypes.data_as(C.c_void_p), C.byref(stats)), self.context.ptr)
        self.last_stats = stats.as_dict()
        return out

    def debug_bilinear(self, side: int, fx, fy):
        """curvis_debug_bilinear: the fp32 tap at explicit continuous texel coordinates."""
        fx = np.ascontiguousarray(fx, dtype=np.float64)
        fy = np.ascontiguousarray(fy, dtype=np.float64)
        out = np.empty(fx.shape + (4,), dtype=np.float32)
        dp = C.POINTER(C.c_double)
        _abi.check(self._lib.curvis_debug_bilinear(self.context.ptr, side, fx.ctypes.data_as(dp), fy.ctypes.data_as(dp),
                                                   out.ctypes.data_as(C.c_void_p), fx.size), self.context.ptr)
        return out
