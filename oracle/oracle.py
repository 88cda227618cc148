"""ctypes wrapper of oracle/libcurvis_oracle.so — the CPU restatement of the reference's
``render_image`` path (oracle/curvis_oracle.c).  TEST INFRASTRUCTURE: imported only by tests/,
bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke().  It shares the
plain-data struct layouts of include/curvis_gpu.h (via curvis_b200._abi) and nothing else with
the product."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from curvis_b200 import _abi

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "libcurvis_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_DIR, "curvis_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        dp, vp = C.POINTER(C.c_double), C.c_void_p
        L.oracle_orientation.argtypes = [dp, dp, dp, dp, dp]
        L.oracle_camera_init.argtypes = [C.POINTER(_abi.CurvisCamera), dp, dp, dp, C.c_double, C.c_double, C.c_uint32, C.c_uint32]
        L.oracle_normalize_theta_phi.argtypes = [C.c_double, C.c_double, dp, dp]
        L.oracle_normalize_theta_phi.restype = None
        L.oracle_vector3_from_theta_phi.argtypes = [C.c_double, C.c_double, dp]
        L.oracle_vector3_from_theta_phi.restype = None
        L.oracle_theta_phi_from_vector3.argtypes = [dp, dp, dp]
        L.oracle_theta_phi_from_vector3.restype = None
        L.oracle_outward_vector_on_camera_space.argtypes = [C.POINTER(_abi.CurvisCamera), C.c_uint32, C.c_uint32, dp]
        L.oracle_outward_vector_on_camera_space.restype = None
        L.oracle_outward_vector_on_world_space.argtypes = [C.POINTER(_abi.CurvisCamera), C.c_uint32, C.c_uint32, dp]
        L.oracle_outward_vector_on_world_space.restype = None
        L.oracle_new_photon.argtypes = [C.POINTER(_abi.CurvisMetric), dp, dp, dp]
        L.oracle_new_photon.restype = None
        L.oracle_step.argtypes = [C.POINTER(_abi.CurvisMetric), dp, C.c_double]
        L.oracle_step.restype = None
        L.oracle_step_rk4.argtypes = [C.POINTER(_abi.CurvisMetric), dp, C.c_double]
        L.oracle_step_rk4.restype = None
        L.oracle_escape_photon.argtypes = [C.POINTER(_abi.CurvisMetric), dp, C.c_double, C.c_uint32, C.c_double, C.POINTER(C.c_uint32)]
        L.oracle_relativistic_vector_to_direction.argtypes = [C.POINTER(_abi.CurvisMetric), dp, dp, dp]
        L.oracle_relativistic_vector_to_direction.restype = None
        L.oracle_lookup_direction.argtypes = [C.POINTER(_abi.CurvisMetric), dp, C.c_int, dp]
        L.oracle_escape_photon_sim.argtypes = [C.POINTER(_abi.CurvisMetric), dp, C.POINTER(_abi.CurvisSim), C.POINTER(C.c_uint32), dp]
        L.oracle_step_adaptive.argtypes = [C.POINTER(_abi.CurvisMetric), dp, C.c_double, C.c_double]
        L.oracle_step_adaptive.restype = None
        L.oracle_squared_norm_cov.argtypes = [C.POINTER(_abi.CurvisMetric), dp, dp]
        L.oracle_squared_norm_cov.restype = C.c_double
        L.oracle_texel_from_vector3.argtypes = [dp, dp, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), dp, dp]
        L.oracle_texel_from_vector3.restype = None
        L.oracle_render_rows.argtypes = [
            C.POINTER(_abi.CurvisMetric), C.POINTER(_abi.CurvisCamera), C.POINTER(_abi.CurvisSim),
            vp, C.c_uint32, C.c_uint32, dp, vp, C.c_uint32, C.c_uint32, dp,
            C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, C.POINTER(_abi.CurvisStats), C.c_int]
        L.oracle_render_rows_ex.argtypes = L.oracle_render_rows.argtypes + [vp]
        L.oracle_bilinear_tap.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.POINTER(C.c_float)]
        L.oracle_bilinear_tap.restype = None
        L.oracle_trajectory.argtypes = [C.POINTER(_abi.CurvisMetric), dp, dp, C.c_double, C.c_uint32, dp]
        L.oracle_trajectory.restype = None
        L.oracle_metric_validate.argtypes = [C.POINTER(_abi.CurvisMetric)]
        L.oracle_rotation_from_two_vectors.argtypes = [dp, dp, dp]
        L.oracle_rotation_matrix_from_theta_phi.argtypes = [C.c_double, C.c_double, dp]
        L.oracle_rotation_matrix_from_theta_phi.restype = None
        L.oracle_compute_escape_angle.argtypes = [C.POINTER(_abi.CurvisMetric), C.c_double, C.c_double, C.c_double, C.c_uint32, C.c_double,
                                                  dp, C.POINTER(C.c_uint32)]
        pdp = C.POINTER(dp)
        L.oracle_sample_escape_angles.argtypes = [C.POINTER(_abi.CurvisMetric), C.c_double, C.c_double, C.c_uint32, C.c_double,
                                                  C.c_double, C.c_double, C.c_uint32, C.c_uint32, C.c_double, C.c_double,
                                                  pdp, pdp, pdp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                                  C.POINTER(C.c_uint32)]
        L.oracle_free.argtypes = [vp]
        L.oracle_free.restype = None
        L.oracle_interp_slice.argtypes = [dp, dp, C.c_uint32, dp, C.c_size_t, dp]
        L.oracle_interp_slice.restype = None
        L.oracle_render_image_efficient.argtypes = [
            C.POINTER(_abi.CurvisMetric), C.POINTER(_abi.CurvisCamera), C.POINTER(_abi.CurvisSim),
            C.c_uint32, C.c_uint32, C.c_double, C.c_double,
            vp, C.c_uint32, C.c_uint32, dp, vp, C.c_uint32, C.c_uint32, dp,
            C.c_uint32, C.c_uint32, vp, dp, dp, dp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def _d(values, n):
    return (C.c_double * n)(*[float(v) for v in values])


def metric(kind: str, rho=1.0, m=0.1, a=1e-4) -> _abi.CurvisMetric:
    k = {"ellis": _abi.METRIC_ELLIS, "interstellar": _abi.METRIC_INTERSTELLAR, "flat": _abi.METRIC_FLAT}[kind.lower()]
    return _abi.CurvisMetric(kind=k, rho=rho, m=m, a=a)


def orientation(forward, up):
    rot, inv, upo = (C.c_double * 9)(), (C.c_double * 9)(), (C.c_double * 3)()
    rc = lib().oracle_orientation(_d(forward, 3), _d(up, 3), rot, inv, upo)
    if rc:
        raise ValueError("Forward and up vectors must not be parallel")
    return np.array(rot).reshape(3, 3), np.array(inv).reshape(3, 3), np.array(upo)


def camera(position, forward, up, focal_length, diagonal, width, height) -> _abi.CurvisCamera:
    cam = _abi.CurvisCamera()
    rc = lib().oracle_camera_init(C.byref(cam), _d(position, 4), _d(forward, 3), _d(up, 3), focal_length, diagonal, width, height)
    if rc:
        raise ValueError(f"oracle_camera_init failed ({rc})")
    return cam


def sim(max_iterations, max_radius, delta, sampling=0, integrator=0, frame=0, coordinates=0, step_tolerance=0.0) -> _abi.CurvisSim:
    return _abi.CurvisSim(max_iterations=max_iterations, max_radius=max_radius, delta=delta, precision=0, sampling=sampling,
                          integrator=integrator, frame=frame, coordinates=coordinates, step_tolerance=step_tolerance)


def bilinear_tap(bg_rgba8, fx, fy):
    """The fp32 bilinear tap (extension) at continuous texel coordinates: float32 (..., 4)."""
    bg = np.ascontiguousarray(bg_rgba8, dtype=np.uint8)
    fx = np.asarray(fx, dtype=np.float64); fy = np.asarray(fy, dtype=np.float64)
    out = np.empty(fx.shape + (4,), dtype=np.float32)
    flat = out.reshape(-1, 4)
    o = (C.c_float * 4)()
    for k, (x, y) in enumerate(zip(fx.reshape(-1), fy.reshape(-1))):
        lib().oracle_bilinear_tap(bg.ctypes.data_as(C.c_void_p), bg.shape[1], bg.shape[0], float(x), float(y), o)
        flat[k] = o[:]
    return out


def normalize_theta_phi(theta, phi):
    t, p = C.c_double(), C.c_double()
    lib().oracle_normalize_theta_phi(theta, phi, C.byref(t), C.byref(p))
    return t.value, p.value


def vector3_from_theta_phi(theta, phi):
    o = (C.c_double * 3)()
    lib().oracle_vector3_from_theta_phi(theta, phi, o)
    return np.array(o)


def theta_phi_from_vector3(v):
    t, p = C.c_double(), C.c_double()
    lib().oracle_theta_phi_from_vector3(_d(v, 3), C.byref(t), C.byref(p))
    return t.value, p.value


def outward_vector(cam, px, py, world=True):
    o = (C.c_double * 3)()
    fn = lib().oracle_outward_vector_on_world_space if world else lib().oracle_outward_vector_on_camera_space
    fn(C.byref(cam), px, py, o)
    return np.array(o)


def new_photon(g, position, direction):
    ph = (C.c_double * 8)()
    lib().oracle_new_photon(C.byref(g), _d(position, 4), _d(direction, 3), ph)
    a = np.array(ph)
    return a[:4].copy(), a[4:].copy()


def step(g, x, p, delta, rk4=False):
    ph = _d(list(x) + list(p), 8)
    (lib().oracle_step_rk4 if rk4 else lib().oracle_step)(C.byref(g), ph, delta)
    a = np.array(ph)
    return a[:4].copy(), a[4:].copy()


def escape_photon(g, x, p, delta, max_iterations, max_radius):
    ph = _d(list(x) + list(p), 8)
    steps = C.c_uint32()
    side = lib().oracle_escape_photon(C.byref(g), ph, delta, max_iterations, max_radius, C.byref(steps))
    a = np.array(ph)
    return side, steps.value, a[:4].copy(), a[4:].copy()


def lookup_direction(g, x, p, frame=0):
    """oracle_lookup_direction: the vector that indexes the background for an escaped photon (x, p) in a curvis_frame.
    Raises where the reference panics (rotation_from_two_vectors on parallel vectors)."""
    ph = _d(list(x) + list(p), 8)
    o = (C.c_double * 3)()
    if lib().oracle_lookup_direction(C.byref(g), ph, frame, o):
        raise ValueError("v1 and v2 must not be parallel")
    return np.array(o)


def escape_photon_sim(g, x, p, s, track=False):
    """escape_photon under a full curvis_sim (integrator extensions): (side, steps, x, p[, (min |sin theta|, stiffness)])."""
    ph = _d(list(x) + list(p), 8)
    steps = C.c_uint32()
    diag = (C.c_double * 2)()
    side = lib().oracle_escape_photon_sim(C.byref(g), ph, C.byref(s), C.byref(steps), diag if track else None)
    a = np.array(ph)
    out = (side, steps.value, a[:4].copy(), a[4:].copy())
    return out + ((diag[0], diag[1]),) if track else out


def direction(g, p, x):
    o = (C.c_double * 3)()
    lib().oracle_relativistic_vector_to_direction(C.byref(g), _d(p, 4), _d(x, 4), o)
    return np.array(o)


def squared_norm_cov(g, p, x):
    return lib().oracle_squared_norm_cov(C.byref(g), _d(p, 4), _d(x, 4))


def texel_from_vector3(v, bg_w, bg_h, inv_rot=None):
    x, y, t, p = C.c_uint32(), C.c_uint32(), C.c_double(), C.c_double()
    inv = _d(np.asarray(inv_rot if inv_rot is not None else np.eye(3)).reshape(-1), 9)
    lib().oracle_texel_from_vector3(inv, _d(v, 3), bg_w, bg_h, C.byref(x), C.byref(y), C.byref(t), C.byref(p))
    return x.value, y.value, t.value, p.value


def trajectory(g, position, direction_, delta, n):
    out = np.zeros((n, 8), dtype=np.float64)
    lib().oracle_trajectory(C.byref(g), _d(position, 4), _d(direction_, 3), delta, n, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def render_rows(g, cam, s, bg_pos, bg_neg, row_begin=0, row_end=None, row_stride=1, threads=1,
                pos_inv_rot=None, neg_inv_rot=None, with_records=True, rgba32f=None):
    """Returns (rgb8 (rows, W, 3), records (rows, W) structured, stats dict).  ``rgba32f``: an
    optional float32 (rows, W, 4) array that receives the unrounded colours."""
    W, H = cam.resolution_width, cam.resolution_height
    if row_end is None:
        row_end = H
    n_rows = 0 if row_end <= row_begin else (row_end - row_begin + row_stride - 1) // row_stride
    bg_pos = np.ascontiguousarray(bg_pos, dtype=np.uint8)
    bg_neg = np.ascontiguousarray(bg_neg, dtype=np.uint8)
    out = np.zeros((n_rows, W, 3), dtype=np.uint8)
    rec = np.zeros((n_rows, W), dtype=_abi.RAY_RECORD_DTYPE) if with_records else None
    st = _abi.CurvisStats()
    dp = C.POINTER(C.c_double)
    pi = np.ascontiguousarray(pos_inv_rot, dtype=np.float64).ctypes.data_as(dp) if pos_inv_rot is not None else None
    ni = np.ascontiguousarray(neg_inv_rot, dtype=np.float64).ctypes.data_as(dp) if neg_inv_rot is not None else None
    rc = lib().oracle_render_rows_ex(
        C.byref(g), C.byref(cam), C.byref(s),
        bg_pos.ctypes.data_as(C.c_void_p), bg_pos.shape[1], bg_pos.shape[0], pi,
        bg_neg.ctypes.data_as(C.c_void_p), bg_neg.shape[1], bg_neg.shape[0], ni,
        row_begin, row_end, row_stride, out.ctypes.data_as(C.c_void_p),
        rec.ctypes.data_as(C.c_void_p) if rec is not None else None, C.byref(st), threads,
        rgba32f.ctypes.data_as(C.c_void_p) if rgba32f is not None else None)
    if rc:
        raise RuntimeError(f"oracle_render_rows status {rc}")
    return out, rec, st.as_dict()


def rotation_from_two_vectors(v1, v2):
    m = (C.c_double * 9)()
    if lib().oracle_rotation_from_two_vectors(_d(v1, 3), _d(v2, 3), m):
        raise ValueError("v1 and v2 must not be parallel")
    return np.array(m).reshape(3, 3)


def rotation_matrix_from_theta_phi(theta, phi):
    m = (C.c_double * 9)()
    lib().oracle_rotation_matrix_from_theta_phi(theta, phi, m)
    return np.array(m).reshape(3, 3)


def compute_escape_angle(g, l, alpha, delta, max_iterations, max_radius):
    """(side, angle, steps): side +1/-1, 0 NotEscaped (angle NaN)."""
    a, st = C.c_double(), C.c_uint32()
    side = lib().oracle_compute_escape_angle(C.byref(g), l, alpha, delta, max_iterations, max_radius, C.byref(a), C.byref(st))
    return side, a.value, st.value


def sample_escape_angles(g, l, delta, max_iterations, max_radius, a_min=-0.1 * np.pi, a_max=1.1 * np.pi, initial_points=100,
                         max_sampling_iterations=100, thr1=1e-5, thr2=1e-5):
    """doubly_sample_function over compute_escape_angle: (alphas, escapes, signs, info)."""
    dp = C.POINTER(C.c_double)
    pa, pe, ps = dp(), dp(), dp()
    n, ev, stp, passes = C.c_uint32(), C.c_uint64(), C.c_uint64(), C.c_uint32()
    rc = lib().oracle_sample_escape_angles(C.byref(g), l, delta, max_iterations, max_radius, a_min, a_max, initial_points,
                                           max_sampling_iterations, thr1, thr2, C.byref(pa), C.byref(pe), C.byref(ps), C.byref(n),
                                           C.byref(ev), C.byref(stp), C.byref(passes))
    out = tuple(np.ctypeslib.as_array(p, shape=(max(n.value, 1),))[: n.value].copy() for p in (pa, pe, ps))
    for p in (pa, pe, ps):
        lib().oracle_free(p)
    if rc:
        raise RuntimeError("the reference would panic while sampling")
    return out + (dict(points=n.value, evaluations=ev.value, steps=stp.value, passes=passes.value),)


def interp_slice(x, y, xp):
    x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    xp = np.ascontiguousarray(xp, dtype=np.float64)
    out = np.empty_like(xp)
    dp = C.POINTER(C.c_double)
    lib().oracle_interp_slice(x.ctypes.data_as(dp), y.ctypes.data_as(dp), min(len(x), len(y)), xp.ctypes.data_as(dp), xp.size,
                              out.ctypes.data_as(dp))
    return out


def render_image_efficient(g, cam, s, bg_pos, bg_neg, alpha_nums=100, max_iterations_sampling=100, thr1=1e-5, thr2=1e-5,
                           row_begin=0, row_end=None, pos_inv_rot=None, neg_inv_rot=None, debug=False):
    """render_image_efficient restated: (rgb8 (rows, W, 3), info[, alpha, angle, space])."""
    W, H = cam.resolution_width, cam.resolution_height
    if row_end is None:
        row_end = H
    rows = max(0, row_end - row_begin)
    bg_pos = np.ascontiguousarray(bg_pos, dtype=np.uint8); bg_neg = np.ascontiguousarray(bg_neg, dtype=np.uint8)
    out = np.zeros((rows, W, 3), dtype=np.uint8)
    dp = C.POINTER(C.c_double)
    dbg = [np.zeros((rows, W)) for _ in range(3)] if debug else [None] * 3
    pi = np.ascontiguousarray(pos_inv_rot, dtype=np.float64).ctypes.data_as(dp) if pos_inv_rot is not None else None
    ni = np.ascontiguousarray(neg_inv_rot, dtype=np.float64).ctypes.data_as(dp) if neg_inv_rot is not None else None
    tp, te, ts = C.c_uint32(), C.c_uint64(), C.c_uint64()
    rc = lib().oracle_render_image_efficient(
        C.byref(g), C.byref(cam), C.byref(s), alpha_nums, max_iterations_sampling, thr1, thr2,
        bg_pos.ctypes.data_as(C.c_void_p), bg_pos.shape[1], bg_pos.shape[0], pi,
        bg_neg.ctypes.data_as(C.c_void_p), bg_neg.shape[1], bg_neg.shape[0], ni,
        row_begin, row_end, out.ctypes.data_as(C.c_void_p),
        *[d.ctypes.data_as(dp) if d is not None else None for d in dbg], C.byref(tp), C.byref(te), C.byref(ts))
    if rc:
        raise RuntimeError(f"oracle_render_image_efficient status {rc}")
    info = dict(table_points=tp.value, table_evaluations=te.value, table_steps=ts.value)
    return (out, info, *dbg) if debug else (out, info)
