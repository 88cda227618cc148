"""Host mirror of reference src/cameras.rs ``Camera``: position on the metric, orientation in
the local tangent space, sensor geometry.  Per-pixel ray generation
(``outward_vector_on_world_space_from_x_y``, src/cameras.rs:150-172) happens inside the render
kernel; this class only owns what ``Camera::new`` precomputes (src/cameras.rs:79-122)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi


class Camera:
    def __init__(self, position, forward_world, up_world, focal_length, sensor_diagonal,
                 resolution_width, resolution_height):
        if resolution_width < 0 or resolution_height < 0:
            raise _abi.CurvisError(_abi.ERR_INVALID_ARGUMENT, "resolution must be non-negative (u32)")
        self._forward = [float(v) for v in forward_world]
        self._up = [float(v) for v in up_world]
        self._focal_length = float(focal_length)
        self._sensor_diagonal = float(sensor_diagonal)
        self._c = _abi.CurvisCamera()
        self._init(position, resolution_width, resolution_height)

    def _init(self, position, w, h):
        _abi.check(_abi.load_library().curvis_camera_init(
            C.byref(self._c), _abi.dvec(position, 4), _abi.dvec(self._forward, 3), _abi.dvec(self._up, 3),
            self._focal_length, self._sensor_diagonal, int(w), int(h)))

    def as_c(self) -> _abi.CurvisCamera:
        return self._c

    def position(self) -> np.ndarray:
        return np.array(self._c.position, dtype=np.float64)

    def update_position(self, new_position) -> None:          # src/cameras.rs:135-140
        for i in range(4):
            self._c.position[i] = float(new_position[i])

    def update_orientation(self, forward_world, up_world) -> None:   # src/cameras.rs:143-146
        self._forward = [float(v) for v in forward_world]
        self._up = [float(v) for v in up_world]
        self._init(list(self._c.position), self._c.resolution_width, self._c.resolution_height)

    def resolution_width(self) -> int:
        return int(self._c.resolution_width)

    def resolution_height(self) -> int:
        return int(self._c.resolution_height)

    def camera_to_world_rotation_matrix(self) -> np.ndarray:
        return np.array(self._c.cam_to_world, dtype=np.float64).reshape(3, 3)

    def sensor_width(self) -> float:
        return float(self._c.sensor_width)

    def sensor_height(self) -> float:
        return float(self._c.sensor_height)

    def focal_length(self) -> float:
        return float(self._c.focal_length)
