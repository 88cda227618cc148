"""Host mirror of reference src/images.rs ``SphericalImage``: an equirectangular background
plus its orientation.  The texel lookup (``get_pixel_from_vector3``, src/images.rs:107-174)
runs in the render kernel's epilogue; this class owns the RGBA8 texels (what
``DynamicImage::get_pixel`` returns, :107-111) and the orientation's inverse rotation
(:132-142).  Decoding image files is host I/O outside the hot path."""
from __future__ import annotations

import numpy as np

from .algebra import Orientation


class SphericalImage:
    def __init__(self, img, forward=None, up=None):
        """``SphericalImage::new(img, forward, up)`` (src/images.rs:71-89).

        ``img``: uint8 array (H, W, 4) RGBA, (H, W, 3) RGB (alpha 255 is added, as
        ``DynamicImage::get_pixel`` does for RGB8 images) or (H, W) luma.
        """
        a = np.asarray(img)
        if a.dtype != np.uint8:
            raise TypeError("SphericalImage expects uint8 texels")
        if a.ndim == 2:
            a = np.repeat(a[:, :, None], 3, axis=2)
        if a.ndim != 3 or a.shape[2] not in (3, 4):
            raise ValueError("SphericalImage expects (H, W), (H, W, 3) or (H, W, 4)")
        if a.shape[2] == 3:
            a = np.concatenate([a, np.full(a.shape[:2] + (1,), 255, np.uint8)], axis=2)
        self.rgba8 = np.ascontiguousarray(a)
        self.height_pixels, self.width_pixels = int(a.shape[0]), int(a.shape[1])
        self._orientation = Orientation(forward if forward is not None else (1.0, 0.0, 0.0),
                                        up if up is not None else (0.0, 0.0, 1.0))

    def orientation(self) -> Orientation:
        return self._orientation

    def set_forward_up(self, forward, up) -> None:                # src/images.rs:102-104
        self._orientation = Orientation(forward, up)

    def dimensions(self):
        return self.width_pixels, self.height_pixels
