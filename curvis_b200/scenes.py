"""Default scene values (reference settings/defaults/*.toml) and the synthetic, *decodable*
equirectangular backgrounds used by bench.py and the parity tests: the
texel index can be read back from its colour, so a wrong lookup is visible in the frame.
The repo ships no backgrounds (the reference's user supplies them, README.md:31-33)."""
from __future__ import annotations

import math

import numpy as np

# settings/defaults/image_settings.toml, camera_settings.toml, simulation_settings.toml,
# ellis_metric_settings.toml, interstellar_metric_settings.toml
DEFAULT_CAMERA_POSITION = (0.0, 5.0, math.pi / 2.0, 0.0)
DEFAULT_FORWARD = (-1.0, 0.0, 0.0)
DEFAULT_UP = (0.0, 0.0, 1.0)
DEFAULT_DIAGONAL = 43.0
DEFAULT_FOCAL_LENGTH = 15.0
DEFAULT_RESOLUTION = (960, 540)
DEFAULT_ESCAPE_RADIUS = 100.0
DEFAULT_MAX_ITERATIONS = 40000          # key `ray_integration_max_itarations` (sic)
DEFAULT_STEP = 0.05
DEFAULT_ELLIS = {"rho": 1.0}
DEFAULT_INTERSTELLAR = {"m": 0.1, "a": 1e-4, "rho": 1.0}


def decodable_background(width: int = 8192, height: int = 4096, negative: bool = False) -> np.ndarray:
    """RGBA8 (H, W, 4): R = x & 255, G = y & 255, B = ((x >> 8) << 4) | (y >> 8), A = 255.
    ``negative`` bit-inverts R and G so the two sides of the wormhole are distinguishable.
    Decodable for width <= 4096*... : x = ((B >> 4) << 8) | R needs width <= 4096; for wider
    images the top x bit is dropped (still a deterministic pattern)."""
    x = np.arange(width, dtype=np.uint32)[None, :]
    y = np.arange(height, dtype=np.uint32)[:, None]
    img = np.empty((height, width, 4), dtype=np.uint8)
    r = np.broadcast_to((x & 255).astype(np.uint8), (height, width))
    g = np.broadcast_to((y & 255).astype(np.uint8), (height, width))
    img[..., 0] = ~r if negative else r
    img[..., 1] = ~g if negative else g
    img[..., 2] = ((((x >> 8) & 15) << 4) | ((y >> 8) & 15)).astype(np.uint8)
    img[..., 3] = 255
    return img


def noise_background(width: int, height: int, seed: int = 20251017) -> np.ndarray:
    """Seeded random RGBA8 texels — every neighbouring texel differs, so an off-by-one lookup
    cannot hide."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(height, width, 4), dtype=np.uint8)
    img[..., 3] = 255
    return img
