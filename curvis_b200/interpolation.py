"""Camera path for video: the CSV reader of reference src/csv.rs:24-62 and the time
interpolator of src/interpolation.rs:45-112 — including its index quirk: after the scan
(t1, t2) bracket way-points (i-1, i) but the vectors of way-points (i, i+1) are blended, and a
time beyond the second-to-last way-point indexes out of range (the reference panics:
README.md:107 "Sometimes ... panics on the last frame").  ``corrected=True`` opts into the
bracketing way-points instead (an extension, off by default)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def load_path(path_to_csv_file: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """csv.rs:24-62: header line skipped; 10 comma-separated floats per row
    (t, l, theta, phi, fx, fy, fz, ux, uy, uz).  Returns positions (n,4), forwards (n,3), ups (n,3)."""
    rows: List[List[float]] = []
    with open(path_to_csv_file, "r") as f:
        for index, line in enumerate(f):
            if index == 0:
                continue
            line = line.rstrip("\n")
            values = [float(x) for x in line.split(",")]       # "Could not parse float"
            if len(values) < 10:
                raise ValueError(f"Could not read column {len(values)} of line {index + 1}")
            rows.append(values[:10])
    a = np.array(rows, dtype=np.float64).reshape(-1, 10)
    return a[:, 0:4].copy(), a[:, 4:7].copy(), a[:, 7:10].copy()


class Interpolator:
    def __init__(self, positions, forward_vectors, up_vectors, corrected: bool = False):
        self.positions = np.asarray(positions, dtype=np.float64)
        self.forward_vectors = np.asarray(forward_vectors, dtype=np.float64)
        self.up_vectors = np.asarray(up_vectors, dtype=np.float64)
        self.corrected = corrected

    @classmethod
    def from_file(cls, path_to_csv_file: str, corrected: bool = False) -> "Interpolator":
        return cls(*load_path(path_to_csv_file), corrected=corrected)

    def min_time(self) -> float:
        return float(self.positions[0][0])

    def max_time(self) -> float:
        return float(self.positions[-1][0])

    def time_indexes_and_frac_from_time(self, t: float) -> Tuple[int, int, float]:
        """interpolation.rs:63-91."""
        if t < self.min_time():
            raise ValueError("Interpolation time cannot be smaller than first time in positions[0].")
        if t > self.max_time():
            raise ValueError("Interpolation time cannot be greater than last time in positions[0].")
        t1, t2, i = self.min_time(), self.max_time(), 0
        while t > self.positions[i][0]:
            t1 = float(self.positions[i][0])
            t2 = float(self.positions[i + 1][0])
            i += 1
        frac = (t - t1) / (t2 - t1)
        if self.corrected and i > 0:
            return i - 1, i, frac
        return i, i + 1, frac

    def _blend(self, table: np.ndarray, t: float) -> np.ndarray:
        i1, i2, frac = self.time_indexes_and_frac_from_time(t)
        if not (0.0 <= frac <= 1.0):                              # interpolation.rs:36-38
            raise ValueError("frac must be between 0 and 1")
        if i2 >= len(table):
            raise IndexError(f"index out of bounds: the len is {len(table)} but the index is {i2}")   # the reference's panic
        v1, v2 = table[i1], table[i2]
        return v1 + frac * (v2 - v1)                              # interpolation.rs:19-29

    def camera_position(self, t: float) -> np.ndarray:
        return self._blend(self.positions, t)

    def camera_up(self, t: float) -> np.ndarray:
        return self._blend(self.up_vectors, t)

    def camera_forward(self, t: float) -> np.ndarray:
        return self._blend(self.forward_vectors, t)
