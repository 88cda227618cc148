"""Settings of the host driver — same TOML files, key names (typo included:
``ray_integration_max_itarations``), defaults and validation messages as reference
src/settings.rs:22-293; the defaults live in curvis_b200/settings/defaults/*.toml (values of the
reference's settings/defaults/*.toml)."""
from __future__ import annotations

import os
import tomllib
from dataclasses import dataclass, fields

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
DEFAULTS_DIR = os.path.join(PACKAGE_DIR, "settings", "defaults")


class SettingsError(ValueError):
    """The reference returns Err(String) (src/settings.rs Validate / FromToml)."""


def resolve_path(path: str) -> str:
    """filepaths.rs:42-47: a relative path is taken relative to the package directory."""
    return path if os.path.isabs(path) else os.path.join(PACKAGE_DIR, path)


class _Settings:
    default_file = ""

    @classmethod
    def from_toml_file(cls, path: str):
        if not path.endswith(".toml") or not os.path.exists(path):       # filepaths.rs:76-92
            raise SettingsError(f"The settings file {path!r} does not exist or is not a toml file.")
        with open(path, "rb") as f:
            try:
                data = tomllib.load(f)
            except tomllib.TOMLDecodeError as e:
                raise SettingsError(f"Could not parse {path!r}: {e}")
        names = [f.name for f in fields(cls)]
        missing = [n for n in names if n not in data]
        if missing:
            raise SettingsError(f"missing field `{missing[0]}` in {path!r}")
        try:
            return cls(**{n: data[n] for n in names})._coerce()
        except (TypeError, ValueError) as e:
            raise SettingsError(f"invalid value in {path!r}: {e}")

    @classmethod
    def default(cls):                                                    # settings.rs:259-293
        return cls.from_toml_file(os.path.join(DEFAULTS_DIR, cls.default_file))

    def _coerce(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if f.type in ("float", float):
                if isinstance(v, bool) or not isinstance(v, (int, float)):
                    raise TypeError(f"{f.name} must be a number")
                setattr(self, f.name, float(v))
            elif f.type in ("int", int):
                if isinstance(v, bool) or not isinstance(v, int) or v < 0 or v > 0xFFFFFFFF:
                    raise TypeError(f"{f.name} must be a u32")
            elif f.type in ("str", str) and not isinstance(v, str):
                raise TypeError(f"{f.name} must be a string")
        return self

    def normalize(self) -> None:
        pass

    def validate(self) -> None:
        pass


@dataclass
class VideoSettings(_Settings):                                          # settings.rs:22-57
    video_name: str
    frame_rate: float
    filepath_to_camera_path: str
    default_file = "video_settings.toml"

    def normalize(self):
        self.filepath_to_camera_path = resolve_path(self.filepath_to_camera_path)

    def validate(self):
        if not self.video_name:
            raise SettingsError("Video name cannot be an empty string.")
        if not self.filepath_to_camera_path.endswith(".csv"):
            raise SettingsError(f"The camera path {self.filepath_to_camera_path!r} is not a csv file.")
        if not os.path.exists(self.filepath_to_camera_path):
            raise SettingsError(f"The camera path {self.filepath_to_camera_path!r} does not exist.")


@dataclass
class ImageSettings(_Settings):                                          # settings.rs:58-81
    image_name: str
    t: float
    l: float
    theta: float
    phi: float
    forward_x: float
    forward_y: float
    forward_z: float
    up_x: float
    up_y: float
    up_z: float
    default_file = "image_settings.toml"

    def validate(self):
        if not self.image_name:
            raise SettingsError("Image name cannot be an empty string.")


@dataclass
class CameraSettings(_Settings):                                         # settings.rs:83-116
    resolution_x: int
    resolution_y: int
    diagonal: float
    focal_length: float
    default_file = "camera_settings.toml"

    def validate(self):
        if self.resolution_x <= 0:
            raise SettingsError("The resolution in the x direction must be larger than zero.")
        if self.resolution_y <= 0:
            raise SettingsError("The resolution in the y direction must be larger than zero.")
        if self.diagonal <= 0.0:
            raise SettingsError("The diagonal of the camera must be larger than zero.")
        if self.focal_length <= 0.0:
            raise SettingsError("The focal length of the camera must be larger than zero.")


@dataclass
class SimulationSettings(_Settings):                                     # settings.rs:118-166
    escape_radius: float
    ray_integration_max_itarations: int
    ray_integration_step: float
    sampling_initial_nums: int
    sampling_max_iterations: int
    sampling_convergence_threshold_1: float
    sampling_convergence_threshold_2: float
    default_file = "simulation_settings.toml"

    def validate(self):
        if self.escape_radius <= 0.0:
            raise SettingsError("The escape radius must be larger than zero.")
        if self.ray_integration_max_itarations <= 0:
            raise SettingsError("The maximum number of iterations for the ray integration must be larger than zero.")
        if self.ray_integration_step <= 0.0:
            raise SettingsError("The step for the ray integration must be larger than zero.")
        if self.sampling_initial_nums <= 1:
            raise SettingsError("The initial number of samples must be larger than two.")
        if self.sampling_max_iterations <= 0:
            raise SettingsError("The maximum number of iterations for the sampling must be larger than zero.")
        if self.sampling_convergence_threshold_1 <= 0.0:
            raise SettingsError("The first convergence threshold for the sampling must be larger than zero.")
        if self.sampling_convergence_threshold_2 <= 0.0:
            raise SettingsError("The second convergence threshold for the sampling must be larger than zero.")


@dataclass
class EllisMetricSettings(_Settings):                                    # settings.rs:168-186
    rho: float
    default_file = "ellis_metric_settings.toml"

    def validate(self):
        if self.rho <= 0.0:
            raise SettingsError("The density parameter rho must be larger than zero.")


@dataclass
class InterstellarMetricSettings(_Settings):                             # settings.rs:188-216
    m: float
    a: float
    rho: float
    default_file = "interstellar_metric_settings.toml"

    def validate(self):
        if self.m <= 0.0:
            raise SettingsError("The mass parameter m must be larger than zero.")
        if self.a <= 0.0:
            raise SettingsError("The spin parameter a must be larger than zero.")
        if self.rho <= 0.0:
            raise SettingsError("The density parameter rho must be larger than zero.")


def metric_settings_from_file(path):
    """cli.rs:233-261: no file -> Ellis defaults; a file is tried as Interstellar first, then Ellis."""
    if path is None:
        return EllisMetricSettings.default()
    try:
        return InterstellarMetricSettings.from_toml_file(path)
    except SettingsError:
        return EllisMetricSettings.from_toml_file(path)
