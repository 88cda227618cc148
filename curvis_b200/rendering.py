"""Host driver: the image and video rendering systems of reference src/rendering.rs, on top of
the GPU renderers.  Like the reference, ``render()`` calls the table-based renderer
(``render_image_efficient``, src/rendering.rs:97 and :299) by default; ``renderer="per_pixel"``
selects the per-pixel integrator ``render_image`` (src/systems.rs:307-330) instead.  File names
follow the reference: ``<out>/<image_name>.png`` and ``<out>/tmp/frame_{index}.png``."""
from __future__ import annotations

import os
import shutil
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .cameras import Camera
from .images import SphericalImage
from .interpolation import Interpolator
from .metrics import EllisMetric, InterstellarMetric
from .settings import (CameraSettings, EllisMetricSettings, ImageSettings, InterstellarMetricSettings, SimulationSettings,
                       VideoSettings)
from . import _abi
from .systems import Context, RelativisticSystem


def _to_u8(samples: np.ndarray) -> np.ndarray:
    """16-bit samples to 8 bits the way the `image` crate converts a pixel (image 0.25: (x + 128) / 257, i.e. rounding
    x * 255 / 65535), which is what DynamicImage::get_pixel returns for a 16-bit image."""
    return ((samples.astype(np.uint32) + 128) // 257).astype(np.uint8)


def load_image(path: str) -> np.ndarray:
    """images.rs:7-11 + what DynamicImage::get_pixel yields (:107-111): RGBA8 texels.  8-bit images go through PIL's
    RGBA conversion (the same bytes the `image` crate hands out).  16-bit images are converted explicitly with the crate's
    rounding — PIL's own conversion of "I;16" clips to 255 and of 16-bit RGB truncates to the high byte, which would put
    texels one level (or, for 16-bit grey, everything) off the reference's."""
    from PIL import Image
    with Image.open(path) as im:
        if im.mode in ("I;16", "I;16L", "I;16B", "I;16N", "I"):          # 16-bit greyscale (PIL loads some as 32-bit "I")
            g = np.asarray(im)
            if g.dtype.itemsize > 2 and int(g.max(initial=0)) > 65535:
                raise ValueError(f"{path}: greyscale samples beyond 16 bits are not supported")
            g8 = _to_u8(g.astype(np.uint32))
            out = np.empty(g8.shape + (4,), dtype=np.uint8)
            out[..., 0] = out[..., 1] = out[..., 2] = g8
            out[..., 3] = 255
            return np.ascontiguousarray(out)
        arr = np.asarray(im)
        if arr.dtype == np.uint16:                                        # 16-bit RGB / RGBA (e.g. decoded by a plugin)
            a8 = _to_u8(arr)
            if a8.ndim == 2:
                a8 = np.repeat(a8[..., None], 3, axis=2)
            if a8.shape[2] == 3:
                a8 = np.concatenate([a8, np.full(a8.shape[:2] + (1,), 255, dtype=np.uint8)], axis=2)
            return np.ascontiguousarray(a8[..., :4])
        return np.ascontiguousarray(np.asarray(im.convert("RGBA"), dtype=np.uint8))


def save_image(rgb8: np.ndarray, path: str, compress_level: int = 6) -> None:
    """images.rs:13-15: the frame is an RGB8 image saved as PNG (lossless: the pixels are what
    matters, not the deflate level)."""
    from PIL import Image
    Image.fromarray(rgb8, mode="RGB").save(path, compress_level=compress_level)


def load_image_as_spherical_image(path: str, forward=None, up=None) -> SphericalImage:   # images.rs:186-193
    return SphericalImage(load_image(path), forward, up)


def instantiate_metric(metric_settings):                         # main.rs:114-132
    if isinstance(metric_settings, InterstellarMetricSettings):
        return InterstellarMetric(metric_settings.m, metric_settings.a, metric_settings.rho)
    return EllisMetric(metric_settings.rho)


@dataclass
class ImageRenderingSettings:                                    # rendering.rs:148-167, filled like main.rs:54-111
    resolution_x: int
    resolution_y: int
    camera_diagonal: float
    camera_focal_length: float
    camera_position: tuple
    camera_forward: tuple
    camera_up: tuple
    path_to_background_image_1: str
    path_to_background_image_2: str
    path_to_output_folder: str
    output_image_name: str
    escape_radius: float
    max_iterations_propagation: int
    ray_integration_step: float
    alphas_num: int
    max_iterations_sampling: int
    sampling_convergence_threshold_1: float
    sampling_convergence_threshold_2: float

    @classmethod
    def from_settings(cls, bg1, bg2, out, image: ImageSettings, camera: CameraSettings, simulation: SimulationSettings):
        for s in (image, camera, simulation):
            s.normalize()
            s.validate()
        return cls(camera.resolution_x, camera.resolution_y, camera.diagonal, camera.focal_length,
                   (image.t, image.l, image.theta, image.phi), (image.forward_x, image.forward_y, image.forward_z),
                   (image.up_x, image.up_y, image.up_z), bg1, bg2, out, image.image_name,
                   simulation.escape_radius, simulation.ray_integration_max_itarations, simulation.ray_integration_step,
                   simulation.sampling_initial_nums,
                   simulation.sampling_initial_nums,      # main.rs:106-107 feeds sampling_initial_nums into BOTH fields
                   simulation.sampling_convergence_threshold_1, simulation.sampling_convergence_threshold_2)


@dataclass
class VideoRenderingSettings:                                    # rendering.rs:355-374, filled like main.rs:14-51
    frame_rate: float
    resolution_x: int
    resolution_y: int
    camera_diagonal: float
    camera_focal_length: float
    filepath_to_camera_path: str
    filepath_to_background_image_1: str
    filepath_to_background_image_2: str
    filepath_to_output_folder: str
    output_video_name: str
    escape_radius: float
    max_iterations_propagation: int
    ray_integration_step: float
    alphas_num: int
    max_iterations_sampling: int
    sampling_convergence_threshold_1: float
    sampling_convergence_threshold_2: float

    @classmethod
    def from_settings(cls, bg1, bg2, out, video: VideoSettings, camera: CameraSettings, simulation: SimulationSettings):
        for s in (video, camera, simulation):
            s.normalize()
            s.validate()
        return cls(video.frame_rate, camera.resolution_x, camera.resolution_y, camera.diagonal, camera.focal_length,
                   video.filepath_to_camera_path, bg1, bg2, out, video.video_name,
                   simulation.escape_radius, simulation.ray_integration_max_itarations, simulation.ray_integration_step,
                   simulation.sampling_initial_nums, simulation.sampling_initial_nums,
                   simulation.sampling_convergence_threshold_1, simulation.sampling_convergence_threshold_2)


# extension: curvis_sim.precision of include/curvis_gpu.h (the table-based renderer accepts f64 and f64_fast)
PRECISIONS = {"f64": _abi.PRECISION_F64, "f64_fast": _abi.PRECISION_F64_FAST, "f32": _abi.PRECISION_F32}


class ImageRenderingSystem:
    """rendering.rs:20-117."""

    def __init__(self, metric, settings: ImageRenderingSettings, context: Optional[Context] = None, renderer: str = "efficient",
                 precision: str = "f64", sim_options: Optional[dict] = None):
        self.image_rendering_settings = settings
        self.renderer = renderer
        self.precision = PRECISIONS[precision]
        self.sim_options = dict(sim_options or {})     # curvis_sim extensions of the per-pixel renderer (frame, coordinates, ...)
        image_1 = load_image_as_spherical_image(settings.path_to_background_image_1)
        image_2 = load_image_as_spherical_image(settings.path_to_background_image_2)
        camera = Camera(settings.camera_position, settings.camera_forward, settings.camera_up, settings.camera_focal_length,
                        settings.camera_diagonal, settings.resolution_x, settings.resolution_y)
        self.relativistic_system = RelativisticSystem(metric, image_1, image_2, camera, context=context)

    def render_frame(self) -> np.ndarray:
        s = self.image_rendering_settings
        if self.renderer == "per_pixel":
            return self.relativistic_system.render_image(s.max_iterations_propagation, s.escape_radius, s.ray_integration_step,
                                                         precision=self.precision, **self.sim_options)
        return self.relativistic_system.render_image_efficient(
            s.max_iterations_propagation, s.escape_radius, s.ray_integration_step, s.alphas_num, s.max_iterations_sampling,
            s.sampling_convergence_threshold_1, s.sampling_convergence_threshold_2, precision=self.precision)

    def render(self) -> str:
        s = self.image_rendering_settings
        if not os.path.exists(s.path_to_output_folder):            # rendering.rs:89-95
            os.mkdir(s.path_to_output_folder)
        frame = self.render_frame()
        path = os.path.join(s.path_to_output_folder, s.output_image_name) + ".png"   # :108
        save_image(frame, path)
        return path


class VideoRenderingSystem:
    """rendering.rs:170-327.  Extensions (absent from the single-threaded reference): ``devices`` — CUDA ordinals the frames
    are spread over; ``sharding`` — "frames" (default for more than one device: frame i is rendered whole by device
    i mod N, one context, host thread and pinned frame ring per device, no exchange at all) or "rows" (every frame
    row-interleaved over the devices of ONE context by curvis_render_image)."""

    def __init__(self, metric, settings: VideoRenderingSettings, context: Optional[Context] = None, renderer: str = "efficient",
                 corrected_interpolation: bool = False, precision: str = "f64", devices: Optional[List[int]] = None,
                 sharding: str = "frames", sim_options: Optional[dict] = None):
        self.video_rendering_settings = settings
        self.renderer = renderer
        self.precision = PRECISIONS[precision]
        self.sim_options = dict(sim_options or {})
        self.interpolator = Interpolator.from_file(settings.filepath_to_camera_path, corrected=corrected_interpolation)
        image_1 = load_image_as_spherical_image(settings.filepath_to_background_image_1)
        image_2 = load_image_as_spherical_image(settings.filepath_to_background_image_2)
        t0 = self.interpolator.min_time()
        camera = Camera(self.interpolator.camera_position(t0), self.interpolator.camera_forward(t0), self.interpolator.camera_up(t0),
                        settings.camera_focal_length, settings.camera_diagonal, settings.resolution_x, settings.resolution_y)
        if sharding not in ("frames", "rows"):
            raise ValueError("sharding must be 'frames' or 'rows'")
        self.sharding = sharding
        if devices is not None and len(devices) > 1 and sharding == "frames" and context is None:
            # one context (and later one host thread) per device: distinct contexts may run concurrently (curvis_gpu.h)
            self.systems = [RelativisticSystem(metric, image_1, image_2, camera, context=Context([d])) for d in devices]
        else:
            ctx = context if context is not None else (Context(list(devices)) if devices else None)
            self.systems = [RelativisticSystem(metric, image_1, image_2, camera, context=ctx)]
        self.relativistic_system = self.systems[0]
        self.last_render_info: Optional[dict] = None

    def times_of_frames(self) -> List[float]:                      # rendering.rs:224-238
        min_time, max_time = self.interpolator.min_time(), self.interpolator.max_time()
        delta_time = 1.0 / self.video_rendering_settings.frame_rate
        times, t = [], min_time
        while t < max_time:
            times.append(t)
            t += delta_time
        return times

    def update_camera(self, t: float) -> None:                     # rendering.rs:242-253
        cam = self.relativistic_system.camera
        cam.update_position(self.interpolator.camera_position(t))
        cam.update_orientation(self.interpolator.camera_forward(t), self.interpolator.camera_up(t))

    def camera_at(self, t: float) -> Camera:
        """The camera update_camera(t) would leave behind, as a fresh object (frames in flight keep their own)."""
        s = self.video_rendering_settings
        return Camera(self.interpolator.camera_position(t), self.interpolator.camera_forward(t), self.interpolator.camera_up(t),
                      s.camera_focal_length, s.camera_diagonal, s.resolution_x, s.resolution_y)

    def render_frame(self, system: Optional[RelativisticSystem] = None, out=None) -> np.ndarray:
        s = self.video_rendering_settings
        system = system if system is not None else self.relativistic_system
        if self.renderer == "per_pixel":
            return system.render_image(s.max_iterations_propagation, s.escape_radius, s.ray_integration_step, out=out,
                                       precision=self.precision, **self.sim_options)
        return system.render_image_efficient(
            s.max_iterations_propagation, s.escape_radius, s.ray_integration_step, s.alphas_num, s.max_iterations_sampling,
            s.sampling_convergence_threshold_1,
            s.sampling_convergence_threshold_1,                     # rendering.rs:305-306 passes threshold_1 twice
            out=out, precision=self.precision)

    def render(self, max_frames: Optional[int] = None, verbose: bool = True, encoder_threads: int = 8, compress_level: int = 3,
               write_frames: bool = True) -> str:
        """The frame loop of rendering.rs:258-327.  Rendering a 4K frame takes milliseconds on
        the GPU while its PNG encode takes a sizeable fraction of a second on one core, so frames
        are handed to a pool of encoder threads (zlib releases the GIL).  With several devices every device has its own
        host thread, context and ring of page-locked frames; frame i goes to device i mod N.  Camera updates are evaluated
        in frame order first: where the reference panics (its interpolator's last-frame bug, interpolation.rs:63-91) the
        frames before that point are rendered and written, then the error is raised — what the reference leaves on disk."""
        import time
        from concurrent.futures import ThreadPoolExecutor
        s = self.video_rendering_settings
        times = self.times_of_frames()
        if max_frames is not None:
            times = times[:max_frames]
        if not os.path.exists(s.filepath_to_output_folder):        # rendering.rs:266-274
            os.mkdir(s.filepath_to_output_folder)
        tmp_folder = os.path.join(s.filepath_to_output_folder, "tmp")
        if os.path.exists(tmp_folder):                              # :277-282
            shutil.rmtree(tmp_folder)
        os.mkdir(tmp_folder)
        if verbose:
            print(f"Rendering {len(times)} frames...")
        cameras, failure = [], None
        for t in times:
            try:
                cameras.append(self.camera_at(t))                   # may raise on the last frame, like the reference panics
            except Exception as e:                                  # noqa: BLE001 — re-raised below, after the frames before it
                failure = e
                break
        n_dev = len(self.systems)
        shape = (s.resolution_y, s.resolution_x, 3)
        ring = 3                                                    # frames in flight per device (render, encode, encode)
        t_start = time.perf_counter()
        render_s = [0.0] * n_dev

        def device_worker(g: int, pool) -> None:
            system = self.systems[g]
            buffers = [np.empty(shape, dtype=np.uint8) for _ in range(ring)]
            for b in buffers:
                system.context.register_host_buffer(b)              # the kernel stores its pixels straight into the frame
            busy = [None] * ring
            try:
                for k, index in enumerate(range(g, len(cameras), n_dev)):
                    slot = k % ring
                    if busy[slot] is not None:
                        busy[slot].result()                          # its previous frame has been encoded
                    if verbose:
                        print(f"Rendering frame {index + 1}/{len(times)}...")
                    system.camera = cameras[index]
                    t0 = time.perf_counter()
                    frame = self.render_frame(system, out=buffers[slot])
                    render_s[g] += time.perf_counter() - t0
                    busy[slot] = pool.submit(save_image, frame, os.path.join(tmp_folder, f"frame_{index}.png"), compress_level) \
                        if write_frames else None
            finally:
                for f in busy:
                    if f is not None:
                        f.result()
                for b in buffers:
                    system.context.unregister_host_buffer(b)

        with ThreadPoolExecutor(max_workers=max(1, encoder_threads)) as pool:
            if n_dev == 1:
                device_worker(0, pool)
            else:
                with ThreadPoolExecutor(max_workers=n_dev) as devices_pool:
                    for f in [devices_pool.submit(device_worker, g, pool) for g in range(n_dev)]:
                        f.result()
        wall = time.perf_counter() - t_start
        self.last_render_info = {"frames": len(cameras), "devices": n_dev, "sharding": self.sharding if n_dev > 1 or self.systems[0].context.device_count() > 1 else "single device",
                                 "wall_s": wall, "frames_per_s": len(cameras) / wall if wall > 0 else None,
                                 "render_s_per_device": render_s, "png_written": write_frames, "compress_level": compress_level}
        if self.systems[0] is self.relativistic_system and cameras:
            self.relativistic_system.camera = cameras[-1]           # where the reference's loop leaves the camera
        if failure is not None:
            raise failure
        return tmp_folder
