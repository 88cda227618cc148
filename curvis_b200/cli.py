"""``python -m curvis_b200 image|video|custom`` — the command tree of reference src/cli.rs:35-122
and src/main.rs:135-235: positionals ``background_image_1 background_image_2 [output_folder]``,
options ``-i/--image-settings``, ``-v/--video-settings``, ``-m/--metric-settings``,
``-c/--camera-settings``, ``-s/--simulation-settings`` (TOML files; missing -> defaults).
Extensions (absent from the reference): ``--renderer``, ``--precision``, ``--devices``, ``--frames``,
``--corrected-interpolation``."""
from __future__ import annotations

import argparse
import os
import sys

from . import settings as S
from .rendering import (ImageRenderingSettings, ImageRenderingSystem, VideoRenderingSettings, VideoRenderingSystem,
                        instantiate_metric)
from .systems import Context


def _common(sub):
    sub.add_argument("background_image_1")
    sub.add_argument("background_image_2")
    sub.add_argument("output_folder", nargs="?", default=None)
    sub.add_argument("-m", "--metric-settings", dest="metric_settings")
    sub.add_argument("-c", "--camera-settings", dest="camera_settings")
    sub.add_argument("-s", "--simulation-settings", dest="simulation_settings")
    sub.add_argument("--renderer", choices=["efficient", "per_pixel"], default="efficient",
                     help="efficient = render_image_efficient (the reference binary's choice); per_pixel = render_image")
    sub.add_argument("--precision", choices=["f64", "f64_fast", "f32"], default="f64",
                     help="f64 = one rounding per reference operation; f64_fast = the same fp64 Euler iteration regrouped for the "
                          "GPU (3x faster, same frames); f32 = fp32 right-hand side (per_pixel renderer only)")
    sub.add_argument("--devices", default=None, help="comma-separated CUDA ordinals (default: all visible)")
    sub.add_argument("--frame", choices=["local", "world", "world_quirk"], default="local",
                     help="per_pixel renderer: local = render_image as written (systems.rs:540-561); world = escaped_photon_to_world_direction "
                          "(systems.rs:144-187) with frame_field_33; world_quirk = the same rotation with metrics.rs:347 as written")
    sub.add_argument("--coordinates", choices=["spherical", "cartesian"], default="spherical",
                     help="per_pixel renderer: cartesian = chart-free angular state, no coordinate poles (f64 only)")
    sub.add_argument("--integrator", choices=["euler", "rk4", "euler_adaptive"], default="euler", help="per_pixel renderer, f64 only")
    sub.add_argument("--step-tolerance", type=float, default=0.01, help="euler_adaptive: largest azimuth / polar advance of one step")


def get_cli() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="curvis", description="Ray tracing of wormhole space-times on B200 GPUs")
    subs = ap.add_subparsers(dest="command")
    img = subs.add_parser("image")
    _common(img)
    img.add_argument("-i", "--image-settings", dest="image_settings")
    vid = subs.add_parser("video")
    _common(vid)
    vid.add_argument("-v", "--video-settings", dest="video_settings")
    vid.add_argument("--frames", type=int, default=None, help="render only the first N frames")
    vid.add_argument("--corrected-interpolation", action="store_true")
    vid.add_argument("--sharding", choices=["frames", "rows"], default="frames",
                     help="several --devices: frames = frame i rendered whole by device i mod N (default); rows = every frame "
                          "row-interleaved over the devices")
    vid.add_argument("--encoder-threads", type=int, default=8)
    vid.add_argument("--compress-level", type=int, default=3, help="PNG deflate level of the written frames")
    vid.add_argument("--no-write", action="store_true", help="render without writing PNG files (timing)")
    subs.add_parser("custom")
    return ap


def _sim_options(args) -> dict:
    """The curvis_sim extension fields (include/curvis_gpu.h) of the per-pixel renderer; all defaults = the reference."""
    from . import _abi
    opts = {}
    if args.frame != "local":
        opts["frame"] = {"world": _abi.FRAME_WORLD, "world_quirk": _abi.FRAME_WORLD_QUIRK}[args.frame]
    if args.coordinates == "cartesian":
        opts["coordinates"] = _abi.COORDINATES_CARTESIAN
    if args.integrator != "euler":
        opts["integrator"] = {"rk4": _abi.INTEGRATOR_RK4, "euler_adaptive": _abi.INTEGRATOR_EULER_ADAPTIVE}[args.integrator]
        if args.integrator == "euler_adaptive":
            opts["step_tolerance"] = args.step_tolerance
    if opts and args.renderer != "per_pixel":
        raise S.SettingsError("--frame / --coordinates / --integrator belong to --renderer per_pixel")
    return opts


def _existing(path: str) -> str:
    if not os.path.exists(path):
        raise S.SettingsError(f"File {path!r} not found.")        # cli.rs:156-161
    return path


def _load(cls, path):
    return cls.default() if path is None else cls.from_toml_file(_existing(path))


def _output_folder(path):
    if path is None:
        return os.getcwd()                                         # cli.rs:196-203
    _existing(path)
    if not os.path.isdir(path):
        raise S.SettingsError(f"{path!r} is not a folder.")
    return path


def main(argv=None) -> int:
    args = get_cli().parse_args(argv)
    try:
        if args.command is None:
            print("Subcommand not found", file=sys.stderr)
            return 1
        if args.command == "custom":
            raise S.SettingsError("The custom subcommand is a placeholder in the reference (src/custom.rs:4-8).")
        bg1, bg2 = _existing(args.background_image_1), _existing(args.background_image_2)
        out = _output_folder(args.output_folder)
        metric = instantiate_metric(S.metric_settings_from_file(args.metric_settings))
        camera, simulation = _load(S.CameraSettings, args.camera_settings), _load(S.SimulationSettings, args.simulation_settings)
        devices = [int(d) for d in args.devices.split(",")] if args.devices else None
        if args.command == "image":
            ctx = Context(devices)
            image = _load(S.ImageSettings, args.image_settings)
            settings = ImageRenderingSettings.from_settings(bg1, bg2, out, image, camera, simulation)
            path = ImageRenderingSystem(metric, settings, context=ctx, renderer=args.renderer, precision=args.precision,
                                        sim_options=_sim_options(args)).render()
            print(f"Saved {path}")
        else:
            video = _load(S.VideoSettings, args.video_settings)
            settings = VideoRenderingSettings.from_settings(bg1, bg2, out, video, camera, simulation)
            system = VideoRenderingSystem(metric, settings, renderer=args.renderer, devices=devices, sharding=args.sharding,
                                          corrected_interpolation=args.corrected_interpolation, precision=args.precision,
                                          sim_options=_sim_options(args))
            try:
                folder = system.render(max_frames=args.frames, encoder_threads=args.encoder_threads, compress_level=args.compress_level,
                                       write_frames=not args.no_write)
            finally:
                if system.last_render_info:
                    info = system.last_render_info
                    print(f"{info['frames']} frames on {info['devices']} device(s) in {info['wall_s']:.2f} s "
                          f"({info['frames_per_s']:.1f} frames/s, PNG {'written' if info['png_written'] else 'not written'})")
            print(f"Frames in {folder}")
        return 0
    except Exception as e:                                          # main.rs:219-227: print the error, exit 1
        print(f"Error: {e}", file=sys.stderr)
        return 1


if __name__ == "__main__":
    sys.exit(main())
