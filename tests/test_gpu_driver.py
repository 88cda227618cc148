"""End to end through the host driver on a GPU: `curvis image` / `curvis video` equivalents write
the PNGs the reference would (names, sizes) and their pixels equal the oracle's renders."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _write_backgrounds(tmp_path):
    from PIL import Image
    from curvis_b200 import scenes
    bp, bn = scenes.decodable_background(1024, 512), scenes.decodable_background(1024, 512, True)
    p1, p2 = str(tmp_path / "bg1.png"), str(tmp_path / "bg2.png")
    Image.fromarray(bp, "RGBA").save(p1)
    Image.fromarray(bn[..., :3].copy(), "RGB").save(p2)           # an RGB file: alpha 255 is implied
    return p1, p2, bp, bn


def test_cli_image_both_renderers(tmp_path, oracle):
    from PIL import Image
    from curvis_b200 import scenes
    from curvis_b200.cli import main
    p1, p2, bp, bn = _write_backgrounds(tmp_path)
    cam = tmp_path / "cam.toml"
    cam.write_text("resolution_x = 128\nresolution_y = 72\ndiagonal = 43.0\nfocal_length = 15.0\n")
    out = tmp_path / "out"
    assert main(["image", p1, p2, str(out)]) == 1                 # output folder must exist (cli.rs:205-209)
    out.mkdir()
    assert main(["image", p1, p2, str(out), "-c", str(cam)]) == 0
    got = np.asarray(Image.open(out / "output_image.png"))
    assert got.shape == (72, 128, 3)
    ocam = oracle.camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 128, 72)
    ref, _ = oracle.render_image_efficient(oracle.metric("ellis"), ocam, oracle.sim(40000, 100.0, 0.05), bp, bn, 100, 100, 1e-5, 1e-5)
    assert (got == ref).all(axis=2).mean() >= 0.9999
    assert main(["image", p1, p2, str(out), "-c", str(cam), "--renderer", "per_pixel"]) == 0
    got = np.asarray(Image.open(out / "output_image.png"))
    ref, _, _ = oracle.render_rows(oracle.metric("ellis"), ocam, oracle.sim(40000, 100.0, 0.05), bp, bn, threads=os.cpu_count() or 1)
    assert (got == ref).all()
    # --precision (extension): the regrouped fp64 kernel renders the same frame with both renderers; fp32 is
    # per-pixel only
    assert main(["image", p1, p2, str(out), "-c", str(cam), "--renderer", "per_pixel", "--precision", "f64_fast"]) == 0
    assert (np.asarray(Image.open(out / "output_image.png")) == ref).all(axis=2).mean() >= 0.9999
    assert main(["image", p1, p2, str(out), "-c", str(cam), "--precision", "f64_fast"]) == 0
    eff, _ = oracle.render_image_efficient(oracle.metric("ellis"), ocam, oracle.sim(40000, 100.0, 0.05), bp, bn, 100, 100, 1e-5, 1e-5)
    assert (np.asarray(Image.open(out / "output_image.png")) == eff).all(axis=2).mean() >= 0.9999
    assert main(["image", p1, p2, str(out), "-c", str(cam), "--renderer", "per_pixel", "--precision", "f32"]) == 0
    assert (np.asarray(Image.open(out / "output_image.png")) == ref).all(axis=2).mean() >= 0.99
    assert main(["image", p1, p2, str(out), "-c", str(cam), "--precision", "f32"]) == 1        # the table-based renderer is fp64 only


def test_cli_video_frames(tmp_path, oracle):
    from PIL import Image
    from curvis_b200.cli import main
    from curvis_b200.interpolation import Interpolator
    p1, p2, bp, bn = _write_backgrounds(tmp_path)
    cam = tmp_path / "cam.toml"
    cam.write_text("resolution_x = 96\nresolution_y = 54\ndiagonal = 43.0\nfocal_length = 15.0\n")
    met = tmp_path / "metric.toml"
    met.write_text("m = 0.1\na = 0.0001\nrho = 1.0\n")
    out = tmp_path / "video"
    out.mkdir()
    (out / "tmp").mkdir()
    (out / "tmp" / "stale.png").write_text("x")                   # a pre-existing tmp folder is wiped (rendering.rs:277-282)
    assert main(["video", p1, p2, str(out), "-c", str(cam), "-m", str(met), "--frames", "3"]) == 0
    names = sorted(os.listdir(out / "tmp"))
    assert names == ["frame_0.png", "frame_1.png", "frame_2.png"]
    it = Interpolator.from_file(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "curvis_b200", "paths", "path_through.csv"))
    for index in range(3):
        t = index / 30.0 if index else 0.0
        t = [0.0, 1.0 / 30.0, 1.0 / 30.0 + 1.0 / 30.0][index]
        ocam = oracle.camera(it.camera_position(t), it.camera_forward(t), it.camera_up(t), 15.0, 43.0, 96, 54)
        # video passes threshold_1 twice (rendering.rs:305-306); identical values at defaults
        ref, _ = oracle.render_image_efficient(oracle.metric("interstellar"), ocam, oracle.sim(40000, 100.0, 0.05), bp, bn, 100, 100, 1e-5, 1e-5)
        got = np.asarray(Image.open(out / "tmp" / f"frame_{index}.png"))
        assert (got == ref).all(axis=2).mean() >= 0.999, index


@pytest.mark.parametrize("sharding", ["frames", "rows"])
def test_video_over_several_devices_through_the_product_api(tmp_path, sharding):
    """VideoRenderingSystem(devices=[...]): the frame loop of rendering.rs:258-327 spread over the visible GPUs — frame
    sharding (one context + host thread per device) or row sharding (one multi-device context) — writes the same PNGs as
    one device.  With one visible GPU the device list is [0, 0]: two contexts on the same device still exercise the path."""
    import torch
    from PIL import Image
    import curvis_b200 as cv
    from curvis_b200 import settings as S
    from curvis_b200.rendering import VideoRenderingSettings, VideoRenderingSystem
    p1, p2, _, _ = _write_backgrounds(tmp_path)
    n = torch.cuda.device_count()
    devices = list(range(n)) if n >= 2 else ([0, 0] if sharding == "frames" else [0])
    camera = S.CameraSettings.default()
    camera.resolution_x, camera.resolution_y = 160, 90
    video, simulation = S.VideoSettings.default(), S.SimulationSettings.default()
    outs = {}
    for label, devs in (("one", [0]), ("many", devices)):
        out = tmp_path / f"video_{label}"
        out.mkdir()
        settings = VideoRenderingSettings.from_settings(p1, p2, str(out), video, camera, simulation)
        system = VideoRenderingSystem(cv.InterstellarMetric(0.1, 1e-4, 1.0), settings, renderer="per_pixel", precision="f64_fast",
                                      devices=devs, sharding=sharding)
        folder = system.render(max_frames=7, verbose=False, encoder_threads=4)
        assert system.last_render_info["frames"] == 7
        outs[label] = [np.asarray(Image.open(os.path.join(folder, f"frame_{i}.png"))) for i in range(7)]
    for a, b in zip(outs["one"], outs["many"]):
        assert a.shape == (90, 160, 3) and (a == b).all()
