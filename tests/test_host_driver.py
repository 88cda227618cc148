"""Host driver: settings / TOML keys / defaults / validation of
src/settings.rs, the CSV path reader (src/csv.rs), the time interpolator with the reference's
index quirk (src/interpolation.rs:63-91), the frame-time loop (src/rendering.rs:224-238) and the
CLI error path (src/main.rs:219-227).  CPU only; the GPU end-to-end runs are in
test_gpu_driver.py."""
import math
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH_CSV = os.path.join(ROOT, "curvis_b200", "paths", "path_through.csv")


def test_default_settings_values():
    from curvis_b200 import settings as S
    cam, sim = S.CameraSettings.default(), S.SimulationSettings.default()
    assert (cam.resolution_x, cam.resolution_y, cam.diagonal, cam.focal_length) == (960, 540, 43.0, 15.0)
    assert (sim.escape_radius, sim.ray_integration_max_itarations, sim.ray_integration_step) == (100.0, 40000, 0.05)
    assert (sim.sampling_initial_nums, sim.sampling_max_iterations) == (100, 50)
    assert (sim.sampling_convergence_threshold_1, sim.sampling_convergence_threshold_2) == (1e-5, 1e-5)
    img = S.ImageSettings.default()
    assert (img.image_name, img.t, img.l, img.theta, img.phi) == ("output_image", 0.0, 5.0, math.pi / 2, 0.0)
    assert (img.forward_x, img.forward_y, img.forward_z, img.up_x, img.up_y, img.up_z) == (-1.0, 0.0, 0.0, 0.0, 0.0, 1.0)
    assert S.EllisMetricSettings.default().rho == 1.0
    inter = S.InterstellarMetricSettings.default()
    assert (inter.m, inter.a, inter.rho) == (0.1, 1e-4, 1.0)
    vid = S.VideoSettings.default()
    assert (vid.video_name, vid.frame_rate, vid.filepath_to_camera_path) == ("output_video", 30.0, "paths/path_through.csv")
    vid.normalize()
    vid.validate()                                            # the shipped sample path exists and is a .csv
    assert os.path.samefile(vid.filepath_to_camera_path, PATH_CSV)


def test_settings_validation_and_toml(tmp_path):
    from curvis_b200 import settings as S
    for field, value, msg in [("escape_radius", 0.0, "escape radius"), ("ray_integration_max_itarations", 0, "maximum number of iterations"),
                              ("ray_integration_step", -1.0, "step for the ray integration"), ("sampling_initial_nums", 1, "initial number of samples"),
                              ("sampling_max_iterations", 0, "sampling"), ("sampling_convergence_threshold_1", 0.0, "first convergence"),
                              ("sampling_convergence_threshold_2", 0.0, "second convergence")]:
        sim = S.SimulationSettings.default()
        setattr(sim, field, value)
        with pytest.raises(S.SettingsError, match=msg):
            sim.validate()
    cam = S.CameraSettings.default()
    cam.diagonal = 0.0
    with pytest.raises(S.SettingsError, match="diagonal"):
        cam.validate()
    f = tmp_path / "sim.toml"
    f.write_text("escape_radius = 25.0\nray_integration_max_itarations = 1000\nray_integration_step = 0.05\nsampling_initial_nums = 100\n"
                 "sampling_max_iterations = 50\nsampling_convergence_threshold_1 = 1e-5\nsampling_convergence_threshold_2 = 1e-5\n")
    sim = S.SimulationSettings.from_toml_file(str(f))
    assert (sim.escape_radius, sim.ray_integration_max_itarations) == (25.0, 1000)
    f.write_text("escape_radius = 25.0\n")
    with pytest.raises(S.SettingsError, match="missing field"):
        S.SimulationSettings.from_toml_file(str(f))
    with pytest.raises(S.SettingsError):
        S.SimulationSettings.from_toml_file(str(tmp_path / "nope.toml"))
    # metric file: tried as Interstellar first, then Ellis (cli.rs:248-256)
    m = tmp_path / "metric.toml"
    m.write_text("rho = 2.5\n")
    assert isinstance(S.metric_settings_from_file(str(m)), S.EllisMetricSettings)
    m.write_text("m = 0.2\na = 0.001\nrho = 2.5\n")
    assert isinstance(S.metric_settings_from_file(str(m)), S.InterstellarMetricSettings)
    assert isinstance(S.metric_settings_from_file(None), S.EllisMetricSettings)


def test_csv_path_and_interpolator_quirk():
    from curvis_b200.interpolation import Interpolator, load_path
    pos, fwd, up = load_path(PATH_CSV)
    assert pos.shape == (1000, 4) and fwd.shape == (1000, 3) and up.shape == (1000, 3)
    assert pos[0].tolist() == [0.0, -4.0, math.pi / 2, 0.0] and pos[-1][:2].tolist() == [20.0, 4.0]
    it = Interpolator(pos, fwd, up)
    assert (it.min_time(), it.max_time()) == (0.0, 20.0)
    assert it.time_indexes_and_frac_from_time(0.0) == (0, 1, 0.0)             # loop does not run: frac = 0
    # t between way-points 2 and 3: bracket (2,3) but indices (3,4) are returned (interpolation.rs:85-90)
    t = 0.5 * (pos[2][0] + pos[3][0])
    i1, i2, frac = it.time_indexes_and_frac_from_time(t)
    assert (i1, i2) == (3, 4) and frac == pytest.approx(0.5)
    np.testing.assert_allclose(it.camera_position(t), pos[3] + frac * (pos[4] - pos[3]))
    fixed = Interpolator(pos, fwd, up, corrected=True)
    assert fixed.time_indexes_and_frac_from_time(t)[:2] == (2, 3)
    with pytest.raises(ValueError):
        it.camera_position(-0.1)
    with pytest.raises(ValueError):
        it.camera_position(20.1)
    with pytest.raises(IndexError):                                           # README.md:107 "panics on the last frame"
        it.camera_position(19.995)
    fixed.camera_position(19.995)


def test_times_of_frames_and_the_301st_frame_panic():
    """30 fps -> 600 frames; at 15 fps the 301st frame time is 19.99999999999995 and the
    reference panics on it."""
    from curvis_b200.interpolation import Interpolator
    from curvis_b200.rendering import VideoRenderingSystem

    class Stub(VideoRenderingSystem):
        def __init__(self, fps):
            self.interpolator = Interpolator.from_file(PATH_CSV)
            self.video_rendering_settings = type("S", (), {"frame_rate": fps})()

    assert len(Stub(30.0).times_of_frames()) == 600
    times = Stub(15.0).times_of_frames()
    assert len(times) == 301 and times[300] == 19.99999999999995
    it = Interpolator.from_file(PATH_CSV)
    it.camera_position(times[299])
    with pytest.raises(IndexError):
        it.camera_position(times[300])


def test_cli_errors_exit_1(capsys):
    from curvis_b200.cli import main
    assert main(["image", "/no/such/a.png", "/no/such/b.png"]) == 1
    assert "not found" in capsys.readouterr().err
    assert main([]) == 1
    assert main(["custom"]) == 1


def test_load_image_converts_16_bit_like_the_image_crate(tmp_path):
    """DynamicImage::get_pixel on a 16-bit image returns (x + 128) / 257 per sample (image 0.25); PIL's own conversions clip
    ("I;16") or truncate (16-bit RGB) instead."""
    import numpy as np
    from PIL import Image
    from curvis_b200.rendering import load_image
    g = (np.arange(6 * 8, dtype=np.uint32).reshape(6, 8) * 1371 % 65536).astype(np.uint16)
    g[0, :4] = [0, 127, 128, 65535]
    Image.fromarray(g).save(tmp_path / "g16.png")
    got = load_image(str(tmp_path / "g16.png"))
    want = ((g.astype(np.uint32) + 128) // 257).astype(np.uint8)
    assert got.shape == (6, 8, 4) and (got[..., 3] == 255).all()
    for c in range(3):
        assert (got[..., c] == want).all()
    rgb = np.random.default_rng(3).integers(0, 256, (5, 7, 3), dtype=np.uint8)
    Image.fromarray(rgb).save(tmp_path / "rgb8.png")
    got8 = load_image(str(tmp_path / "rgb8.png"))
    assert (got8[..., :3] == rgb).all() and (got8[..., 3] == 255).all()


def test_shipped_camera_paths_are_the_reference_files():
    """paths/path_through.csv and path_orbit.csv are INPUT DATA of BASELINE config 5: shipped byte for byte (round 1
    regenerated them with numpy and 0.33 % of the entries differed in the last bit)."""
    import hashlib
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "curvis_b200", "paths")
    sums = {f: hashlib.sha256(open(os.path.join(root, f), "rb").read()).hexdigest() for f in ("path_through.csv", "path_orbit.csv")}
    assert sums == {"path_through.csv": "747b87a2179125188d3cae3f79cace8571aaaa76ebf3b1f27f18f6a9d0e9ac3d", "path_orbit.csv": "fe3872182c0743643b358bbdc0128145a862a7e54db3bbcf23823ffefec3b5ed"}
