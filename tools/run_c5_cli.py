"""BASELINE.json configs[4] through the PRODUCT's command line (GPU box):

    python tools/run_c5_cli.py [n_frames] [out.json]

`python -m curvis_b200 video bg1 bg2 out -v video.toml -c camera.toml -s simulation.toml -m metric.toml --renderer per_pixel
 --precision f64_fast --devices 0,...,N-1`: paths/path_through.csv at 15 frames per second (frames 0..299; the 301st frame is
where the reference's interpolator panics), 3840x2160, Interstellar metric, 2000 iterations / step 0.05 / escape radius 45,
every visible GPU, frame sharding, PNG frames written to <out>/tmp like the reference.  Reports frames/s with and without the
PNG encode, and checks three frames of the run against a single-device render of the same frames."""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image
from curvis_b200 import scenes
from curvis_b200.cli import main

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 300
n_dev = torch.cuda.device_count()
work = tempfile.mkdtemp(prefix="curvis_c5_")
bp, bn = scenes.decodable_background(8192, 4096), scenes.decodable_background(8192, 4096, negative=True)
Image.fromarray(bp, "RGBA").save(os.path.join(work, "bg1.png"), compress_level=1)
Image.fromarray(bn, "RGBA").save(os.path.join(work, "bg2.png"), compress_level=1)
open(os.path.join(work, "video.toml"), "w").write('video_name = "output_video"\nframe_rate = 15.0\nfilepath_to_camera_path = "%s"\n' %
                                                  os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "curvis_b200", "paths", "path_through.csv"))
open(os.path.join(work, "camera.toml"), "w").write("resolution_x = 3840\nresolution_y = 2160\ndiagonal = 43.0\nfocal_length = 15.0\n")
open(os.path.join(work, "simulation.toml"), "w").write("escape_radius = 45.0\nray_integration_max_itarations = 2000\nray_integration_step = 0.05\n"
                                                       "sampling_initial_nums = 100\nsampling_max_iterations = 50\n"
                                                       "sampling_convergence_threshold_1 = 1e-5\nsampling_convergence_threshold_2 = 1e-5\n")
open(os.path.join(work, "metric.toml"), "w").write("m = 0.1\na = 0.0001\nrho = 1.0\n")
devices = ",".join(str(d) for d in range(n_dev))
common = ["video", os.path.join(work, "bg1.png"), os.path.join(work, "bg2.png")]
opts = ["-v", os.path.join(work, "video.toml"), "-c", os.path.join(work, "camera.toml"), "-s", os.path.join(work, "simulation.toml"),
        "-m", os.path.join(work, "metric.toml"), "--renderer", "per_pixel", "--precision", "f64_fast"]
result = {"config": "BASELINE configs[4]: path_through.csv, 15 fps, 3840x2160, Interstellar, 2000 / 0.05 / 45", "devices": n_dev, "frames": n_frames}
for label, extra in (("png_written", ["--encoder-threads", "16", "--compress-level", "1"]), ("render_only", ["--no-write"])):
    out = os.path.join(work, "out_" + label)
    os.mkdir(out)
    t0 = time.perf_counter()
    rc = main(common + [out] + opts + ["--devices", devices, "--frames", str(n_frames)] + extra)
    dt = time.perf_counter() - t0
    result[label] = {"rc": rc, "wall_s_including_scene_setup": round(dt, 2)}
    if label == "png_written":
        names = os.listdir(os.path.join(out, "tmp"))
        result[label]["files"] = len(names)
# the frames/s the product itself reports (the frame loop: render + encode, without decoding the backgrounds)
from curvis_b200 import settings as S
from curvis_b200.rendering import VideoRenderingSettings, VideoRenderingSystem, instantiate_metric
video = S.VideoSettings.from_toml_file(os.path.join(work, "video.toml"))
camera = S.CameraSettings.from_toml_file(os.path.join(work, "camera.toml"))
simulation = S.SimulationSettings.from_toml_file(os.path.join(work, "simulation.toml"))
metric = instantiate_metric(S.metric_settings_from_file(os.path.join(work, "metric.toml")))
for label, devs, write in (("all_devices_render_only", list(range(n_dev)), False), ("all_devices_png", list(range(n_dev)), True), ("one_device_render_only", [0], False)):
    out = os.path.join(work, "api_" + label)
    os.mkdir(out)
    settings = VideoRenderingSettings.from_settings(os.path.join(work, "bg1.png"), os.path.join(work, "bg2.png"), out, video, camera, simulation)
    system = VideoRenderingSystem(metric, settings, renderer="per_pixel", precision="f64_fast", devices=devs)
    frames = n_frames if len(devs) > 1 or n_frames <= 60 else 60
    system.render(max_frames=frames, verbose=False, encoder_threads=16, compress_level=1, write_frames=write)
    result[label] = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in system.last_render_info.items() if k != "render_s_per_device"}
    if label == "all_devices_png":
        many = {i: np.asarray(Image.open(os.path.join(out, "tmp", f"frame_{i}.png"))) for i in (0, n_frames // 2, n_frames - 1)}
    del system
# three frames of the multi-device run against a single device
out = os.path.join(work, "check")
os.mkdir(out)
settings = VideoRenderingSettings.from_settings(os.path.join(work, "bg1.png"), os.path.join(work, "bg2.png"), out, video, camera, simulation)
one = VideoRenderingSystem(metric, settings, renderer="per_pixel", precision="f64", devices=[0])
times = one.times_of_frames()
diff = {}
for i, frame in many.items():
    one.relativistic_system.camera = one.camera_at(times[i])
    ref = one.render_frame()
    diff[i] = int((ref != frame).any(axis=2).sum())
result["differing_pixels_vs_one_device_f64"] = diff
print(json.dumps(result))
if len(sys.argv) > 2:
    json.dump(result, open(sys.argv[2], "w"), indent=1)
