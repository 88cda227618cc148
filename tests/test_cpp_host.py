"""The C++ host mirror (curvis_b200/host/curvis.hpp) compiles, links against the C-ABI library,
refuses to run without a GPU, and on a GPU renders the same frame as the Python mirror."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "curvis_b200", "host", "curvis_image")


def test_cpp_driver_builds_and_fails_loudly_without_gpu(built, tmp_path):
    import torch
    assert os.path.exists(EXE)
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    r = subprocess.run([EXE, "ellis", "16", "9", "10", "100", "0.05", str(tmp_path / "x.ppm")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_driver_matches_oracle(built, oracle, tmp_path):
    from curvis_b200 import scenes
    out = tmp_path / "frame.ppm"
    r = subprocess.run([EXE, "ellis", "64", "36", "200", "10", "0.1", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = out.read_bytes()
    header = b"P6\n64 36\n255\n"
    assert raw.startswith(header)
    frame = np.frombuffer(raw[len(header):], dtype=np.uint8).reshape(36, 64, 3)
    cam = oracle.camera(scenes.DEFAULT_CAMERA_POSITION, scenes.DEFAULT_FORWARD, scenes.DEFAULT_UP, 15.0, 43.0, 64, 36)
    bp, bn = scenes.decodable_background(1024, 512), scenes.decodable_background(1024, 512, True)
    ref, _, st = oracle.render_rows(oracle.metric("ellis"), cam, oracle.sim(200, 10.0, 0.1), bp, bn, with_records=False)
    assert (frame == ref).all()
    assert f"steps={st['total_steps']} " in r.stdout
