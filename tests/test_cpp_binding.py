"""The C++ drop-in header (include/och_h_octree_b200.hpp) compiles against the library and, on a GPU box,
the headless demo runs the reference's frame/edit flow."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_demo(tmp_path, ort, name="headless_demo"):
    exe = str(tmp_path / name)
    libdir = os.path.dirname(ort.LIB_PATH)
    cmd = [shutil.which("g++") or "g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", name + ".cpp"), "-L" + libdir, "-lort_b200", "-Wl,-rpath," + libdir, "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_header_compiles_and_fails_loudly_without_gpu(ort, tmp_path):
    import torch
    exe = build_demo(tmp_path, ort)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    out = subprocess.run([exe, "1"], capture_output=True, text=True)
    assert out.returncode == 2 and "no CUDA device" in out.stderr
    exe2 = build_demo(tmp_path, ort, "octree_demo")
    out = subprocess.run([exe2], capture_output=True, text=True)
    assert out.returncode == 2 and "no CUDA device" in out.stderr


@pytest.mark.gpu
def test_headless_demo_runs(ort, tmp_path):
    exe = build_demo(tmp_path, ort)
    out = subprocess.run([exe, "10"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("frame")]
    assert len(lines) == 10
    fills = {int(l.split("tabled nodes")[1].split()[0]) for l in lines}
    assert len(fills) > 1, "the T/Z edits should have changed the table"


@pytest.mark.gpu
def test_octree_demo_runs(ort, tmp_path):
    exe = build_demo(tmp_path, ort, "octree_demo")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    h0 = int(lines[0].split("hits")[1])
    h1 = int(lines[1].split("hits")[1])
    assert h0 > h1 > 0                       # one box was unset between the passes
    assert "dir 6 voxel 0 t 0" in lines[2]   # och::octree's MISS reports hit_time 0.0F
