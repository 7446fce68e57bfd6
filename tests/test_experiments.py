"""The measurement build (libort_b200_exp.so = the product sources + -DORT_EXPERIMENTS): the kernels that were measured
and not adopted (csrc/ort_experiments.cuh) stay selectable there so that DESIGN.md's decisions can be re-measured -- and
they must stay the same function as the product kernels.  A library is chosen per process (ORT_B200_EXPERIMENTS), so the
GPU leg re-runs the variant tests of test_gpu_parity.py in a child process with the measurement build loaded."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_the_product_library_carries_no_experiment_kernel(ort):
    """nm over the product library: none of the experiment kernels' names; the measurement build has them."""
    from octree_ray_tracing_b200 import build
    build.build()
    names = ("trace_frame_tight_kernel", "trace_frame_pipe_kernel", "trace_frame_probe_kernel", "trace_frame_deferred_kernel",
             "trace_frame_staged_kernel", "trace_frame_tiles_kernel", "trace_frame_shaped_kernel", "trace_frame_walker_kernel",
             "trace_frame_walker_beam_kernel")
    prod = subprocess.run(["nm", "-C", build.LIB], capture_output=True, text=True, check=True).stdout
    exp = subprocess.run(["nm", "-C", build.LIB_EXP], capture_output=True, text=True, check=True).stdout
    for n in names:
        assert n not in prod, f"{n} is linked into the product library"
        assert n in exp, f"{n} is missing from the measurement build"
    assert b"experiments" not in ort.lib().ort_version() or os.environ.get("ORT_B200_EXPERIMENTS") == "1"


@pytest.mark.gpu
def test_experiment_kernels_equal_the_product_kernels():
    env = dict(os.environ, ORT_B200_EXPERIMENTS="1")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-x", "-q", "-m", "gpu",
                          "-k", "kernel_variants_agree or degenerate_rays_and_corner_cameras"], env=env, capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "2 passed" in out.stdout, out.stdout[-500:]
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_beam.py"), "-x", "-q", "-m", "gpu",
                          "-k", "beam_experiment_walkers"], env=env, capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and "1 passed" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]


@pytest.mark.gpu
def test_the_product_library_refuses_experiment_variants(ort):
    import numpy as np
    ctx = ort.TraceContext(2)
    nodes8 = np.zeros((2, 8), np.uint32); nodes8[0, :] = 2; nodes8[1, :] = 7
    ctx.upload_full(nodes8, 1)
    ctx.set_option("variant", 7)
    rot, fov = ort.camera_coeffs(0.3, -0.2)
    with pytest.raises(ort.OrtError, match="ORT_EXPERIMENTS"):
        ctx.trace_frame(np.array([1.2, 1.3, 1.4], np.float32), rot, fov, 64, 32)
    ctx.set_option("variant", 13)
    v, f, t = ctx.trace_frame(np.array([1.2, 1.3, 1.4], np.float32), rot, fov, 64, 32)
    assert (f == 7).all() and (v == 7).all()            # the origin sits inside a solid voxel
    # a successful call leaves no stale message behind
    assert not ort.lib().ort_last_error(ctx.h)
