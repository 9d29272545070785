"""The C ABI from plain C: tests/c_abi_smoke.c is compiled with gcc against include/cub200.h + libcub200.so + cudart
(no torch, no C++) and must reproduce the reference's LV example mask (11 961 pixels)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import port

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_consumer(tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    gcc = shutil.which("gcc")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if gcc is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc / CUDA headers not available")
    libdir = os.path.join(ROOT, "contouring_uncertainty_b200", "lib")
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.check_call([gcc, "-std=c99", "-O1", os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-I", os.path.join(cuda, "include"),
                           "-L", libdir, "-L", os.path.join(cuda, "lib64"), "-lcub200", "-lcudart", "-Wl,-rpath," + libdir,
                           "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    words = out.stdout.split()
    vals = {words[i]: words[i + 1] for i in (0, 2)}
    assert vals["sm"] == "100"
    areas = [int(w) for w in words[5:8]]
    lv = port.LV_EXAMPLE_SHAPE
    want = [int(port.us_contour_to_mask(lv + np.array([i, 0.0])).sum()) for i in range(3)]
    assert areas == want and areas[0] == 11961
    count_sum, pred_sum = int(words[9]), int(words[11])
    masks = np.stack([port.us_contour_to_mask(lv + np.array([i, 0.0])) for i in range(3)])
    assert count_sum == int(masks.sum()) and pred_sum == int((masks.mean(0).round()).sum())
    assert "einval -1 empty 0" in out.stdout
    line = [l for l in out.stdout.splitlines() if l.startswith("lv_length")][0].split()
    lengths = [float(v) for v in line[1:4]]
    want_len = [port.lv_longitudinal_length(m.astype(bool)) for m in masks]
    np.testing.assert_allclose(lengths, want_len, rtol=1e-11)
    assert [int(v) for v in line[5:8]] == want
