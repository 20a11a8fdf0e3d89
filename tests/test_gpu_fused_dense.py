"""-m gpu: the single-kernel dense relation attention (gtos_rel_attn_fwd: projection GEMM + scores + key-padding mask +
softmax + dropout + P.V in the tcgen05 score kernel's epilogue; graph_transformer.py:122-159).  It is opt-in
(GTOS_REL_FUSED_FWD=1 switches both the tile chooser - full-row tiles - and the module path; it measured slower than the
two-kernel default), and the switch is read once per process, so the module / full-size / generator-level
parity tests are re-run here in a child process with the switch on."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parity_suite_with_the_fused_dense_kernel():
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    env = dict(os.environ, GTOS_REL_FUSED_FWD="1")
    probe = subprocess.run([sys.executable, "-c",
                            "from gtos_b200 import _lib; l = _lib.load(); "
                            "print(l.gtos_rel_attn_fusable(41, 64, 512, 8), l.gtos_rel_attn_fusable(61, 16, 512, 8), "
                            "l.gtos_rel_attn_fusable(257, 32, 512, 8))"], cwd=ROOT, env=env, capture_output=True, text=True)
    assert probe.stdout.split() == ["1", "1", "0"], probe.stdout + probe.stderr
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "tests/test_gpu_modules.py", "tests/test_gpu_fullsize.py",
                        "tests/test_gpu_edge.py", "tests/test_gpu_generator.py", "-k",
                        "graph_transformer or banked or padding or permutation or one_graph or config3 or hot_path or generator"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout
