"""CPU (-m "not gpu"): the C-ABI shared library builds, loads, and exports every symbol that
include/gtos_b200.h declares; host-side helpers behave; the product fails loudly without a GPU."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from gtos_b200 import _lib, build
    build.build()
    return _lib.load()


def test_header_symbols_are_exported_and_bound(lib):
    from gtos_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "gtos_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gtos_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version_and_error_channel(lib):
    from gtos_b200 import _lib
    assert lib.gtos_abi_version() == 5
    if not torch.cuda.is_available():
        assert lib.gtos_device_check() != 0
        assert "CUDA" in _lib.last_error() or "device" in _lib.last_error()


def test_rel_tiling_host_logic(lib):
    """tile chooser: bi*bj <= 128 rows, q/k staging budget, full coverage of the N x N grid."""
    from gtos_b200 import ops
    for N, B, D, H in [(17, 8, 128, 8), (41, 64, 512, 8), (61, 16, 512, 8), (257, 32, 512, 8), (1, 1, 128, 8)]:
        t = ops.rel_tiling(N, B, D, H)
        assert t["bi"] * t["bj"] <= 128
        bi8, bj8 = (t["bi"] + 7) // 8 * 8, (t["bj"] + 7) // 8 * 8
        full_rows = t["bj"] == N and t["bi"] <= 4             # every key of a query in one tile: the fused attention tail
        assert bi8 + bj8 <= 48 or (full_rows and (bi8 + 2 * bj8) * 512 <= 96 * 1024)
        fused_on = os.environ.get("GTOS_REL_FUSED_FWD", "0") == "1"
        assert full_rows == (fused_on and N in (41, 61, 1)) or N == 1, (N, t)   # configs 2 and 3 can fuse, the 257-node config cannot
        assert t["ni_blk"] * t["bi"] >= N and t["nj_blk"] * t["bj"] >= N
        assert t["tiles"] == B * t["ni_blk"] * t["nj_blk"]
        util = N * N / (t["ni_blk"] * t["nj_blk"] * 128)
        assert util > 0.5 or N < 8, (N, t, util)
    import ctypes as C
    out = (C.c_int32 * 5)()
    assert lib.gtos_rel_tiling(17, 8, 100, 4, out) != 0        # D % 128 != 0 -> refused, message set


def test_no_cpu_fallback():
    from gtos_b200 import _lib
    from gtos_b200.graph_transformer import GraphTransformer
    m = GraphTransformer(1, 128, 256, 8, 0.0)
    x, rel = torch.randn(5, 2, 128), torch.randn(5, 5, 2, 128)
    if not torch.cuda.is_available():
        with pytest.raises(_lib.GtosLibraryError):
            m(x, rel)


def test_state_dict_keys_match_reference_names():
    """parameter names / shapes the reference checkpoints use (SURVEY §5 checkpoint row)."""
    from gtos_b200.graph_transformer import GraphTransformer
    from gtos_b200.transformer import Transformer
    sd = GraphTransformer(1, 128, 256, 8, 0.1).state_dict()
    assert set(sd) == {"layers.0." + k for k in [
        "self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.relation_in_proj.weight",
        "self_attn.out_proj.weight", "self_attn.out_proj.bias", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias",
        "attn_layer_norm.weight", "attn_layer_norm.bias", "ff_layer_norm.weight", "ff_layer_norm.bias"]}
    assert sd["layers.0.self_attn.relation_in_proj.weight"].shape == (256, 128)
    sd = Transformer(1, 128, 256, 8, 0.1, with_external=True).state_dict()
    assert "layers.0.external_attn.in_proj_weight" in sd and "layers.0.external_layer_norm.bias" in sd


def test_precision_switch_host_logic():
    """ops.set_precision / ops.precision_mode: validation, nesting and restore (the modules read ops.fp32_mode() on every
    forward; GTOS_PRECISION sets the initial value)"""
    import pytest
    from gtos_b200 import ops
    start = ops.precision()
    assert start in ("bf16", "fp32")
    with ops.precision_mode("fp32"):
        assert ops.fp32_mode() and ops.precision() == "fp32"
        with ops.precision_mode("bf16"):
            assert not ops.fp32_mode()
        assert ops.fp32_mode()
    assert ops.precision() == start
    with pytest.raises(ValueError):
        ops.set_precision("fp16")
    with pytest.raises(ValueError):
        with ops.precision_mode("tf32"):
            pass
    assert ops.precision() == start
    # the fp32-mode Functions exist and keep the no-CPU-fallback rule
    import torch
    from gtos_b200 import _lib, ops32
    with pytest.raises(_lib.GtosLibraryError):
        ops32.split3(torch.zeros(4, 8), 0)
    with pytest.raises(_lib.GtosLibraryError):
        ops.zero_regions([torch.ones(4, 8)])
    ops.zero_regions([])                      # nothing to do: no launch, no error
