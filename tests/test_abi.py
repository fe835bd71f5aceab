"""CPU: the C-ABI library builds, loads, exports every symbol include/pf_b200.h declares, and
its argument checks / topology queries (no compute calls -- those need a GPU)."""
import ctypes as C
import os
import re

import pytest

from oracle import bg_oracle
from panoptic_forecasting_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(pf_lib):
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(pf_lib, s), "missing export %s" % s
    assert set(_lib.SIGNATURES) == set(syms)
    assert pf_lib.pf_version() >= 100


def test_workspace_queries_and_arg_checks(pf_lib):
    assert pf_lib.pf_zsplat_workspace_bytes(1, 3, 1024, 2048) >= 1024 * 2048 * 8
    assert pf_lib.pf_zsplat_workspace_bytes(0, 3, 4, 4) == 0
    net = C.c_void_p()
    assert pf_lib.pf_bgnet_create(C.byref(net), 11, 3, 1, 7) == -1          # bad precision -> PF_EINVAL
    assert b"precision" in pf_lib.pf_last_error()
    assert pf_lib.pf_bgnet_create(C.byref(net), 99, 3, 1, 0) == -1
    assert pf_lib.pf_bgnet_create(C.byref(net), 11, 3, 1, 0) == 0
    try:
        assert pf_lib.pf_bgnet_workspace_bytes(net, 1, 102, 128) == 0        # H not a multiple of 4
        assert pf_lib.pf_bgnet_workspace_bytes(net, 1, 128, 72) == 0         # W not a multiple of 16
        assert pf_lib.pf_bgnet_workspace_bytes(net, 1, 100, 176) > 0
        assert pf_lib.pf_bgnet_workspace_bytes(net, 1, 1024, 2048) > 100e6
        # null pointers are rejected before any CUDA call
        assert pf_lib.pf_bgnet_forward(net, None, None, None, 1, 64, 64, 64, 64, None, None, None, None, None, 0, None) == -1
        assert pf_lib.pf_zsplat_forward(*([None] * 8), 1, 1, 4, 4, 1, *([None] * 5), 0, None) == -1
        assert pf_lib.pf_bgnet_launches_per_forward(net) > 70
    finally:
        pf_lib.pf_bgnet_destroy(net)


def test_native_topology_matches_reference_channel_arithmetic(pf_lib):
    """conv names / shapes of the native plan vs hardnet.py:177-194,262-327 restated in the oracle."""
    net = C.c_void_p()
    assert pf_lib.pf_bgnet_create(C.byref(net), 11, 3, 1, 0) == 0
    try:
        n = pf_lib.pf_bgnet_num_convs(net)
        assert n == 69
        info = _lib.ConvInfo()
        got = {}
        for i in range(n + 1):
            assert pf_lib.pf_bgnet_conv_info(net, i, C.byref(info)) == 0
            got[info.name.decode()] = (info.cin, info.cout, info.ksize, info.stride)
        assert pf_lib.pf_bgnet_conv_info(net, n + 1, C.byref(info)) == -1
    finally:
        pf_lib.pf_bgnet_destroy(net)
    exp = {"model.base.0": (36, 16, 3, 2), "model.base.1": (16, 24, 3, 1), "model.base.2": (24, 32, 3, 2),
           "model.base.3": (32, 48, 3, 1)}
    ch, idx, skip = 48, 4, []
    for i in range(5):
        oc_total = 0
        for l in range(bg_oracle.N_LAYERS[i]):
            oc, ic, _ = bg_oracle.get_link(l + 1, ch, bg_oracle.GR[i], bg_oracle.GRMUL)
            exp["model.base.%d.layers.%d" % (idx, l)] = (ic, oc, 3, 1)
            if l % 2 == 0 or l == bg_oracle.N_LAYERS[i] - 1:
                oc_total += oc
        exp["model.base.%d" % (idx + 1)] = (oc_total, bg_oracle.CH_LIST[i], 1, 1)
        skip.append(oc_total)
        ch = bg_oracle.CH_LIST[i]
        idx += 3 if i < 4 else 2
    prev = ch
    for j in range(4):
        i = 3 - j
        cat = prev + skip[i]
        exp["model.conv1x1_up.%d" % j] = (cat, cat // 2, 1, 1)
        oc_total = 0
        for l in range(bg_oracle.N_LAYERS[i]):
            oc, ic, _ = bg_oracle.get_link(l + 1, cat // 2, bg_oracle.GR[i], bg_oracle.GRMUL)
            exp["model.denseBlocksUp.%d.layers.%d" % (j, l)] = (ic, oc, 3, 1)
            if l % 2 == 0 or l == bg_oracle.N_LAYERS[i] - 1:
                oc_total += oc
        prev = oc_total
    exp["model.finalConv"] = (prev, 11, 1, 1)
    assert got == exp


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.PFError):
        _lib.lib()
