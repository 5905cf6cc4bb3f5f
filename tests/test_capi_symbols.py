"""The C ABI: every function declared in include/mgb200.h is exported by the built libraries (no compute calls)."""
import ctypes as C
import os
import re

import pytest

import mgtest as T


def declared_symbols():
    txt = open(os.path.join(T.REPO, "include", "mgb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mgb?_[a-z0-9_]+)\s*\(", txt)))


def test_header_lists_the_reference_entry_points():
    syms = declared_symbols()
    for s in ("mg_index", "mg_idx_destroy", "mg_tbuf_init", "mg_tbuf_destroy", "mg_map", "mg_map_frag", "mg_map_batch", "mg_map_batch_frag", "mg_gchain_free"):
        assert s in syms


def test_hostsim_exports_all():
    lib = T.load_hostsim()
    for s in declared_symbols():
        assert hasattr(lib, s), s


def test_product_exports_all():
    path = os.path.join(T.REPO, "minigraph_b200", "libmgb200.so")
    if not os.path.exists(path):
        pytest.skip("libmgb200.so not built (run __graft_entry__.build())")
    out = os.popen("nm -D --defined-only %s" % path).read()
    for s in declared_symbols():
        assert re.search(r"\sT %s\b" % s, out), s


def test_product_refuses_without_gpu():
    """no CPU fallback: without a device mg_index() fails loudly"""
    path = os.path.join(T.REPO, "minigraph_b200", "libmgb200.so")
    if not os.path.exists(path):
        pytest.skip("libmgb200.so not built")
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from minigraph_b200 import capi, options
    lib = capi.load_product()
    g = lib.mgb_gfa_read(os.path.join(T.FIX, "MT.gfa").encode())
    io, mo = options.opt_set("lr")
    gi = lib.mg_index(g, C.byref(io), 1, C.byref(mo))
    assert not gi
    assert b"no CUDA device" in lib.mgb_last_error()
