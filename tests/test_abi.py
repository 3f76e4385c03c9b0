"""CPU tier: the C-ABI library loads and exports exactly the symbols include/semseg_b200.h declares, and the
ctypes binding lists every one of them (no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

from semseg_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "semseg_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(semseg_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_symbols():
    syms = header_symbols()
    assert "semseg_conv_fprop" in syms and "semseg_psamask_fwd" in syms and len(syms) >= 20


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), "library does not export %s" % s


def test_no_undeclared_exports():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (semseg_[a-z0-9_]+)", out)))
    assert exported == header_symbols()


def test_ctypes_binding_covers_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()
    lib = _lib.load()
    assert lib.semseg_abi_version() == 1
    assert lib.semseg_launch_count() == 0 or lib.semseg_launch_count() > 0


def test_struct_layout_matches_header():
    """sizeof / offsets of the ABI structs as the C compiler sees them == the ctypes mirrors."""
    import tempfile
    prog = r'''
#include <stdio.h>
#include "semseg_b200.h"
#include <stddef.h>
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(semseg_conv_desc), sizeof(semseg_wgrad_desc), sizeof(semseg_pack_item),
         offsetof(semseg_pack_item, Cout), offsetof(semseg_pack_item, tile0), offsetof(semseg_pack_item, tiles_ci));
  return 0;
}
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        a, b, c, o_cout, o_tile0, o_tiles = map(int, subprocess.check_output([exe]).split())
    assert a == ctypes.sizeof(_lib.ConvDesc)
    assert b == ctypes.sizeof(_lib.WgradDesc)
    # the item table of semseg_pack_weights_multi is built by ctypes and read by the device code
    assert c == ctypes.sizeof(_lib.PackItem)
    assert (o_cout, o_tile0, o_tiles) == (_lib.PackItem.Cout.offset, _lib.PackItem.tile0.offset,
                                          _lib.PackItem.tiles_ci.offset)


def test_invalid_arguments_return_error_codes_without_gpu():
    lib = _lib.load()
    # argument validation happens before any CUDA call
    assert lib.semseg_psamask_fwd(3, None, None, 1, 1, 1, 1, 1, None) == -1
    assert b"psa_type" in lib.semseg_last_error()
    assert lib.semseg_psamask_fwd(0, ctypes.c_void_p(16), ctypes.c_void_p(16), 1, 2, 2, 4, 3, None) == -1
    assert b"odd" in lib.semseg_last_error()
    d = _lib.ConvDesc()
    assert lib.semseg_conv_fprop(ctypes.byref(d), None) == -1
    assert lib.semseg_conv_stats_rows(1, 8, 16, 64) == 4
    assert lib.semseg_iou_hist(ctypes.c_void_p(16), ctypes.c_void_p(16), 10, 0, 255, 1, ctypes.c_void_p(16), None) == -1
    assert b"iou_hist" in lib.semseg_last_error()
    assert lib.semseg_im2col3x3s2(ctypes.c_void_p(16), 8, 1, 9, 9, 4, ctypes.c_void_p(16), None) == -1
    assert b"im2col3x3s2" in lib.semseg_last_error()
    assert lib.semseg_pack_weights_multi(None, 1, 1, 9, None) == -1
