"""The C-ABI library loads and exports every symbol include/rc_b200.h declares; struct layouts match
the header (compiled with gcc at test time).  No compute calls: this runs without a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

import radiancecascade_b200 as rc
from radiancecascade_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rc_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rc_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _ffi.load()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in rc_b200.h but not exported by librc_b200.so"
        assert n in _ffi.SYMBOLS, f"{n} has no ctypes signature in _ffi.SYMBOLS"
    for n in _ffi.SYMBOLS:
        assert n in names, f"{n} bound in _ffi but not declared in the header"
    assert lib.rc_abi_version() == 1


def test_struct_layouts_match_the_header():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "rc_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(rc_config), sizeof(rc_camera), sizeof(rc_light),
         sizeof(rc_level_info), sizeof(rc_scene_info), offsetof(rc_config, scene_path), offsetof(rc_config, tile_x0),
         offsetof(rc_level_info, texel_offset));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        got = [int(x) for x in subprocess.check_output([exe]).split()]
    want = [C.sizeof(_ffi.rc_config), C.sizeof(_ffi.rc_camera), C.sizeof(_ffi.rc_light), C.sizeof(_ffi.rc_level_info),
            C.sizeof(_ffi.rc_scene_info), _ffi.rc_config.scene_path.offset, _ffi.rc_config.tile_x0.offset,
            _ffi.rc_level_info.texel_offset.offset]
    assert got == want
    assert C.sizeof(_ffi.rc_camera) == 80 and C.sizeof(_ffi.rc_light) == 16   # UniformCamera / UniformLight (src/camera.rs:9-15, src/primitives.rs:14-18)


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="needs a box WITHOUT a CUDA device")
def test_create_fails_loudly_without_a_device():
    """No CPU fallback: without CUDA rc_create must refuse, not render on the host."""
    with pytest.raises(rc.RcError) as e:
        rc.DefaultRenderer.new(0, (64, 64), rc.AppState(), rc.scenes.scene_path("cube"))
    assert e.value.status == _ffi.RC_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_bad_arguments_return_status_codes():
    lib = _ffi.load()
    h = C.c_void_p()
    assert lib.rc_create(None, C.byref(h)) == _ffi.RC_ERR_INVALID_ARG
    cfg = _ffi.rc_config()
    cfg.struct_size = 4   # wrong size: ABI mismatch must be caught
    assert lib.rc_create(C.byref(cfg), C.byref(h)) == _ffi.RC_ERR_INVALID_ARG
    assert lib.rc_update(None, None, None, 0, 0) == _ffi.RC_ERR_INVALID_ARG
    assert lib.rc_render(None, None) == _ffi.RC_ERR_INVALID_ARG
    s = C.c_void_p()
    assert lib.rc_scene_load(b"/nonexistent.obj", 0, C.byref(s)) == _ffi.RC_ERR_SCENE_LOAD   # ≙ .unwrap() panic, src/renderer.rs:176
    assert b"cannot open" in lib.rc_last_error(None)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under radiancecascade_b200/ may reference it."""
    pkg = os.path.join(ROOT, "radiancecascade_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(d, f), errors="replace").read()
                assert "oracle" not in text.replace("rc_oracle.c with OpenMP", ""), os.path.join(d, f)


def test_div32767_sequence_is_the_correctly_rounded_division(tmp_path):
    """rc_device.cuh div32767 (oct_decode's division without the generic slow path) == a / 32767.0f for all int16 a."""
    import subprocess
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "div32767_check.c")
    exe = str(tmp_path / "div32767_check")
    subprocess.check_call(["gcc", "-O1", "-ffp-contract=off", "-o", exe, src, "-lm"])
    assert subprocess.run([exe], capture_output=True, text=True).stdout.strip() == "0"
