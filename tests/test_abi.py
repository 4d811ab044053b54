"""The C-ABI library loads on a CPU-only box and exports every symbol include/gort.h declares; without a
CUDA device it fails loudly instead of falling back (no compute calls here)."""
import ctypes
import os
import re

import pytest

import common as Cm


def header_functions():
    text = open(os.path.join(Cm.ROOT, "include", "gort.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(gort_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_symbols_exported(gort):
    lib = gort.load_library()
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libgort.so does not export %s" % n
    assert sorted(gort.EXPORTED_SYMBOLS) == names


def test_abi_version(gort):
    assert gort.load_library().gort_abi_version() == gort.ABI_VERSION == 3


def test_struct_sizes_match_header(gort):
    # offsets computed by the C compiler for the same declarations
    import subprocess
    import tempfile
    src = '#include "gort.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(gort_scene_desc), sizeof(gort_render_params), sizeof(gort_stats));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(Cm.ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(gort.SceneDesc), ctypes.sizeof(gort.RenderParams), ctypes.sizeof(gort.Stats)]


def test_no_cpu_fallback(gort):
    lib = gort.load_library()
    n = lib.gort_device_count()
    if n > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(gort.GortError) as e:
        gort.NewParallelRenderer(1)
    assert e.value.code == -3  # GORT_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_shard_slab_bytes(gort):
    assert gort.shard_slab_bytes(800, 600, 1) == 25 * 19 * 4096
    assert gort.shard_slab_bytes(800, 600, 8) == 60 * 4096  # ceil(475/8) tiles, padded equal for every rank
    assert gort.shard_slab_bytes(33, 1, 2) == 4096
