"""The C-ABI library loads on a CPU-only box and exports exactly what include/mqvs.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from mediquery_rag_b200 import _native


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mqvs.h"), encoding="utf-8").read()
    return sorted(set(re.findall(r"MQVS_API[^;(]*?\b(mqvs_\w+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    names = declared_symbols()
    for must in ("mqvs_create", "mqvs_add", "mqvs_search", "mqvs_search_async", "mqvs_merge_candidates", "mqvs_count",
                 "mqvs_save", "mqvs_load", "mqvs_last_error", "mqvs_destroy", "mqvs_get_stats"):
        assert must in names


def test_library_exports_every_declared_symbol_and_binding_covers_them():
    lib = _native.load()
    names = declared_symbols()
    assert sorted(_native.SIGNATURES) == names
    for n in names:
        assert isinstance(getattr(lib, n), ctypes._CFuncPtr)
    assert lib.mqvs_version() == 200 and lib.mqvs_dim_pad() == 768


def test_the_wider_build_is_the_same_abi_and_the_binding_picks_the_narrowest_build_that_fits():
    """Row width is a build-time constant of the kernels: build() also makes libmqvs_w1024.so from the same sources.  Same
    symbols, its own MQVS_DIM_PAD; Index chooses by dimension and a dimension no build holds fails loudly."""
    assert _native.built_widths()[:2] == [768, 1024]
    wide = _native.load(1024)
    for n in declared_symbols():
        assert isinstance(getattr(wide, n), ctypes._CFuncPtr)
    assert wide.mqvs_dim_pad() == 1024 and wide.mqvs_version() == 200
    assert [_native.width_for(d) for d in (1, 384, 768, 769, 1000, 1024)] == [768, 768, 768, 1024, 1024, 1024]
    with pytest.raises(ValueError, match="WIDTH=1536"):
        _native.width_for(1500)
    h = ctypes.c_void_p()
    dev = (ctypes.c_int * 1)(0)
    assert _native.load().mqvs_create(1, dev, 1024, 0, 0, ctypes.byref(h)) == _native.E_INVALID   # the 768 build refuses
    assert b"WIDTH" in _native.load().mqvs_last_error()


def test_errors_are_codes_with_messages_not_crashes():
    lib = _native.load()
    h = ctypes.c_void_p(None)
    rc = lib.mqvs_create(2, None, 768, 0, 0, ctypes.byref(h))
    assert rc == _native.E_INVALID and b"dev_ids is NULL" in lib.mqvs_last_error()
    rc = lib.mqvs_create(_native.MAX_DEVICES + 1, None, 768, 0, 0, ctypes.byref(h))
    assert rc == _native.E_INVALID and b"n_dev" in lib.mqvs_last_error()
    rc = lib.mqvs_create(1, None, 4096, 0, 0, ctypes.byref(h))
    assert rc == _native.E_INVALID and b"dim" in lib.mqvs_last_error()
    assert lib.mqvs_search(None, None, 0, 1, 10, None, 0, None, None, 0) == _native.E_INVALID
    assert lib.mqvs_count(None, None) == _native.E_INVALID


def test_no_gpu_means_a_loud_error_not_a_fallback():
    from conftest import have_gpu
    if have_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(_native.MqvsError):
        _native.Index(768, "fp32")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mediquery-rag_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "liboracle" not in src, f


def test_header_is_plain_c_and_a_c_program_links_against_the_library(tmp_path):
    """include/mqvs.h compiles as strict C99 and examples/mqvs_demo.c (the reference's three call sites from plain C)
    links against libmqvs.so and runs: on a GPU box it must find its planted row, without one it must report the
    library's error and exit 2 -- never crash."""
    import shutil
    import subprocess
    from conftest import have_gpu
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "mqvs_demo")
    lib_dir = os.path.join(ROOT, "mediquery-rag_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "mqvs_demo.c"), "-L" + lib_dir, "-lmqvs", "-Wl,-rpath," + lib_dir, "-lm",
                    "-o", exe], check=True)
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert "libmqvs version 200" in run.stdout
    if have_gpu():
        assert run.returncode == 0 and "row 42" in run.stdout.splitlines()[2], run.stdout
    else:
        assert run.returncode == 2 and "no usable GPU" in run.stdout, run.stdout


def _sass_mnemonics(obj, pattern):
    import shutil
    import subprocess
    path = os.path.join(ROOT, "mediquery-rag_b200", "csrc", "build", obj)
    if shutil.which("cuobjdump") is None or not os.path.exists(path):
        pytest.skip("cuobjdump or the build directory is not available")
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    return set(re.findall(pattern, sass))


def test_the_hot_kernels_are_blackwell_native_sass():
    """What the built objects contain (cuobjdump -sass, sm_100a): the tensor path really issues tcgen05.mma from TMA-staged
    shared memory with TMEM accumulators (SASS names: UTCHMMA, UTMALDG, LDTM -- /opt/skills/guides/B200_PROFILING.md), the
    SIMT scan really streams rows with 128-bit no-allocate loads, and the cross-GPU exchange uses system-scope stores, loads
    and fences on peer memory."""
    tc = _sass_mnemonics("kernels_tc.o", r"\b(UTCHMMA|UTMALDG|LDTM|UTCBAR)\b")
    assert {"UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"} <= tc
    fused = _sass_mnemonics("kernels_tc_fused_space.o", r"\b(UTCHMMA|UTMALDG|LDTM)\b")
    assert {"UTCHMMA", "UTMALDG", "LDTM"} <= fused
    scan = _sass_mnemonics("kernels_scan.o", r"\bLDG\.E\.NA\.128\.CONSTANT\b|\bLDG\.E\.NA\.64\.CONSTANT\b")
    assert "LDG.E.NA.128.CONSTANT" in scan and "LDG.E.NA.64.CONSTANT" in scan
    xchg = _sass_mnemonics("kernels_xchg.o", r"\b(STG\.E\.STRONG\.SYS|LDG\.E\.STRONG\.SYS|MEMBAR\.\w+\.SYS)\b")
    assert {"STG.E.STRONG.SYS", "LDG.E.STRONG.SYS"} <= xchg and any(m.startswith("MEMBAR") for m in xchg)


def test_block_cyclic_placement_of_multi_device_handles_is_a_bijection():
    """mqvs_shard_of_row is pure host arithmetic in libmqvs.so (no GPU): every global row maps to one (shard, local row),
    shards fill densely in global order, and the inverse documented in include/mqvs.h holds."""
    blk = _native.SHARD_BLOCK
    for W in (1, 2, 3, 8):
        seen = {}
        counts = [0] * W
        for row in list(range(0, 3 * W * blk + 17, 97)) + [blk - 1, blk, W * blk - 1, W * blk, 2 ** 33 + 5]:
            shard, local = _native.shard_of_row(W, row)
            assert 0 <= shard < W
            assert ((local // blk) * W + shard) * blk + local % blk == row
            assert seen.setdefault((shard, local), row) == row
        for row in range(0, 2 * W * blk + 100):  # dense prefix: every shard's local rows are 0, 1, 2, ... in global order
            shard, local = _native.shard_of_row(W, row)
            if row % 61 == 0 or local % blk in (0, blk - 1):
                assert local <= counts[shard]
            counts[shard] = max(counts[shard], local + 1)
        assert sum(counts) == 2 * W * blk + 100 and max(counts) - min(counts) <= blk
