"""The C-ABI library loads and exports every symbol include/sihl_od.h declares (no compute calls: no GPU here)."""
import ctypes as C
import os
import re

from sihl_b200 import _native, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "sihl_od.h")).read()
    return sorted(set(re.findall(r"SIHL_OD_API\s+[\w\s\*]+?\b(sihl_od_\w+)\s*\(", text)))


def test_library_builds_and_loads():
    path = build.build()
    assert os.path.exists(path)
    lib = _native.load()
    assert lib.sihl_od_version() >= 100
    assert lib.sihl_od_last_error_string() is not None


def test_every_declared_symbol_is_exported_and_bound():
    declared = _declared()
    assert len(declared) >= 20
    lib = C.CDLL(build.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in sihl_od.h but not exported"
    assert sorted(_native.EXPORTED_SYMBOLS) == declared, "ctypes signatures and header disagree"


def test_argument_errors_do_not_need_a_gpu():
    lib = _native.load()
    # n_levels out of range is rejected on the host, before any CUDA call
    rc = lib.sihl_od_anchors(None, 0, 640, 640, None, None, None, None)
    assert rc == 1
    assert b"level" in lib.sihl_od_last_error_string()
    rc = lib.sihl_od_topk(None, 1, 100, 5, None, None, None)
    assert rc == 1
    assert lib.sihl_od_nms_workspace_bytes(64, 256) == 0
    assert lib.sihl_od_nms_workspace_bytes(64, 8525) > 0


def test_sm100a_only_and_blackwell_bulk_copy_in_sass():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    out = subprocess.run([cuobjdump, "-lelf", build.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
    sass = subprocess.run([cuobjdump, "-sass", build.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass      # cp.async.bulk + mbarrier (TMA-class bulk copies)


def test_ops_reject_cpu_tensors():
    import pytest
    import torch

    from sihl_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.topk_locations(torch.zeros(2, 300), 100)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.dense_loss(torch.zeros(10), None, torch.zeros(10), torch.zeros(8, dtype=torch.float64))


def test_exchange_and_split_helpers_without_a_gpu():
    """Size queries and argument checks of the multi-GPU exchange / class-split NMS entries are host-only."""
    lib = _native.load()
    assert lib.sihl_od_exchange_region_bytes(0) == 0 and lib.sihl_od_exchange_region_bytes(17) == 0
    assert lib.sihl_od_exchange_region_bytes(8) == 1280                # (18 * 8 + 1) words of 8 bytes, rounded to 256
    assert lib.sihl_od_exchange_region_bytes(1) == 256
    assert lib.sihl_od_nms_split_workspace_bytes(0, 100, 100) == 0
    small, big = lib.sihl_od_nms_split_workspace_bytes(1, 8525, 100), lib.sihl_od_nms_split_workspace_bytes(16, 34100, 100)
    assert 0 < small < big
    rc = lib.sihl_od_pos_loss_tiles_exchange(None, None, None, 1, 100, None, None, 10, 10, None, None, None, None, None, 0,
                                             None, None, None, None, 99, 0, None)
    assert rc == 1 and b"world" in lib.sihl_od_last_error_string()
    rc = lib.sihl_od_quad_matching(None, 10, None, None, 1, 5, 0, None, None, None, None, None, None, None, None)
    assert rc == 1 and b"topk" in lib.sihl_od_last_error_string()


def test_header_is_plain_c99():
    """include/sihl_od.h is the boundary a non-C++ host would bind: it must compile as C."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        return
    header = os.path.join(ROOT, "include", "sihl_od.h")
    res = subprocess.run([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", header], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_integration_stub_matches_the_abi():
    """The ctypes stub shown in INTEGRATION.md declares as many arguments as the library takes."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for name in ("sihl_od_assign_select", "sihl_od_assign_resolve"):
        m = re.search(r"_lib\." + name + r"\.argtypes = \[(.*?)\]", text, re.S)
        assert m, name
        n = len([x for x in m.group(1).replace("\n", " ").split(",") if x.strip()])
        assert n == len(_native._SIGNATURES[name][1]), (name, n)
