"""The C-ABI library builds, loads, and exports every symbol include/b2r.h declares.
No compute calls (no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b2r.h")).read()
    return sorted(set(re.findall(r"B2R_API\s+[\w\s\*]+?\b(b2r_\w+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = _declared()
    for n in ["b2r_fps", "b2r_gather_fwd", "b2r_gather_bwd", "b2r_ball_query", "b2r_group_fwd",
              "b2r_group_bwd", "b2r_three_nn", "b2r_three_interp_fwd", "b2r_three_interp_bwd",
              "b2r_query_group_fwd", "b2r_query_group_bwd"]:
        assert n in names


def test_library_builds_loads_and_exports_all_symbols():
    from backtoreality_b200 import _lib, build
    path = build.build()
    assert os.path.isfile(path)
    lib = ctypes.CDLL(path)
    for n in _declared():
        assert hasattr(lib, n), "libb2r.so does not export %s" % n
    # every declared symbol has a ctypes signature in the Python binding, and vice versa
    assert sorted(_lib.SIGNATURES) == _declared()


def test_host_only_entry_points():
    from backtoreality_b200 import _lib
    from oracle import cpu_ops
    l = _lib.lib()
    assert l.b2r_version() >= 100
    assert l.b2r_status_string(0) == b"ok"
    for n in [1, 2, 3, 7, 8, 255, 256, 511, 512, 513, 1000, 1024, 2048, 40000, 100000]:
        assert l.b2r_ref_block_threads(n) == cpu_ops.block_threads(n)
    c, t, p, s = (ctypes.c_int() for _ in range(4))
    assert l.b2r_fps_plan(8, 40000, c, t, p, s) == 0
    assert c.value * t.value * p.value >= 40000 and 1 <= c.value <= 16  # any cluster size <= 16
    assert l.b2r_fps_plan(1, 10_000_000, c, t, p, s) == -3       # B2R_ERR_UNSUPPORTED
    assert b"capacity" in l.b2r_last_error()
    # the ctypes mirrors of the descriptor structs have the library's layout
    assert l.b2r_struct_bytes(0) == ctypes.sizeof(_lib.SaLayer)
    assert l.b2r_struct_bytes(1) == ctypes.sizeof(_lib.SaLayerBwd)
    assert l.b2r_struct_bytes(2) == ctypes.sizeof(_lib.DenseLayer)
    assert l.b2r_struct_bytes(3) == ctypes.sizeof(_lib.DenseLayerBwd)
    assert l.b2r_struct_bytes(7) == -1
    # the cluster hint of b2r_fps_ex is validated like every other argument
    assert l.b2r_fps_ex(None, 1, 10, 4, None, 17, None) == -1
    # argument validation happens before any CUDA call
    assert l.b2r_fps(None, 1, 10, 4, None, None) == -1
    assert l.b2r_ball_query(None, None, -1, 1, 1, 0.1, 1, None, None) == -1


def test_reference_import_path_shim_exposes_the_nine_functions():
    import pointnet2._ext as ext
    for n in ["gather_points", "gather_points_grad", "furthest_point_sampling", "three_nn",
              "three_interpolate", "three_interpolate_grad", "ball_query", "group_points",
              "group_points_grad"]:
        assert callable(getattr(ext, n))
