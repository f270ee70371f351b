"""CPU: the C-ABI library loads, exports every symbol include/dpgicp.h declares, mirrors the
reference's parameter defaults, refuses to run without a GPU (no CPU fallback), and its host-only
helper (the runIcp guess) matches the reference's math_utils bit for bit."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from dpg_slam_b200 import _abi
from dpg_slam_b200._abi import Params, Result

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dpgicp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dpgicp_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound():
    declared = _declared_symbols()
    assert len(declared) >= 20
    lib = _abi.load_library()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dpgicp.h but not exported"
    assert sorted(_abi.EXPORTS) == declared, "python binding list out of sync with the header"
    out = subprocess.run(["nm", "-D", "--defined-only", _abi.library_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (dpgicp_\w+)", out))
    assert set(declared) <= exported


def test_struct_layouts_match_header(tmp_path):
    assert C.sizeof(Params) == 96 and C.sizeof(Result) == 112
    assert Result.cov.offset == 40 and Result.mse.offset == 32 and Result.status.offset == 24
    assert _abi.RESULT_DTYPE.fields["cov"][1] == 40 and _abi.RESULT_DTYPE.itemsize == 112
    # the header itself, compiled as C: sizes and the offset of every params field against the ctypes mirror
    fields = [f[0] for f in Params._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dpgicp.h"\nint main(void) {\n'
                   '  printf("%zu %zu %zu %d\\n", sizeof(dpgicp_params), sizeof(dpgicp_result), sizeof(dpgicp_factor), DPGICP_ABI_VERSION);\n'
                   + "".join(f'  printf("{f} %zu\\n", offsetof(dpgicp_params, {f}));\n' for f in fields)
                   + "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[0].split() == [str(C.sizeof(Params)), "112", str(_abi.FACTOR_DTYPE.itemsize), str(_abi.ABI_VERSION)]
    for line, f in zip(out[1:], fields):
        name, off = line.split()
        assert name == f and int(off) == getattr(Params, f).offset, f


def test_default_params_are_the_reference_values():
    lib = _abi.load_library()
    p = Params()
    assert lib.dpgicp_default_params(C.byref(p)) == 0
    # parameters.h:146,159,173,191,201,374,385,396,402; cov_func_point_to_point.h:307,554
    assert (p.max_iterations, p.use_reciprocal, p.ransac_iterations, p.downsample_divisor) == (500, 1, 50, 5)
    assert (p.transformation_epsilon, p.max_correspondence_distance, p.cov_sensor_variance) == (5e-9, 0.6, 0.01)
    assert (p.laser_x_variance, p.laser_y_variance) == (0.5, 0.5) and p.laser_theta_variance == np.float32(0.3)
    assert p.cov_cap == 200 and p.cov_mode == _abi.COV_REFERENCE_LIVE and p.metric == _abi.METRIC_POINT_TO_POINT
    assert p.outlier_mode == _abi.OUTLIER_NONE and p.outlier_param == 0.0      # dpg_slam.cc:408-412 registers no rejector
    assert bytes(p) == bytes(Params.defaults())
    assert lib.dpgicp_default_params(None) == -1


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _abi.load_library()
    h = C.c_void_p()
    rc = lib.dpgicp_create(0, C.byref(h))
    assert rc == -2 and not h.value                                   # DPGICP_E_NODEVICE, never a CPU path
    assert b"no CPU fallback" in lib.dpgicp_last_error(None)
    from dpg_slam_b200.scanmatch import DpgIcpError, ScanMatcher
    with pytest.raises(DpgIcpError):
        ScanMatcher(0)
    # NULL context is rejected by every entry point
    assert lib.dpgicp_run(None, C.byref(Params.defaults())) == -1
    assert lib.dpgicp_scan_count(None) == -1


def test_relative_guess_matches_reference_math_utils(math_golden):
    from dpg_slam_b200.scanmatch import relative_guess
    for row, want in zip(math_golden["pose_pairs"], math_golden["inv_out"]):
        g = relative_guess(row[0:3], row[3:6])
        assert np.array_equal(g.view(np.uint32), want.view(np.uint32))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dpg_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".c", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                # comments may NAME the oracle (the shared arithmetic contract); code must not reach it
                assert not re.search(r'#\s*include\s*[<"][^>"]*oracle', text), f
                assert not re.search(r"^\s*(from|import)\s+\S*oracle", text, flags=re.M), f
                assert "libdpgoracle" not in text and "oracle_py" not in text and "-ldpgoracle" not in text, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        assert not re.search(r'#\s*include\s*[<"][^>"]*oracle', open(os.path.join(ROOT, "include", f)).read())


def test_batch_runner_fails_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    exe = os.path.join(ROOT, "dpg_slam_b200", "dpg_batch_runner")
    r = subprocess.run([exe, "--scans", "8", "--beams", "64"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr
