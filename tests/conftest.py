"""Shared fixtures.  `-m "not gpu"` covers the oracle against the reference-derived golden vectors,
the host logic, and that the C-ABI library loads and exports every declared symbol; `-m gpu` are the
parity tests proper (CUDA path vs oracle, through the C ABI)."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    need = [os.path.join(ROOT, "dpg_slam_b200", "libdpgicp.so"), os.path.join(ROOT, "dpg_slam_b200", "libdpgsynth.so"),
            os.path.join(ROOT, "oracle", "libdpgoracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g
        g.build()


@pytest.fixture(scope="session", autouse=True)
def built():
    _ensure_built()


def hex_f32(h):
    return np.array([int(x, 16) for x in h], np.uint32).view(np.float32)


def hex_f64(h):
    return np.array([int(x, 16) for x in h], np.uint64).view(np.float64)


@pytest.fixture(scope="session")
def cov_golden():
    with open(os.path.join(GOLDEN, "cov_ref.json")) as f:
        d = json.load(f)
    out = []
    for c in d["cases"]:
        out.append(dict(name=c["name"], singular=c["singular"], n_d_used=c["n_d_used"], sensor_var=c["sensor_var"],
                        live_in=c["live_in"],
                        P=hex_f32(c["P_hex"]).reshape(-1, 2), Q=hex_f32(c["Q_hex"]).reshape(-1, 2),
                        T_colmajor=hex_f32(c["T_colmajor_hex"]), live_cov=hex_f64(c["live_cov_hex"]).reshape(3, 3),
                        cov3=hex_f64(c["cov3_hex"]).reshape(3, 3), H3=hex_f64(c["H3_hex"]).reshape(3, 3)))
    return out


@pytest.fixture(scope="session")
def math_golden():
    with open(os.path.join(GOLDEN, "math_utils_ref.json")) as f:
        d = json.load(f)
    return dict(angle_in=hex_f32(d["angle_mod_in_hex"]), angle_out=hex_f32(d["angle_mod_out_hex"]),
                pose_pairs=hex_f32(d["pose_pairs_hex"]).reshape(-1, 6),
                inv_out=hex_f32(d["inverse_transform_out_hex"]).reshape(-1, 3),
                fwd_out=hex_f32(d["transform_out_hex"]).reshape(-1, 3))


@pytest.fixture(scope="session")
def gpu_matcher():
    """One ScanMatcher on cuda:0 for the whole GPU session (fails loudly if the library or GPU is missing)."""
    from dpg_slam_b200.scanmatch import ScanMatcher
    sm = ScanMatcher(0)
    yield sm
    sm.close()
