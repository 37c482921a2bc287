"""The device restatement of glibc's log() (flac_b200/csrc/encode_kernels.cuh: fb_log) must equal the
host libm's log() bit for bit: the reference calls the host's log() on its decision path
(lpc.c:1594, fixed.c:284), so this pins those decisions by construction."""
import ctypes as C
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_log(x):
    import flac_b200
    L = flac_b200.lib()
    L.fb200_debug_log.restype = C.c_int
    L.fb200_debug_log.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    assert L.fb200_debug_log(x.ctypes.data, y.ctypes.data, x.size, 0) == 0
    return y


def _host_log(x):
    # the C library's log(), the function libFLAC calls (numpy may use its own SIMD loops)
    libm = C.CDLL("libm.so.6")
    libm.log.restype = C.c_double
    libm.log.argtypes = [C.c_double]
    return np.array([libm.log(float(v)) for v in x], dtype=np.float64)


def test_log_bit_exact_against_host_libm():
    rng = np.random.default_rng(1)
    parts = [
        np.exp(rng.uniform(-700, 700, 200000)),                       # whole exponent range
        rng.uniform(0.9, 1.1, 200000),                                # around the near-1 branch boundaries
        1.0 + rng.uniform(-1, 1, 50000) * 2.0 ** rng.integers(-52, -3, 50000),
        rng.uniform(1e-3, 1e6, 200000),                               # what the encoder feeds (error_scale * lpc_error)
        np.array([1.0, 0.9375, 1.064697265625, 0.5, 2.0, 4096.0, 1e-300, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308]),
        np.ldexp(rng.uniform(0.5, 1.0, 20000), -1060).astype(np.float64),  # subnormals
    ]
    x = np.concatenate(parts)
    got = _device_log(x)
    want = _host_log(x)
    bad = np.nonzero(got.view(np.uint64) != want.view(np.uint64))[0]
    assert bad.size == 0, f"{bad.size} of {x.size} differ, e.g. x={x[bad[0]]!r}: device {got[bad[0]].hex()} host {want[bad[0]].hex()}"


def test_log_special_values():
    x = np.array([0.0, -0.0, -1.0, np.inf, np.nan])
    got = _device_log(x)
    assert got[0] == -np.inf and got[1] == -np.inf and math.isnan(got[2]) and got[3] == np.inf and math.isnan(got[4])
