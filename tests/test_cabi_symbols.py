"""CPU-side checks of the drop-in boundary: the shared library loads (no GPU needed for that) and
exports every function include/*.h declares; creating an engine without a CUDA device fails
loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set()
    for m in re.finditer(r"\b((?:fb200|FLAC__stream_(?:encoder|decoder))_\w+)\s*\(", src):
        names.add(m.group(1))
    return sorted(names)


@pytest.fixture(scope="module")
def lib():
    import flac_b200
    from flac_b200 import build
    build.build()
    return flac_b200.lib()


@pytest.mark.parametrize("header", ["flac_b200.h", "flac_b200_stream.h"])
def test_every_declared_symbol_is_exported(lib, header):
    names = _declared_functions(header)
    assert len(names) > 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"{header}: not exported: {missing}"


def test_string_tables_exported(lib):
    for name in ("FLAC__StreamEncoderStateString", "FLAC__StreamEncoderInitStatusString", "FLAC__StreamDecoderStateString",
                 "FLAC__VERSION_STRING", "FLAC__VENDOR_STRING"):
        assert hasattr(lib, name), name
    arr = (C.c_char_p * 9).in_dll(lib, "FLAC__StreamEncoderStateString")
    assert arr[0] == b"FLAC__STREAM_ENCODER_OK" and arr[8] == b"FLAC__STREAM_ENCODER_MEMORY_ALLOCATION_ERROR"


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import flac_b200
    with pytest.raises(flac_b200.FlacB200Error) as ei:
        flac_b200.Encoder(flac_b200.preset(2, 16, 44100, 5))
    assert ei.value.code == -1  # FB200_ERR_CUDA
    with pytest.raises(flac_b200.FlacB200Error):
        flac_b200.Decoder(2, 16, 44100, 4096)
    # the object API reports it through the reference's own error channel
    lib.FLAC__stream_encoder_new.restype = C.c_void_p
    e = lib.FLAC__stream_encoder_new()
    WRITE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p)
    cb = WRITE(lambda *a: 0)
    lib.FLAC__stream_encoder_init_stream.argtypes = [C.c_void_p, WRITE, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    st = lib.FLAC__stream_encoder_init_stream(e, cb, None, None, None, None)
    assert st == 1  # FLAC__STREAM_ENCODER_INIT_STATUS_ENCODER_ERROR
    lib.FLAC__stream_encoder_delete.argtypes = [C.c_void_p]
    lib.FLAC__stream_encoder_delete(e)


def test_setters_and_validation_without_device(lib):
    """Setter/getter state machine and init validation order (reference: src/test_libFLAC/encoders.c)."""
    L = lib
    L.FLAC__stream_encoder_new.restype = C.c_void_p
    e = C.c_void_p(L.FLAC__stream_encoder_new())
    for fn in ("set_channels", "set_bits_per_sample", "set_sample_rate", "set_compression_level", "set_blocksize", "set_max_lpc_order"):
        getattr(L, "FLAC__stream_encoder_" + fn).argtypes = [C.c_void_p, C.c_uint32]
    L.FLAC__stream_encoder_get_state.argtypes = [C.c_void_p]
    assert L.FLAC__stream_encoder_get_state(e) == 1  # UNINITIALIZED
    assert L.FLAC__stream_encoder_set_channels(e, 9)
    WRITE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p)
    cb = WRITE(lambda *a: 0)
    L.FLAC__stream_encoder_init_stream.argtypes = [C.c_void_p, WRITE, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    assert L.FLAC__stream_encoder_init_stream(e, cb, None, None, None, None) == 4   # INVALID_NUMBER_OF_CHANNELS
    L.FLAC__stream_encoder_set_channels(e, 2)
    L.FLAC__stream_encoder_set_bits_per_sample(e, 3)
    assert L.FLAC__stream_encoder_init_stream(e, cb, None, None, None, None) == 5   # INVALID_BITS_PER_SAMPLE
    L.FLAC__stream_encoder_set_bits_per_sample(e, 16)
    L.FLAC__stream_encoder_set_blocksize(e, 8)
    assert L.FLAC__stream_encoder_init_stream(e, cb, None, None, None, None) == 7   # INVALID_BLOCK_SIZE
    L.FLAC__stream_encoder_set_blocksize(e, 8192)
    assert L.FLAC__stream_encoder_init_stream(e, cb, None, None, None, None) == 11  # NOT_STREAMABLE (subset: <=4608 at 44.1k)
    L.FLAC__stream_encoder_set_blocksize(e, 0)
    L.FLAC__stream_encoder_set_max_lpc_order(e, 33)
    assert L.FLAC__stream_encoder_init_stream(e, cb, None, None, None, None) == 8   # INVALID_MAX_LPC_ORDER
    NOCB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p)
    L.FLAC__stream_encoder_set_max_lpc_order(e, 8)
    assert L.FLAC__stream_encoder_init_stream(e, C.cast(None, NOCB), None, None, None, None) == 3  # INVALID_CALLBACKS
    L.FLAC__stream_encoder_init_ogg_stream.argtypes = [C.c_void_p] + [C.c_void_p] * 6
    assert L.FLAC__stream_encoder_init_ogg_stream(e, None, None, None, None, None, None) == 2      # UNSUPPORTED_CONTAINER
    L.FLAC__stream_encoder_delete.argtypes = [C.c_void_p]
    L.FLAC__stream_encoder_delete(e)


def test_crc16_power_table_constants():
    """kCrcXPow2 in device_common.cuh (x^(2^j) mod x^16+x^15+x^2+1, used to combine chunk CRCs) recomputed from
    scratch, and the combine identity crc(A||B) = crc(A) * x^(8|B|) + crc(B) checked against a bytewise CRC-16."""
    import re
    src = open(os.path.join(os.path.dirname(__file__), "..", "flac_b200", "csrc", "device_common.cuh")).read()
    m = re.search(r"kCrcXPow2\[15\]\s*=\s*\{([^}]*)\}", src)
    table = [int(v, 16) for v in m.group(1).replace(" ", "").split(",")]

    def mul(a, b):
        r = 0
        for i in range(15, -1, -1):
            r = ((r << 1) ^ 0x8005) & 0xFFFF if r & 0x8000 else (r << 1)
            if (b >> i) & 1:
                r ^= a
        return r

    v, want = 2, []
    for _ in range(15):
        want.append(v)
        v = mul(v, v)
    assert table == want and v == 2  # x has order 2^15 - 1: the table is periodic

    def crc16(data, crc=0):
        for byte in data:
            crc ^= byte << 8
            for _ in range(8):
                crc = ((crc << 1) ^ 0x8005) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
        return crc

    def xpow(e):
        r, j = 1, 0
        while e:
            if e & 1:
                r = mul(r, table[j % 15])
            e >>= 1
            j += 1
        return r

    rng = np.random.default_rng(3)
    a, b = bytes(rng.integers(0, 256, 517, dtype=np.uint8)), bytes(rng.integers(0, 256, 4099, dtype=np.uint8))
    assert crc16(a + b) == mul(crc16(a), xpow(8 * len(b))) ^ crc16(b)
