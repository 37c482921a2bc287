"""ctypes bindings for the prebuilt reference libFLAC (oracle/_ref/*.so, built by
oracle/Makefile from /root/reference) -- test infrastructure only."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "..", "oracle", "_ref")


class RefEncOpts(C.Structure):
    _fields_ = [
        ("exhaustive", C.c_int32), ("mid_side", C.c_int32), ("loose_mid_side", C.c_int32),
        ("max_lpc_order", C.c_int32), ("qlp_precision", C.c_int32), ("min_part_order", C.c_int32),
        ("max_part_order", C.c_int32), ("disable_isa", C.c_int32), ("streamable_subset", C.c_int32),
        ("limit_min_bitrate", C.c_int32), ("prec_search", C.c_int32), ("apodization", C.c_char_p),
    ]

    def __init__(self, **kw):
        super().__init__(-1, -1, -1, -1, -1, -1, -1, 0, -1, -1, -1, None)
        for k, v in kw.items():
            if k == "apodization" and isinstance(v, str):
                v = v.encode()
            setattr(self, k, v)


_libs = {}


def available(variant="default"):
    return os.path.exists(_path(variant))


def _path(variant):
    name = {"default": "libFLAC_ref.so", "strict": "libFLAC_ref_strict.so"}[variant]
    return os.path.join(REF_DIR, name)


def lib(variant="default"):
    if variant not in _libs:
        L = C.CDLL(_path(variant))
        L.ref_encode.restype = C.c_int
        L.ref_encode.argtypes = [
            C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
            C.POINTER(RefEncOpts), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
            C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
        L.ref_decode.restype = C.c_int
        L.ref_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64),
                                 C.POINTER(C.c_uint32 * 4), C.c_int]
        L.ref_version.restype = C.c_char_p
        if hasattr(L, "ref_encode_parallel"):
            L.ref_encode_parallel.restype = C.c_int
            L.ref_encode_parallel.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                              C.c_uint32, C.POINTER(RefEncOpts), C.POINTER(C.c_double), C.POINTER(C.c_uint64),
                                              C.POINTER(C.c_uint64)]
        _libs[variant] = L
    return _libs[variant]


def encode(pcm, bps, rate=44100, level=5, blocksize=0, threads=1, md5=False, variant="default", opts=None,
           want_bytes=True):
    """pcm: int32 [samples, channels]. Returns (stream_bytes, header_len, [frame_bytes...])."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int32)
    n, ch = pcm.shape
    L = lib(variant)
    cap = 1 << 16
    cap += int(n * ch * (bps // 8 + 2) * 1.1) + 64 * (n // 16 + 2)
    out = np.empty(cap, dtype=np.uint8) if want_bytes else None
    max_frames = n // 16 + 4 if blocksize and blocksize < 256 else n // 192 + 4
    fs = np.zeros(max_frames, dtype=np.uint32)
    out_len, hdr_len, nfr, aux = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), C.c_int(0)
    rc = L.ref_encode(pcm.ctypes.data, n, ch, bps, rate, level, blocksize, threads, int(md5),
                      C.byref(opts) if opts is not None else None,
                      out.ctypes.data if out is not None else None, cap, C.byref(out_len), C.byref(hdr_len),
                      fs.ctypes.data, max_frames, C.byref(nfr), C.byref(aux))
    if rc != 0:
        raise RuntimeError(f"ref_encode failed rc={rc} aux={aux.value}")
    assert nfr.value <= max_frames
    if not want_bytes:
        return None, hdr_len.value, fs[:nfr.value].copy()
    stream = out[:out_len.value]
    sizes = fs[:nfr.value]
    frames = []
    pos = hdr_len.value
    for s in sizes:
        frames.append(stream[pos:pos + int(s)].tobytes())
        pos += int(s)
    assert pos == out_len.value
    return stream.tobytes(), hdr_len.value, frames


def encode_parallel(pcm, bps, rate, level, blocksize, workers, variant="default", opts=None):
    """One reference encoder per host thread over contiguous block ranges (bytes discarded).
    pcm: int32 [nblocks*blocksize, channels]. Returns (seconds, frames, frame_bytes)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int32)
    n, ch = pcm.shape
    assert n % blocksize == 0
    sec, nfr, nby = C.c_double(0), C.c_uint64(0), C.c_uint64(0)
    rc = lib(variant).ref_encode_parallel(pcm.ctypes.data, n // blocksize, ch, bps, rate, level, blocksize, workers,
                                          C.byref(opts) if opts is not None else None, C.byref(sec), C.byref(nfr), C.byref(nby))
    if rc != 0:
        raise RuntimeError(f"ref_encode_parallel failed rc={rc}")
    return sec.value, nfr.value, nby.value


def decode(stream, max_samples, channels, variant="default", md5=False):
    """Returns (pcm int32 [samples, channels], info(channels,bps,rate,errors))."""
    buf = np.frombuffer(stream, dtype=np.uint8)
    out = np.zeros((max_samples, channels), dtype=np.int32)
    ns = C.c_uint64(0)
    info = (C.c_uint32 * 4)()
    rc = lib(variant).ref_decode(buf.ctypes.data, buf.size, out.ctypes.data, max_samples, C.byref(ns), C.byref(info), int(md5))
    if rc != 0:
        raise RuntimeError(f"ref_decode failed rc={rc}")
    return out[:ns.value], tuple(info)
