"""ctypes bindings for oracle/liboracle.so (the CPU restatement) -- test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.abspath(os.path.join(_HERE, "..", "oracle"))
FO_MAX_LPC_ORDER = 32
FO_MAX_PARTITIONS = 256


class Apod(C.Structure):
    _fields_ = [("type", C.c_int32), ("p", C.c_float), ("parts", C.c_int32), ("start", C.c_float), ("end", C.c_float)]


class Config(C.Structure):
    _fields_ = [
        ("channels", C.c_uint32), ("bits_per_sample", C.c_uint32), ("sample_rate", C.c_uint32), ("blocksize", C.c_uint32),
        ("do_mid_side", C.c_int32), ("loose_mid_side", C.c_int32),
        ("max_lpc_order", C.c_uint32), ("qlp_coeff_precision", C.c_uint32),
        ("do_qlp_coeff_prec_search", C.c_int32), ("do_exhaustive_model_search", C.c_int32),
        ("min_residual_partition_order", C.c_uint32), ("max_residual_partition_order", C.c_uint32),
        ("num_apodizations", C.c_uint32), ("apodizations", Apod * 32),
        ("disable_constant_subframes", C.c_int32), ("disable_fixed_subframes", C.c_int32),
        ("disable_verbatim_subframes", C.c_int32), ("limit_min_bitrate", C.c_int32), ("x86_avx2_fixed_guess", C.c_int32),
    ]


class SubframePlan(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("order", C.c_int32), ("wasted_bits", C.c_int32), ("subframe_bps", C.c_int32),
        ("qlp_precision", C.c_int32), ("qlp_shift", C.c_int32), ("qlp_coeff", C.c_int32 * FO_MAX_LPC_ORDER),
        ("rice_method", C.c_int32), ("partition_order", C.c_int32), ("rice_params", C.c_int32 * FO_MAX_PARTITIONS),
        ("estimate_bits", C.c_uint32),
    ]


class FramePlan(C.Structure):
    _fields_ = [("channel_assignment", C.c_int32), ("sub", SubframePlan * 8), ("cand", SubframePlan * 4),
                ("cand_valid", C.c_uint32 * 4)]


class StreamInfo(C.Structure):
    _fields_ = [("channels", C.c_uint32), ("bits_per_sample", C.c_uint32), ("sample_rate", C.c_uint32)]


_lib = None


def build():
    subprocess.run(["make", "-s", "oracle"], cwd=ORACLE_DIR, check=True)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.fo_config_preset.argtypes = [C.POINTER(Config), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.fo_encoder_new.restype = C.c_void_p
        L.fo_encoder_new.argtypes = [C.POINTER(Config)]
        L.fo_encoder_delete.argtypes = [C.c_void_p]
        L.fo_encoder_config.restype = C.POINTER(Config)
        L.fo_encoder_config.argtypes = [C.c_void_p]
        L.fo_encode_frame.restype = C.c_size_t
        L.fo_encode_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(FramePlan)]
        L.fo_encode_stream.restype = C.c_int
        L.fo_encode_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                       C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.fo_decode_frame.restype = C.c_size_t
        L.fo_decode_frame.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(StreamInfo), C.c_void_p, C.c_size_t,
                                      C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.fo_crc8.restype = C.c_uint8
        L.fo_crc8.argtypes = [C.c_void_p, C.c_size_t]
        L.fo_crc16.restype = C.c_uint16
        L.fo_crc16.argtypes = [C.c_void_p, C.c_size_t]
        L.fo_window_tukey.argtypes = [C.c_void_p, C.c_int32, C.c_float]
        L.fo_config_set_apodization.argtypes = [C.POINTER(Config), C.c_char_p]
        L.fo_window.restype = C.c_int
        L.fo_window.argtypes = [C.POINTER(Apod), C.c_void_p, C.c_int32]
        L.fo_autocorrelation.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


def preset(channels, bps, rate, level, blocksize=0, **over):
    cfg = Config()
    lib().fo_config_preset(C.byref(cfg), channels, bps, rate, level, blocksize)
    for k, v in over.items():
        if k == "apodization":
            lib().fo_config_set_apodization(C.byref(cfg), v.encode() if isinstance(v, str) else v)
        else:
            setattr(cfg, k, v)
    return cfg


def window(apod, length):
    out = np.empty(length, dtype=np.float32)
    if not lib().fo_window(C.byref(apod), out.ctypes.data, length):
        raise ValueError("unknown apodization type")
    return out


class Encoder:
    def __init__(self, cfg):
        self.h = lib().fo_encoder_new(C.byref(cfg))
        if not self.h:
            raise ValueError("configuration outside oracle scope")
        self.cfg = lib().fo_encoder_config(self.h).contents

    def close(self):
        if self.h:
            lib().fo_encoder_delete(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def encode_frame(self, pcm, frame_number=0, want_plan=False):
        pcm = np.ascontiguousarray(pcm, dtype=np.int32)
        n, ch = pcm.shape
        cap = 64 + n * ch * 5
        out = np.zeros(cap, dtype=np.uint8)
        plan = FramePlan() if want_plan else None
        ln = lib().fo_encode_frame(self.h, pcm.ctypes.data, n, frame_number, out.ctypes.data, cap,
                                   C.byref(plan) if want_plan else None)
        if ln == 0:
            raise RuntimeError("fo_encode_frame failed")
        return (out[:ln].tobytes(), plan) if want_plan else out[:ln].tobytes()

    def encode_stream(self, pcm):
        pcm = np.ascontiguousarray(pcm, dtype=np.int32)
        n, ch = pcm.shape
        bs = self.cfg.blocksize
        cap = 1024 + n * ch * 5 + 64 * (n // bs + 2)
        out = np.zeros(cap, dtype=np.uint8)
        maxf = n // bs + 2
        fs = np.zeros(maxf, dtype=np.uint32)
        out_len, nf = C.c_size_t(0), C.c_size_t(0)
        rc = lib().fo_encode_stream(self.h, pcm.ctypes.data, n, out.ctypes.data, cap, C.byref(out_len), fs.ctypes.data, maxf, C.byref(nf))
        if rc != 0:
            raise RuntimeError("fo_encode_stream failed")
        frames, pos = [], 0
        for s in fs[:nf.value]:
            frames.append(out[pos:pos + int(s)].tobytes())
            pos += int(s)
        return frames


def decode_frames(stream_bytes, channels, bps, rate, max_samples):
    """Decode consecutive frames (no metadata header). Returns int32 [samples, channels]."""
    buf = np.frombuffer(stream_bytes, dtype=np.uint8)
    si = StreamInfo(channels, bps, rate)
    out = np.zeros((max_samples, channels), dtype=np.int32)
    pos, done = 0, 0
    bs, ch, b, num = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_uint64(0)
    L = lib()
    while pos < buf.size:
        used = L.fo_decode_frame(buf.ctypes.data + pos, buf.size - pos, C.byref(si), out.ctypes.data + done * channels * 4,
                                 max_samples - done, C.byref(bs), C.byref(ch), C.byref(b), C.byref(num))
        if used == 0:
            raise RuntimeError(f"fo_decode_frame failed at byte {pos}")
        pos += used
        done += bs.value
    return out[:done]
