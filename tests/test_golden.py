"""Committed golden vectors (tests/golden/frames_v1.json, produced from the compiled reference by
tests/golden/make_golden.py): the oracle on CPU, and the CUDA path on the GPU box, must
reproduce every reference frame -- also where oracle/_ref is not available."""
import hashlib
import json
import os

import numpy as np
import pytest

import oraclelib
import signals

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frames_v1.json")))


def _input(case):
    x = getattr(signals, case["generator"])(**case["kwargs"])
    assert hashlib.sha256(x.tobytes()).hexdigest() == case["input_sha256"], "signal generator drifted"
    return x


def _over(case):
    return {"apodization": case["apodization"]} if "apodization" in case else {}


def _check(frames, case):
    assert [len(f) for f in frames] == case["frame_sizes"]
    assert [hashlib.sha256(f).hexdigest() for f in frames] == case["frame_sha256"]
    assert frames[0][:48].hex() == case["frame0_head_hex"]


@pytest.mark.parametrize("name", sorted(GOLD["cases"]))
def test_oracle_reproduces_golden_frames(name):
    case = GOLD["cases"][name]
    x = _input(case)
    enc = oraclelib.Encoder(oraclelib.preset(x.shape[1], case["bps"], case["rate"], case["level"], case["blocksize"], **_over(case)))
    _check(enc.encode_stream(x), case)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD["cases"]))
def test_cuda_reproduces_golden_frames(name):
    import flac_b200
    case = GOLD["cases"][name]
    x = _input(case)
    enc = flac_b200.Encoder(flac_b200.preset(x.shape[1], case["bps"], case["rate"], case["level"], case["blocksize"], **_over(case)))
    _check(enc.encode_frames(x), case)
    enc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD["cases"]))
def test_cuda_decodes_golden_inputs(name):
    """decode(our frames) == input, checked against the input hash from the fixture."""
    import flac_b200
    case = GOLD["cases"][name]
    x = _input(case)
    enc = flac_b200.Encoder(flac_b200.preset(x.shape[1], case["bps"], case["rate"], case["level"], case["blocksize"], **_over(case)))
    stream, offs = enc.encode(x)
    dec = flac_b200.Decoder(x.shape[1], case["bps"], case["rate"], enc.cfg.blocksize)
    y = dec.decode(stream, offs, total_samples=x.shape[0])
    assert hashlib.sha256(np.ascontiguousarray(y).tobytes()).hexdigest() == case["input_sha256"]
    enc.close(); dec.close()
