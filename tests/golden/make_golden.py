"""Generates tests/golden/frames_v1.json from the COMPILED REFERENCE (oracle/_ref/libFLAC_ref.so,
shipped flags) -- run in the build container where oracle/_ref exists:

    python tests/golden/make_golden.py

For every case the fixture stores the SHA-256 of each reference frame, the frame sizes and the
first 48 bytes of frame 0 (header + start of the first subframe, human-checkable). Inputs are
regenerated from tests/signals.py by name/seed, so the fixture stays small. Both reference
builds (shipped flags / source-order FP) must agree on a case for it to be recorded.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))

import reflib  # noqa: E402
import signals  # noqa: E402

CASES = [
    # name, generator, kwargs, bps, rate, level, blocksize
    ("music16_l0", "music_like", dict(nsamples=4096 * 3 + 100, channels=2, bps=16, rate=44100, seed=1), 16, 44100, 0, 0),
    ("music16_l2", "music_like", dict(nsamples=4096 * 3 + 100, channels=2, bps=16, rate=44100, seed=1), 16, 44100, 2, 0),
    ("music16_l5", "music_like", dict(nsamples=4096 * 3 + 100, channels=2, bps=16, rate=44100, seed=1), 16, 44100, 5, 0),
    ("music16_l8", "music_like", dict(nsamples=4096 * 3 + 100, channels=2, bps=16, rate=44100, seed=1), 16, 44100, 8, 0),
    ("mono16_l5", "music_like", dict(nsamples=4096 * 2 + 7, channels=1, bps=16, rate=44100, seed=2), 16, 44100, 5, 0),
    ("music24_l8", "music_like", dict(nsamples=4096 * 2 + 33, channels=2, bps=24, rate=96000, seed=11), 24, 96000, 8, 0),
    ("surround24_l8", "music_like", dict(nsamples=4096 + 500, channels=8, bps=24, rate=192000, seed=12), 24, 192000, 8, 0),
    ("noise16_l8", "white_noise", dict(nsamples=4096 * 2, channels=2, bps=16, seed=5), 16, 44100, 8, 0),
    ("silence_l5", "silence", dict(nsamples=4096 * 2 + 5, channels=2), 16, 44100, 5, 0),
    ("wasted3_l8", "wasted_bits", dict(nsamples=4096 * 2, channels=2, bps=16, wasted=3), 16, 44100, 8, 0),
    ("noisy_sine_l8", "noisy_sine", dict(nsamples=4096 * 2, channels=2, bps=16), 16, 44100, 8, 0),
    ("bs1000_l5", "music_like", dict(nsamples=3300, channels=2, bps=16, rate=44100, seed=3), 16, 44100, 5, 1000),
    ("bs4608_l8", "music_like", dict(nsamples=4608 * 2 + 10, channels=2, bps=16, rate=44100, seed=3), 16, 44100, 8, 4608),
    # apodization specification strings (8th field), FLAC__stream_encoder_set_apodization
    ("music16_l8_pre14_windows", "music_like", dict(nsamples=4096 * 2 + 100, channels=2, bps=16, rate=44100, seed=21), 16, 44100, 8, 0,
     "tukey(0.25);partial_tukey(2);punchout_tukey(3)"),
    ("music16_l5_gauss", "music_like", dict(nsamples=4096 * 2 + 100, channels=2, bps=16, rate=44100, seed=22), 16, 44100, 5, 0, "gauss(0.2)"),
    ("music24_l8_welch_hann", "music_like", dict(nsamples=4096 * 2 + 9, channels=2, bps=24, rate=96000, seed=23), 24, 96000, 8, 0, "welch;hann"),
    ("music16_l5_bs1152_triangle_connes", "music_like", dict(nsamples=1152 * 3 + 5, channels=2, bps=16, rate=44100, seed=24), 16, 44100, 5, 1152,
     "triangle;connes;rectangle"),
]


def main():
    out = {"reference": reflib.lib().ref_version().decode(), "build": "oracle/Makefile ref (gcc -O3, shipped FP flags)", "cases": {}}
    for name, gen, kw, bps, rate, level, bs, *rest in CASES:
        apod = rest[0] if rest else None
        x = getattr(signals, gen)(**kw)
        opts = reflib.RefEncOpts(streamable_subset=0, **({"apodization": apod} if apod else {}))
        _, _, fd = reflib.encode(x, bps, rate=rate, level=level, blocksize=bs, variant="default", opts=opts)
        _, _, fs = reflib.encode(x, bps, rate=rate, level=level, blocksize=bs, variant="strict", opts=opts)
        assert fd == fs, f"{name}: reference builds disagree"
        out["cases"][name] = {
            "generator": gen, "kwargs": kw, "bps": bps, "rate": rate, "level": level, "blocksize": bs,
            **({"apodization": apod} if apod else {}),
            "input_sha256": hashlib.sha256(x.tobytes()).hexdigest(),
            "frame_sizes": [len(f) for f in fd],
            "frame_sha256": [hashlib.sha256(f).hexdigest() for f in fd],
            "frame0_head_hex": fd[0][:48].hex(),
        }
    with open(os.path.join(HERE, "frames_v1.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
