"""Small encode + decode workload for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Covers the fast-path kernels (4096), the general kernels (blocksize 1000) and the decoder."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import flac_b200  # noqa: E402
import signals  # noqa: E402

for ch, bps, level, bs, n in ((2, 16, 8, 0, 4096 * 6 + 100), (2, 24, 8, 0, 4096 * 3), (2, 16, 5, 1000, 3300), (1, 16, 2, 0, 1152 * 4 + 7)):
    x = signals.music_like(n, ch, bps, 44100, seed=3)
    enc = flac_b200.Encoder(flac_b200.preset(ch, bps, 44100, level, bs), max_blocks_per_launch=4)
    stream, offs = enc.encode(x)
    dec = flac_b200.Decoder(ch, bps, 44100, enc.cfg.blocksize)
    y = dec.decode(stream, offs, total_samples=n)
    assert np.array_equal(x, y)
    enc.close(); dec.close()
print("sanitize smoke ok")
