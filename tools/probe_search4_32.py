"""Deterministic probe for k_search4's 32-tap instantiation (orders 13..32) against the oracle.

    python tools/probe_search4_32.py

Encodes a matrix of (bps, max_lpc_order, blocksize, exhaustive, seed) on the GPU, compares every frame with
oracle/flac_oracle.c and, for frames that differ, prints the per-signal plan differences (type/order/partition
order/estimate) from fb200_debug_copy_plans -- the data needed to root-cause a mis-evaluated candidate."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import flac_b200  # noqa: E402
import oraclelib  # noqa: E402
import signals  # noqa: E402


def main():
    bad_total = 0
    n_total = 0
    cases = []
    for bps in (16, 12, 24, 20):
        for mlo in (13, 16, 17, 20, 24, 32):
            for bs in (4096, 4608, 1024, 2304):
                for ex in (0, 1):
                    if ex and (mlo > 17 or bs != 4096):
                        continue
                    cases.append((bps, mlo, bs, ex))
    for ci, (bps, mlo, bs, ex) in enumerate(cases):
        seed = 2 + ci % 5
        x = signals.music_like(bs * 3 + 99, 2, bps, 44100, seed=seed)
        over = dict(max_lpc_order=mlo)
        if ex:
            over["do_exhaustive_model_search"] = 1
        enc = flac_b200.Encoder(flac_b200.preset(2, bps, 44100, 8, bs, **over))
        try:
            got = enc.encode_frames(x)
        finally:
            enc.close()
        want = oraclelib.Encoder(oraclelib.preset(2, bps, 44100, 8, bs, **over)).encode_stream(x)
        bad = [i for i, (a, b) in enumerate(zip(got, want)) if a != b]
        n_total += len(want)
        bad_total += len(bad)
        print(f"bps {bps:2d} mlo {mlo:2d} bs {bs:5d} ex {ex} seed {seed}: {len(bad)}/{len(want)} frames differ {bad[:4]}", flush=True)
    print(f"TOTAL {bad_total}/{n_total} frames differ (FB200_SEARCH_KERNEL={os.environ.get('FB200_SEARCH_KERNEL', 'default')})")
    return 1 if bad_total else 0


if __name__ == "__main__":
    sys.exit(main())
