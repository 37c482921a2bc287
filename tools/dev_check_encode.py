"""Developer diagnostic (GPU box): encode with the CUDA path and the oracle and print the
first differences at plan level (which signal, which field). Not part of the product."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import flac_b200  # noqa: E402
import oraclelib  # noqa: E402
import signals  # noqa: E402

TYPES = {0: "const", 1: "verb", 2: "fixed", 3: "lpc", -1: "inactive"}


def plan_str_gpu(p):
    n = 1 << p.porder
    return f"{TYPES.get(p.type)} ord{p.order} w{p.wasted} bps{p.bps} prec{p.precision} sh{p.shift} m{p.method} po{p.porder} est{p.est_bits} qlp{list(p.qlp[:p.order]) if p.type == 3 else ''} k{list(p.params[:min(n, 8)])}"


def plan_str_or(p):
    n = 1 << p.partition_order
    return f"{TYPES.get(p.type)} ord{p.order} w{p.wasted_bits} bps{p.subframe_bps} prec{p.qlp_precision} sh{p.qlp_shift} m{p.rice_method} po{p.partition_order} est{p.estimate_bits} qlp{list(p.qlp_coeff[:p.order]) if p.type == 3 else ''} k{list(p.rice_params[:min(n, 8)])}"


def check(x, bps, rate, level, bs=0, max_show=4, **over):
    ch = x.shape[1]
    enc = flac_b200.Encoder(flac_b200.preset(ch, bps, rate, level, bs, **over))
    bsz = enc.cfg.blocksize
    frames = enc.encode_frames(x)
    nfull = x.shape[0] // bsz
    plans, ca = enc.debug_plans(min(nfull, 4096)) if nfull else (None, None)
    oover = {}
    for k, v in over.items():
        oover[{"do_mid_side_stereo": "do_mid_side", "loose_mid_side_stereo": "loose_mid_side"}.get(k, k)] = v
    o = oraclelib.Encoder(oraclelib.preset(ch, bps, rate, level, bs, **oover))
    shown = 0
    nbad = 0
    for i, fr in enumerate(frames):
        blk = x[i * bsz:(i + 1) * bsz]
        want, fp = o.encode_frame(blk, i, want_plan=True)
        if want != fr:
            nbad += 1
            if shown < max_show:
                shown += 1
                first = next((j for j in range(min(len(want), len(fr))) if want[j] != fr[j]), None)
                print(f"  frame {i}: len gpu {len(fr)} oracle {len(want)} first diff byte {first}")
                if plans is not None and i < nfull and i < 4096 and (x.shape[0] // bsz) <= 4096:
                    print(f"    ca gpu {ca[i]} oracle {fp.channel_assignment}")
                    for s in range(enc.nsig):
                        g = plans[i * enc.nsig + s]
                        if ch == 2 and fp.cand_valid[s if s < 2 else s]:
                            w = fp.cand[s]
                        elif s < ch:
                            w = fp.sub[s]
                        else:
                            continue
                        gs, ws = plan_str_gpu(g), plan_str_or(w)
                        print(f"    sig{s} {'OK ' if gs == ws else 'DIFF'} gpu: {gs}")
                        if gs != ws:
                            print(f"    sig{s}      ora: {ws}")
    print(f"level {level} ch{ch} bps{bps} bs{bsz} {over}: {nbad}/{len(frames)} frames differ")
    enc.close()
    return nbad


if __name__ == "__main__":
    print(flac_b200.lib().fb200_version().decode(), "devices:", flac_b200.lib().fb200_device_count())
    total = 0
    x = signals.music_like(4096 * 6 + 777, 2, 16, 44100, seed=1)
    for lvl in (0, 1, 2, 3, 5, 8):
        total += check(x, 16, 44100, lvl)
    total += check(signals.music_like(4096 * 3, 1, 16, 44100, seed=2), 16, 44100, 5)
    total += check(signals.music_like(4096 * 3, 2, 24, 96000, seed=11), 24, 96000, 8)
    total += check(signals.music_like(4096 * 2, 8, 24, 192000, seed=12), 24, 192000, 8)
    total += check(signals.silence(4096 * 2 + 5, 2), 16, 44100, 5)
    total += check(signals.white_noise(4096 * 2, 2, 16), 16, 44100, 8)
    total += check(signals.wasted_bits(4096 * 2, 2, 16, 3), 16, 44100, 8)
    print("TOTAL differing frames:", total)
