"""G3 evidence: frame mismatch counts of the reference's own pure-sine fixture families between
  * the CPU restatement oracle/flac_oracle.c (source-order floating point == what the CUDA path computes; the GPU parity
    tests pin CUDA == oracle) and
  * the two builds of the UNMODIFIED reference: "strict" (no fast-math flags) and "shipped" (-fassociative-math ...).
Writes profiles/g3_tonal_mismatch.json. CPU only.

The fixtures restate /root/reference/src/test_streams/main.c:435-664 (generate_sine{16,24}_{1,2}) with the parameter rows
of :1405-1460 (theta accumulates by delta per sample; (int)(val + 0.5) truncation toward zero)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oraclelib  # noqa: E402
import reflib  # noqa: E402

MONO = [(48000.0, 441.0, 0.50, 441.0, 0.49), (96000.0, 441.0, 0.61, 661.5, 0.37), (44100.0, 441.0, 0.50, 882.0, 0.49),
        (44100.0, 441.0, 0.50, 4410.0, 0.49), (44100.0, 8820.0, 0.70, 4410.0, 0.29)]
STEREO = [(48000.0, 441.0, 0.50, 441.0, 0.49, 1.0), (48000.0, 441.0, 0.61, 661.5, 0.37, 1.0), (96000.0, 441.0, 0.50, 882.0, 0.49, 1.0),
          (44100.0, 441.0, 0.50, 4410.0, 0.49, 1.0), (44100.0, 8820.0, 0.70, 4410.0, 0.29, 1.0), (44100.0, 441.0, 0.50, 441.0, 0.49, 0.5),
          (44100.0, 441.0, 0.61, 661.5, 0.37, 2.0), (44100.0, 441.0, 0.50, 882.0, 0.49, 0.7), (44100.0, 441.0, 0.50, 4410.0, 0.49, 1.3),
          (44100.0, 8820.0, 0.70, 4410.0, 0.29, 0.1)]


def sine(bps, rate, n, f1, a1, f2, a2, fmult=None):
    full = (1 << (bps - 1)) - 1
    d1, d2 = 2.0 * np.pi / (rate / f1), 2.0 * np.pi / (rate / f2)
    # theta_i = i * delta accumulated in double like the reference's running sum
    t1 = np.concatenate([[0.0], np.cumsum(np.full(n - 1, d1))])
    t2 = np.concatenate([[0.0], np.cumsum(np.full(n - 1, d2))])
    left = np.trunc((a1 * np.sin(t1) + a2 * np.sin(t2)) * full + 0.5).astype(np.int32)
    if fmult is None:
        return np.ascontiguousarray(left[:, None])
    right = np.trunc(-(a1 * np.sin(t1 * fmult) + a2 * np.sin(t2 * fmult)) * full + 0.5).astype(np.int32)
    return np.ascontiguousarray(np.stack([left, right], axis=1))


def main():
    n = 200000
    rows = []
    for bps in (16, 24):
        fixtures = [(f"sine{bps}-{i:02d}", p, None) for i, p in enumerate(MONO)] + [(f"sine{bps}-{10 + i}", p[:5], p[5]) for i, p in enumerate(STEREO)]
        for name, (rate, f1, a1, f2, a2), fmult in fixtures:
            x = sine(bps, rate, n, f1, a1, f2, a2, fmult)
            for level in (5, 8):
                ours = oraclelib.Encoder(oraclelib.preset(x.shape[1], bps, int(rate), level)).encode_stream(x)
                _, _, strict = reflib.encode(x, bps, rate=int(rate), level=level, variant="strict")
                _, _, shipped = reflib.encode(x, bps, rate=int(rate), level=level, variant="default")
                row = {"fixture": name, "level": level, "frames": len(ours),
                       "differ_vs_strict_build": sum(a != b for a, b in zip(ours, strict)),
                       "differ_vs_shipped_build": sum(a != b for a, b in zip(ours, shipped)),
                       "strict_vs_shipped_builds_differ": sum(a != b for a, b in zip(strict, shipped)),
                       "bytes_ours": sum(map(len, ours)), "bytes_shipped": sum(map(len, shipped))}
                rows.append(row)
                print(row, flush=True)
    out = {"what": "frames that differ between the source-order FP path (oracle restatement == CUDA path) and the two builds of the unmodified reference, "
                   "on the reference's noise-free sine fixture families (src/test_streams/main.c:1405-1460), 200 000 samples each",
           "summary": {"encodes": len(rows), "encodes_identical_to_strict_build": sum(r["differ_vs_strict_build"] == 0 for r in rows),
                       "encodes_identical_to_shipped_build": sum(r["differ_vs_shipped_build"] == 0 for r in rows),
                       "frames": sum(r["frames"] for r in rows), "frames_differ_vs_strict_build": sum(r["differ_vs_strict_build"] for r in rows),
                       "frames_differ_vs_shipped_build": sum(r["differ_vs_shipped_build"] for r in rows)},
           "rows": rows}
    with open(os.path.join(ROOT, "profiles", "g3_tonal_mismatch.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out["summary"]))


if __name__ == "__main__":
    main()
