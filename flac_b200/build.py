"""Builds flac_b200/libflac_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m flac_b200.build [--force]

Flags that matter for bit-exactness: -fmad=false (no FMA contraction of the LPC analysis),
host side -ffp-contract=off (window tables use the host libm exactly like the reference).
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", f) for f in ("encoder.cu", "decoder.cu", "stream_api.cu")]
HDR = sorted(glob.glob(os.path.join(HERE, "csrc", "*.h")) + glob.glob(os.path.join(HERE, "csrc", "*.cuh"))) + [
    os.path.join(HERE, "..", "include", "flac_b200.h"), os.path.join(HERE, "..", "include", "flac_b200_stream.h")]
OUT = os.path.join(HERE, "libflac_b200.so")

NVCC_FLAGS = [
    "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=default", "-shared", "-lcudart",
]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in SRC + HDR + [__file__])


def build(force=False, verbose=False):
    srcs = [s for s in SRC if os.path.exists(s)]
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libflac_b200.so")
    if verbose:
        print(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
