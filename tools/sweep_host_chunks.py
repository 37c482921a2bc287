"""End-to-end (packed 16-bit host buffers through fb200_encode_host_packed) step time against the number of copy/compute
chunks per call (FB200_HOST_CHUNKS). Prints one line per (workload, chunks).
    python tools/sweep_host_chunks.py [chunks ...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import flac_b200  # noqa: E402

CHUNKS = [int(a) for a in sys.argv[1:]] or [3, 5, 8, 12, 16, 24]
for name, (ch, bps, rate, level) in {"cfg2": (2, 16, 44100, 5), "cfg2_l8": (2, 16, 44100, 8)}.items():
    blocks, bs = 10000, 4096
    x = bench.make_pcm(ch, bps, rate, blocks, bs, seed=1)
    pk = flac_b200.pack_pcm(x, 2)
    h_pcm = torch.empty(pk.shape, dtype=torch.uint8, pin_memory=True)
    h_pcm.numpy()[:] = pk
    ns = x.shape[0]
    for nchunks in CHUNKS:
        os.environ["FB200_HOST_CHUNKS"] = str(nchunks)
        enc = flac_b200.Encoder(flac_b200.preset(ch, bps, rate, level, bs), max_blocks_per_launch=blocks)
        out_cap = blocks * enc.max_frame_bytes + 64
        h_out = torch.empty(out_cap, dtype=torch.uint8, pin_memory=True)
        h_offs = torch.empty(blocks + 1, dtype=torch.int64, pin_memory=True)
        for _ in range(3):
            enc.encode_packed(h_pcm.numpy(), 2, ns, 0, out=h_out.numpy(), offsets=h_offs.numpy().view(np.uint64))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 10
        for _ in range(n):
            enc.encode_packed(h_pcm.numpy(), 2, ns, 0, out=h_out.numpy(), offsets=h_offs.numpy().view(np.uint64))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print(f"{name} host_chunks={nchunks}: {dt * 1e3:.3f} ms/step  {blocks * bs * ch / dt / 1e6:.0f} Msamples/s", flush=True)
        enc.close()
        del h_out, h_offs
