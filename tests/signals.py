"""Deterministic synthetic PCM generators used by tests and bench.py.

Families follow SURVEY.md §8(d):
  (i)   music_like  -- AM sinusoids + low-passed Gaussian noise (carries a noise floor),
  (ii)  reference-fixture restatements (noisy-sine LCG family, full-scale deflection, wasted bits),
  (iii) stress inputs (white noise, silence, DC, two-tone, pure sines).
All return int32 arrays shaped [samples, channels] (interleaved order, C-contiguous).
"""
import numpy as np


def _fit(x, bps):
    lim = (1 << (bps - 1)) - 1
    return np.clip(np.rint(x), -lim - 1, lim).astype(np.int32)


def music_like(nsamples, channels=2, bps=16, rate=44100, seed=1):
    rng = np.random.default_rng(seed)
    hires = rate >= 88200
    ntones = 24 if hires else 12
    t = np.arange(nsamples, dtype=np.float64) / rate
    taps = np.exp(-np.arange(64) / 8.0)
    taps /= taps.sum()

    def voice(r):
        x = np.zeros(nsamples)
        for _ in range(ntones):
            f = r.uniform(30, 20000) if hires else r.uniform(60, 6000)
            a = r.uniform(0.02, 0.2)
            am = r.uniform(0.1, 3.0)
            ph = r.uniform(0, 2 * np.pi)
            x += a * (0.6 + 0.4 * np.sin(2 * np.pi * am * t + ph)) * np.sin(2 * np.pi * f * t + ph)
        n = r.standard_normal(nsamples + 63)
        x += (0.01 if hires else 0.05) * np.convolve(n, taps, mode="valid")
        return x

    common = voice(rng)
    chans = []
    for c in range(channels):
        ind = voice(np.random.default_rng(seed * 1000 + 100 + c))
        chans.append(common + 0.3 * ind)
    x = np.stack(chans, axis=1)
    x *= 0.95 * ((1 << (bps - 1)) - 1) / np.max(np.abs(x))
    return np.ascontiguousarray(_fit(x, bps))


def white_noise(nsamples, channels=2, bps=16, seed=5, scale=1.0):
    rng = np.random.default_rng(seed)
    lim = int(((1 << (bps - 1)) - 1) * scale)
    return np.ascontiguousarray(rng.integers(-lim - 1, lim + 1, size=(nsamples, channels), dtype=np.int64).astype(np.int32))


def silence(nsamples, channels=2):
    return np.zeros((nsamples, channels), dtype=np.int32)


def dc(nsamples, channels=2, value=1234):
    return np.full((nsamples, channels), value, dtype=np.int32)


def sine(nsamples, channels=1, bps=16, rate=44100, freq=441.0, amp=0.9, freq2=None):
    t = np.arange(nsamples, dtype=np.float64) / rate
    x = np.sin(2 * np.pi * freq * t)
    if freq2 is not None:
        x = 0.5 * (x + np.sin(2 * np.pi * freq2 * t))
    x = x * amp * ((1 << (bps - 1)) - 1)
    return np.ascontiguousarray(np.repeat(_fit(x, bps)[:, None], channels, axis=1))


def noisy_sine(nsamples, channels=2, bps=16, rate=44100):
    """Sine + LCG noise, after the reference fixture generator
    (/root/reference/src/test_streams/main.c:1090-1131: state = 11117*state + 211231)."""
    t = np.arange(nsamples, dtype=np.float64) / rate
    out = np.zeros((nsamples, channels), dtype=np.float64)
    state = 12345
    noise = np.empty(nsamples * channels)
    for i in range(nsamples * channels):
        state = (11117 * state + 211231) & 0xFFFFFFFF
        noise[i] = ((state >> 8) & 0xFFFF) / 65536.0 - 0.5
    noise = noise.reshape(nsamples, channels)
    for c in range(channels):
        out[:, c] = 0.6 * np.sin(2 * np.pi * (441.0 * (c + 1)) * t) + 0.1 * noise[:, c]
    return np.ascontiguousarray(_fit(out * ((1 << (bps - 1)) - 1), bps))


def wasted_bits(nsamples, channels=2, bps=16, wasted=3, seed=7):
    x = music_like(nsamples, channels, bps - wasted, seed=seed)
    return np.ascontiguousarray(x << wasted)


def full_scale_deflection(nsamples, channels=1, bps=16, period=8):
    """Square-ish full-scale pattern, after the fsd* fixtures (test_streams/main.c:306-433)."""
    hi, lo = (1 << (bps - 1)) - 1, -(1 << (bps - 1))
    pat = np.where((np.arange(nsamples) // period) % 2 == 0, hi, lo).astype(np.int32)
    return np.ascontiguousarray(np.repeat(pat[:, None], channels, axis=1))
