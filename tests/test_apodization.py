"""Apodization coverage: every window function of FLAC__stream_encoder_set_apodization
(src/libFLAC/stream_encoder.c:1940-2065, src/libFLAC/window.c).

CPU (-m "not gpu"): the oracle's and the engine's host-side window tables against the compiled
reference's FLAC__window_* symbols, bit for bit; oracle frames against reference frames for
specification strings.  GPU: CUDA frames against the oracle and the compiled reference."""
import ctypes as C
import os

import numpy as np
import pytest

import oraclelib
import reflib
import signals
from conftest import require_ref

# (our type id, reference symbol, extra float args)
PLAIN = [
    (2, "FLAC__window_bartlett"), (3, "FLAC__window_bartlett_hann"), (4, "FLAC__window_blackman"),
    (5, "FLAC__window_blackman_harris_4term_92db_sidelobe"), (6, "FLAC__window_connes"), (7, "FLAC__window_flattop"),
    (9, "FLAC__window_hamming"), (10, "FLAC__window_hann"), (11, "FLAC__window_kaiser_bessel"), (12, "FLAC__window_nuttall"),
    (13, "FLAC__window_rectangle"), (14, "FLAC__window_triangle"), (17, "FLAC__window_welch"),
]
LENGTHS = [2, 3, 16, 17, 192, 577, 1152, 4096, 4097, 4608, 16384]

SPECS = [
    "hann", "bartlett;welch", "blackman_harris_4term_92db;flattop", "gauss(0.2)", "gauss(0.5);hamming",
    "bartlett_hann;connes;kaiser_bessel;nuttall;rectangle;triangle;blackman",
    "partial_tukey(2)", "partial_tukey(3/0.3/0.5)", "punchout_tukey(3)", "punchout_tukey(2/0.25/0.1)",
    "tukey(0.25);partial_tukey(2);punchout_tukey(3)",            # the pre-1.4 flac -8 family
    "partial_tukey(1/0.2/0.7)",                                  # parts<=1 degenerates to tukey(p)
    "tukey(0);tukey(1)", "subdivide_tukey(4/0.8)", "subdivide_tukey(2);hann;gauss(0.1)",
    "nonsense;welch", "nonsense", "gauss(0.9)",                  # unknown / rejected items fall back (:2058-2062)
    "partial_tukey(2);subdivide_tukey(3/0.5)",                   # '/' search runs past the ';' (:1994-1997)
    "partial_tukey(40)", "punchout_tukey(16);partial_tukey(16)",  # num + parts < 32 rule (:2004, :2025)
    "partial_tukey(2/0.999)", "partial_tukey(2/0.1/0)", "punchout_tukey(2/0.1/1.5)",
]


def _ref_window(L, sym, n, *params):
    f = getattr(L, sym)
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_int32] + [C.c_float] * len(params)
    out = np.full(n + 8, np.float32(-77.0))
    f(out.ctypes.data, n, *params)
    assert np.all(out[n:] == np.float32(-77.0))
    return out[:n]


# The shipped build flags (-fassociative-math ..., oracle/Makefile REF_FAST) let GCC re-associate the
# three-or-more-term float sums of these six generators: their tables differ from source order by
# <= 3 ulp there (measured below).  Source order (the strict build) is the semantics we pin; the other
# eleven generators are bit-identical under both builds.  Same G1/G2 split as DESIGN.md "FP semantics".
REASSOCIATED_BY_SHIPPED_FLAGS = ("bartlett_hann", "blackman", "blackman_harris_4term_92db", "flattop", "kaiser_bessel", "nuttall")


def _reassociated(name):
    return any(name == "FLAC__window_" + r or name == "FLAC__window_" + r + "_sidelobe" or r in name.split(";") for r in REASSOCIATED_BY_SHIPPED_FLAGS)


def _same_bits(a, b):
    return np.array_equal(a.view(np.uint32), b.view(np.uint32))


def _window_cases():
    cases = [(t, sym, ()) for t, sym in PLAIN]
    cases += [(8, "FLAC__window_gauss", (s,)) for s in (0.01, 0.2, 0.25, 0.5, 0.7, -1.0)]
    cases += [(0, "FLAC__window_tukey", (p,)) for p in (0.0, 0.01, 0.25, 0.5, 0.99, 1.0)]
    for p in (0.2, 0.05, 0.95, 0.0, 1.0):
        for start, end in ((0.0, 0.55), (0.45, 1.0), (0.3, 0.7), (0.0, 1.0)):
            cases.append((15, "FLAC__window_partial_tukey", (p, start, end)))
            cases.append((16, "FLAC__window_punchout_tukey", (p, start, end)))
    return cases


def _apod(cls, t, params):
    a = cls()
    a.type = t
    if len(params) >= 1:
        a.p = params[0]
    if len(params) == 3:
        a.start, a.end = params[1], params[2]
    return a


@pytest.mark.parametrize("variant", ["strict", "default"])
def test_oracle_window_tables_match_reference(variant):
    require_ref(variant)
    L = reflib.lib(variant)
    for t, sym, params in _window_cases():
        for n in LENGTHS:
            want = _ref_window(L, sym, n, *params)
            got = oraclelib.window(_apod(oraclelib.Apod, t, params), n)
            if variant == "default" and _reassociated(sym):
                assert np.abs(got - want).max() <= 3 * 2.0 ** -24, f"{sym} L={n}: more than 3 ulp(1.0) from the shipped build"
                continue
            assert _same_bits(got, want), f"{sym}{params} L={n} [{variant}]: {np.flatnonzero(got != want)[:5]}"


@pytest.mark.parametrize("variant", ["strict", "default"])
def test_engine_window_tables_match_reference(variant):
    """The product's host-side generators (flac_b200/csrc/windows.h via fb200_window; no GPU needed)."""
    require_ref(variant)
    import flac_b200
    L = reflib.lib(variant)
    for t, sym, params in _window_cases():
        for n in LENGTHS:
            want = _ref_window(L, sym, n, *params)
            got = flac_b200.window(_apod(flac_b200.Apodization, t, params), n)
            if variant == "default" and _reassociated(sym):
                assert np.abs(got - want).max() <= 3 * 2.0 ** -24, f"{sym} L={n}"
                continue
            assert _same_bits(got, want), f"{sym}{params} L={n} [{variant}]"


def test_engine_and_oracle_parse_specifications_identically():
    """Both parsers against each other, field by field (the reference's parsed list is private;
    its effect is checked through the frames below)."""
    import flac_b200
    for spec in SPECS:
        a = flac_b200.preset(2, 16, 44100, 5, apodization=spec)
        b = oraclelib.preset(2, 16, 44100, 5, apodization=spec)
        assert a.num_apodizations == b.num_apodizations, spec
        for i in range(a.num_apodizations):
            x, y = a.apodizations[i], b.apodizations[i]
            assert (x.type, x.parts) == (y.type, y.parts), (spec, i)
            for f in ("p", "start", "end"):
                assert np.float32(getattr(x, f)).view(np.uint32) == np.float32(getattr(y, f)).view(np.uint32), (spec, i, f)
    assert flac_b200.preset(2, 16, 44100, 5, apodization="partial_tukey(40)").num_apodizations == 1
    assert flac_b200.preset(2, 16, 44100, 5, apodization="punchout_tukey(16);partial_tukey(16)").num_apodizations == 16


@pytest.mark.parametrize("spec", SPECS)
def test_oracle_frames_match_reference_for_specification(spec):
    require_ref("strict")
    x = signals.music_like(4096 * 2 + 321, 2, 16, 44100, seed=23)
    for level, variant in ((5, "strict"), (8, "strict"), (5, "default")):
        if variant == "default" and _reassociated(spec):
            continue  # G2 is not gated where the shipped flags re-associate the window sum (see above)
        enc = oraclelib.Encoder(oraclelib.preset(2, 16, 44100, level, apodization=spec))
        got = enc.encode_stream(x)
        _, _, ref = reflib.encode(x, 16, rate=44100, level=level, variant=variant, opts=reflib.RefEncOpts(apodization=spec))
        assert len(got) == len(ref)
        bad = [i for i, (a, b) in enumerate(zip(got, ref)) if a != b]
        assert not bad, f"{spec!r} level {level} [{variant}]: frames {bad} differ"


@pytest.mark.parametrize("bs,ch,bps", [(1152, 1, 16), (4608, 2, 24), (577, 2, 16)])
def test_oracle_frames_match_reference_other_shapes(bs, ch, bps):
    require_ref("strict")
    x = signals.music_like(bs * 2 + 50, ch, bps, 48000, seed=29)
    for spec in ("tukey(0.25);partial_tukey(2);punchout_tukey(3)", "gauss(0.15);flattop;welch"):
        enc = oraclelib.Encoder(oraclelib.preset(ch, bps, 48000, 8, bs, apodization=spec))
        got = enc.encode_stream(x)
        _, _, ref = reflib.encode(x, bps, rate=48000, level=8, blocksize=bs, variant="strict", opts=reflib.RefEncOpts(apodization=spec))
        assert got == ref, spec


# --------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("spec", SPECS)
def test_gpu_frames_for_specification(spec):
    import flac_b200
    x = signals.music_like(4096 * 3 + 321, 2, 16, 44100, seed=23)
    for level in (5, 8):
        enc = flac_b200.Encoder(flac_b200.preset(2, 16, 44100, level, apodization=spec))
        got = enc.encode_frames(x)
        enc.close()
        want = oraclelib.Encoder(oraclelib.preset(2, 16, 44100, level, apodization=spec)).encode_stream(x)
        assert got == want, f"oracle: {spec!r} level {level}"
        if reflib.available("strict"):
            for variant in ("strict", "default"):
                if variant == "default" and _reassociated(spec):
                    continue
                _, _, ref = reflib.encode(x, 16, rate=44100, level=level, variant=variant, opts=reflib.RefEncOpts(apodization=spec))
                assert got == ref, f"reference[{variant}]: {spec!r} level {level}"


@pytest.mark.gpu
@pytest.mark.parametrize("bs,ch,bps", [(1152, 1, 16), (4608, 2, 24), (577, 2, 16), (4096, 8, 24)])
def test_gpu_frames_other_shapes(bs, ch, bps):
    import flac_b200
    x = signals.music_like(bs * 3 + 50, ch, bps, 48000, seed=29)
    for spec in ("tukey(0.25);partial_tukey(2);punchout_tukey(3)", "gauss(0.15);flattop;welch", "hann;subdivide_tukey(3)"):
        enc = flac_b200.Encoder(flac_b200.preset(ch, bps, 48000, 8, bs, apodization=spec))
        got = enc.encode_frames(x)
        enc.close()
        want = oraclelib.Encoder(oraclelib.preset(ch, bps, 48000, 8, bs, apodization=spec)).encode_stream(x)
        assert got == want, spec


@pytest.mark.gpu
def test_gpu_stream_api_set_apodization_string():
    """FLAC__stream_encoder_set_apodization on the object API reaches the same engine configuration."""
    import test_gpu_stream_api as T
    x = signals.music_like(4096 * 2 + 11, 2, 16, 44100, seed=5)
    for spec in ("tukey(0.25);partial_tukey(2);punchout_tukey(3)", "hann", "bogus"):
        _, frames = T.encode_with_api(x, 16, 44100, 8, apodization=spec, verify=True)
        want = oraclelib.Encoder(oraclelib.preset(2, 16, 44100, 8, apodization=spec)).encode_stream(x)
        assert frames == want, spec
