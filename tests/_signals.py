"""Seeded synthetic 48 kHz streams shared by the parity tests, the golden-vector generator and bench.py.

white(): counter-based white noise per channel with the gain ladder of SURVEY.md §8(d)
(channels span -6 .. -36 dBFS so that histograms differ per channel)."""
import numpy as np

SEED = 0x42B200


def gain_ladder(nch):
    c = np.arange(nch)
    return (10.0 ** (-(6.0 + 30.0 * (c % 97) / 96.0) / 20.0)).astype(np.float32)


def white(nch, n, seed=SEED, block=0):
    """[nch, n] float32 uniform(-1,1) * gain_c ; `block` selects an independent stretch of the stream."""
    rng = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, block, 0]))
    x = rng.random((nch, n), dtype=np.float32) * 2.0 - 1.0
    return np.ascontiguousarray(x * gain_ladder(nch)[:, None], dtype=np.float32)


def sine(n, f, fs=48000.0, amp=1.0, phase=0.0, start=0):
    t = (np.arange(n) + start) / fs
    return (amp * np.sin(2 * np.pi * f * t + phase)).astype(np.float32)


def nasty(nch, n, seed=7):
    """white noise sprinkled with NaN, +-Inf, denormals, zeros and full-scale clicks."""
    rng = np.random.default_rng(seed)
    x = white(nch, n, seed=seed)
    idx = rng.integers(0, n, size=(nch, 6))
    vals = np.array([np.nan, np.inf, -np.inf, 1e-42, 0.0, 1.0], np.float32)
    for c in range(nch):
        x[c, idx[c]] = vals
    return x


def lcg_stereo(nframes, s0=12345):
    """the survey's LCG noise: L[i], R[i] drawn alternately, +-0.25."""
    a, c, s = 1664525, 1013904223, s0
    v = np.empty(2 * nframes, np.uint32)
    for i in range(2 * nframes):
        s = (s * a + c) & 0xFFFFFFFF
        v[i] = s
    f = (0.25 * ((v >> 8).astype(np.float64) / 2 ** 23 - 1)).astype(np.float32)
    return np.ascontiguousarray(np.stack([f[0::2], f[1::2]]))
