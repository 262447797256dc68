"""Independent reference for the phasewheel / stereoscope FFT path: numpy's float64 FFT (pocketfft) of the windowed frame.

The reference plugin calls FFTW3 (gui/fft.c:234), which is neither vendored nor installed here; neither this repo's CUDA kernel
nor its CPU restatement may therefore serve as the other's only witness.  What IS pinned by the reference source is everything
around the transform: the Hann window and its normalisation (gui/fft.c:69-79,122-161), the float multiplication of ring and
window (:329-333), the half-complex layout (Re = out[i], Im = out[N - i], :170-177) and X_k = sum x_n exp(-2 pi i n k / N),
FFTW's documented R2HC definition, which is numpy.fft.rfft's.
"""
import numpy as np


def hann_window(N):
    i = np.arange(N)
    w = (0.5 - 0.5 * np.cos(2.0 * np.pi / (N - 1.0) * i)).astype(np.float32)           # window[i] = a - b cos (c i), stored as float
    return (w.astype(np.float64) * (2.0 / w.astype(np.float64).sum())).astype(np.float32)   # window[i] *= 2 / sum (double), stored as float


def spectra(frames):
    """frames: [rows, N] float32, the last N ring samples in time order -> (X complex128 [rows, N/2+1], power, phase)"""
    frames = np.ascontiguousarray(frames, np.float32)
    N = frames.shape[1]
    xin = frames * hann_window(N)[None, :]                  # float32 product, as fft_in[i] *= window[i]
    X = np.fft.rfft(xin.astype(np.float64), axis=1)
    return X, X.real ** 2 + X.imag ** 2, np.arctan2(X.imag, X.real)


def compare(power, phase, X, lo_db=20.0):
    """engine power / phase (float32 [bins]) of one channel vs the float64 spectrum X[0 .. bins]: returns
    (max |dX| / max |X| over bins 1 .. bins-2, max dB error and max phase error [rad] over the bins within lo_db of the peak)"""
    bins = power.shape[0]
    k = np.arange(1, bins - 1)
    ref = X[k]
    got = np.sqrt(power[k].astype(np.float64)) * np.exp(1j * phase[k].astype(np.float64))
    amax = np.abs(ref).max()
    rel = np.abs(got - ref).max() / amax if amax > 0 else 0.0
    strong = np.abs(ref) ** 2 >= (amax ** 2) * 10 ** (-lo_db / 10)
    if not strong.any() or amax == 0:
        return rel, 0.0, 0.0
    db = np.abs(10 * np.log10(power[k][strong].astype(np.float64) / np.abs(ref[strong]) ** 2)).max()
    dph = np.abs(np.angle(np.exp(1j * (phase[k][strong].astype(np.float64) - np.angle(ref[strong]))))).max()
    return rel, db, dph
