"""Exhaustive device sweep of the engine's log10f against the host libm.

Every loudness value, dB port and histogram bin index of the engine goes through csrc/common.cuh::log10f_glibc, a
restatement of glibc's log10f (reference call sites: ebumeter/ebu_r128_proc.cc:116-122,140-141,259 feeding the integer bins
of :66-79).  This test evaluates it ON THE DEVICE for all 2^31 non-negative float bit patterns (+0 .. +Inf .. NaNs) plus a
band of negative inputs and requires bit equality with log10f of the libm this process is linked against; the outcome is
appended to gpurun_out/log10f_sweep.txt (a copy of one run is committed as profiles/r2_log10f_sweep.txt).
"""
import os
import time

import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHUNK = 1 << 24


def _sweep(first, total):
    import torch
    import meters_lv2_b200 as B
    L = O.load("best")
    threads = O.cpu_info("best")[0]
    dev = [torch.empty(CHUNK, dtype=torch.float32, device="cuda") for _ in range(2)]
    host = [torch.empty(CHUNK, dtype=torch.float32).pin_memory() for _ in range(2)]
    bad = np.zeros(3, np.uint32)
    mism, done, k = 0, 0, 0
    pending = None
    while done < total or pending is not None:
        cur = None
        if done < total:
            n = min(CHUNK, total - done)
            b = k & 1
            rc = B.lib().b200m_selftest_log10f(0, (first + done) & 0xFFFFFFFF, n, dev[b].data_ptr(), None)
            assert rc == 0, B.lib().b200m_last_error()
            host[b][:n].copy_(dev[b][:n], non_blocking=True)
            ev = torch.cuda.Event(); ev.record()
            cur = (first + done, n, b, ev)
            done += n; k += 1
        if pending is not None:                               # check chunk k-1 on the host while chunk k runs / copies
            f0, n0, b0, ev0 = pending
            ev0.synchronize()
            m = L.orc_log10f_check(f0 & 0xFFFFFFFF, n0, host[b0].data_ptr(), threads, bad.ctypes.data)
            if m and not mism:
                first_bad = bad.copy()
            mism += m
        pending = cur
    return mism, (first_bad if mism else None)


def test_log10f_device_equals_host_libm_for_every_nonnegative_float():
    t0 = time.time()
    mism, bad = _sweep(0x00000000, 1 << 31)                   # +0, denormals, normals, +Inf, all positive NaNs
    mneg, badn = _sweep(0x80000000, 1 << 24)                  # -0 and the smallest negatives (domain error path: NaN / -Inf)
    mneg2, badn2 = _sweep(0xBF000000, 1 << 24)                # around -0.5 .. -2
    dt = time.time() - t0
    import platform
    line = "log10f sweep: 2^31 non-negative + 2^25 negative bit patterns, device vs host libm (%s): %d + %d + %d mismatches, %.1f s\n" % (
        platform.libc_ver()[1] or "libc ?", mism, mneg, mneg2, dt)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "log10f_sweep.txt"), "a") as f:
        f.write(line)
    assert mism == 0, "first mismatch: x=%08x libm=%08x device=%08x" % tuple(bad)
    assert mneg == 0 and mneg2 == 0, (badn, badn2)
