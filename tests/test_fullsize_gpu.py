"""GPU parity at the BASELINE.json batch sizes.

The other GPU tests use <= 130 instances; these run the banks at the sizes bench.py times — grids of thousands of CTAs,
the 4-slice host path of b200m_r128_run_host, the residue-class gate scheduling over 8192 instances, 32-bit index
arithmetic at 16384 rows — and compare EVERY instance with oracle/_ref (the reference's own classes):

 * C2 + headline: 8192 stereo EBUr128 cycle, 480 blocks of 1024 frames (I and LRA gates live), device path, host path
   (b200m_r128_run_host, 4 copy/compute slices) and two half banks (the multi-GPU shard layout): nine floats, dBTP hold and
   every histogram bin of every instance, bit for bit;
 * C3: 8192 stereo TPnRMS (TruePeakdsp::process + Kmeterdsp::process, read every block), bit for bit;
 * C4: 4096 stereo spectr30, 60 ports of every instance, bit for bit;
 * C5: 2048 stereo Stcorrdsp bit for bit; phasewheel spectra of all 2048 instances against numpy's float64 FFT.
Input: a ring of 8 distinct blocks per row (what bench.py uses) so that host memory stays at 512 MiB.
"""
import ctypes as C

import numpy as np
import pytest

import _oracle as O
import _signals as S

pytestmark = pytest.mark.gpu
NFRAM, RING = 1024, 8
NAMES = ("loudness_M", "maxloudn_M", "loudness_S", "maxloudn_S", "integrated", "integ_thr", "range_min", "range_max", "range_thr")


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _ring(rows, seed):
    return S.white(rows, RING * NFRAM, seed=seed)


def _threads():
    return O.cpu_info("best")[0]


def _blk_ptr(x, b):
    return C.c_void_p(x.ctypes.data + 4 * NFRAM * (b % RING))


def test_c2_ebur128_8192_stereo_device_host_and_sharded():
    import torch
    import meters_lv2_b200 as B
    n_inst, nb = 8192, 480   # 480 blocks = 204 fragments: >= 50 M-points (integrated) and >= 20 S-points (LRA)
    x = _ring(2 * n_inst, seed=101)
    x[2 * 4000:2 * 4000 + 2] = 0.0                               # one silent instance (gates never open)
    x[2 * 77] *= 30.0                                            # one hot channel (+5 dB bins clamp, _error counts)
    xd = torch.from_numpy(x).cuda()
    stride, base = xd.stride(0), xd.data_ptr()
    dev = B.EBUr128(n_inst, 48000.0, True); dev.control(B.EBUr128.START)
    host = B.EBUr128(n_inst, 48000.0, True); host.control(B.EBUr128.START)
    half = [B.EBUr128(n_inst // 2, 48000.0, True) for _ in range(2)]
    for h in half:
        h.control(B.EBUr128.START)
    hbuf = [B.host_alloc(2 * n_inst, NFRAM) for _ in range(2)]   # pinned, dense: the e2e path of bench.py

    L = O.load("best"); thr = _threads()
    oe = O.Ebu(n_inst, 2); ot = O.TruePeak(2 * n_inst); oe.integr("start")
    tpmax = np.full(n_inst, -np.inf, np.float32)
    m = np.empty(2 * n_inst, np.float32); p = np.empty(2 * n_inst, np.float32)
    for b in range(nb):
        r = b % RING
        dev.run_ptr(base + 4 * NFRAM * r, stride, NFRAM)
        hb = hbuf[b & 1]
        if b >= 2:
            torch.cuda.synchronize()                             # the previous use of this host buffer has been copied
        hb[:] = x[:, r * NFRAM:(r + 1) * NFRAM]
        host.run_ptr(hb.ctypes.data, NFRAM, NFRAM, host=True)
        for k, h in enumerate(half):
            h.run_ptr(base + 4 * (stride * n_inst * k + NFRAM * r), stride, NFRAM)
        # the reference, one ebur128_run audio cycle per instance (src/ebulv2.cc:341-367)
        L.orc_ebu_process(oe.h, _blk_ptr(x, b), RING * NFRAM, NFRAM, thr)
        L.orc_tp_process(ot.h, _blk_ptr(x, b), RING * NFRAM, NFRAM, 1, thr)
        L.orc_tp_read(ot.h, O.ptr(m), O.ptr(p))
        v = np.maximum(m[0::2], m[1::2])
        with np.errstate(divide="ignore"):
            t = np.where(v == 0, -np.inf, (20.0 * np.log10(v.astype(np.float32)).astype(np.float64)).astype(np.float32))
        tpmax = np.maximum(tpmax, t)
    ref = oe.read()
    assert (ref[:, 4] > -100).sum() > n_inst - 8 and (ref[:, 7] > -100).sum() > n_inst - 8, "integrated loudness / LRA must be live"

    def check(bank, lo, hi, what):
        res, tp = bank.results()
        for i, k in enumerate(NAMES):
            bad = np.nonzero(u32(res[k]) != u32(ref[lo:hi, i]))[0]
            assert bad.size == 0, (what, k, bad[:5] + lo)
        # np.log10 is not the libm log10f the reference calls: the hold is compared to 2 ulp here, bit-exactness of the dB
        # conversion is pinned by tests/test_log10f_sweep_gpu.py and tests/test_lv2_ebur128_gpu.py (reference plugin's own tp_max)
        fin = np.isfinite(tpmax[lo:hi])
        assert np.array_equal(np.isfinite(tp), fin)
        assert np.abs(tp[fin].astype(np.float64) - tpmax[lo:hi][fin].astype(np.float64)).max() <= 8e-6, what      # 2 ulp at 32 dB
        return res

    rd = check(dev, 0, n_inst, "device path")
    rh = check(host, 0, n_inst, "host path (4 slices)")
    check(half[0], 0, n_inst // 2, "shard 0"); check(half[1], n_inst // 2, n_inst, "shard 1")
    for k in ("hist_M_count", "hist_S_count"):
        assert np.array_equal(rd[k], rh[k])
    # every histogram bin of every instance
    hm = np.empty(751, np.int32); hs = np.empty(751, np.int32); c4 = np.empty(4, np.int32)
    for i in range(n_inst):
        L.orc_ebu_hist(oe.h, i, O.ptr(hm), O.ptr(hs), O.ptr(c4))
        gm, gs = dev.ebu.histogram(i)
        assert np.array_equal(gm, hm) and np.array_equal(gs, hs), ("device path histogram", i)
        assert rd["hist_M_count"][i] == c4[0] and rd["hist_S_count"][i] == c4[1], i
        if i % 16 == 5:
            gm, gs = host.ebu.histogram(i)
            assert np.array_equal(gm, hm) and np.array_equal(gs, hs), ("host path histogram", i)
            bk, j = half[i // (n_inst // 2)], i % (n_inst // 2)
            gm, gs = bk.ebu.histogram(j)
            assert np.array_equal(gm, hm) and np.array_equal(gs, hs), ("shard histogram", i)


def test_c3_tpnrms_8192_stereo():
    import torch
    import meters_lv2_b200 as B
    nch, nb = 16384, 12
    x = _ring(nch, seed=103)
    xd = torch.from_numpy(x).cuda()
    g = B.TruePeakKmeter(nch)
    ot = O.TruePeak(nch); ok = O.Kmeter(nch)
    L = O.load("best"); thr = _threads()
    for b in range(nb):
        g.process_ptr(xd.data_ptr() + 4 * NFRAM * (b % RING), xd.stride(0), NFRAM)
        L.orc_tp_process(ot.h, _blk_ptr(x, b), RING * NFRAM, NFRAM, 0, thr)
        L.orc_km_process(ok.h, _blk_ptr(x, b), RING * NFRAM, NFRAM, thr)
        r = g.read(); m, p = ot.read(); rms, pk = ok.read()
        for got, want, what in ((r["tp_m"], m, "tp_m"), (r["tp_p"], p, "tp_p"), (r["km_rms"], rms, "km_rms"), (r["km_peak"], pk, "km_peak")):
            bad = np.nonzero(u32(got) != u32(want))[0]
            assert bad.size == 0, (what, b, bad[:5])


def test_c4_spectr30_4096_stereo():
    import torch
    import meters_lv2_b200 as B
    n_inst, nb = 4096, 5
    x = _ring(2 * n_inst, seed=105)
    xd = torch.from_numpy(x).cuda()
    g = B.Spectr30(n_inst, 2); o = O.Spectr30(n_inst, 2)
    L = O.load("best"); thr = _threads()
    for b in range(nb):
        g.process_ptr(xd.data_ptr() + 4 * NFRAM * (b % RING), xd.stride(0), NFRAM)
        L.orc_spec_process(o.h, _blk_ptr(x, b), RING * NFRAM, NFRAM, C.c_float(1.0), C.c_float(-4.0), thr)
    got, want = g.read(), o.read()
    bad = np.argwhere(u32(got) != u32(want))
    assert bad.size == 0, bad[:5]


def test_c5_stcorr_and_phasewheel_2048_stereo():
    import torch
    import meters_lv2_b200 as B
    n_inst, nb = 2048, 6
    x = _ring(2 * n_inst, seed=107)
    xd = torch.from_numpy(x).cuda()
    co = B.Stcorrdsp(n_inst); oc = O.Stcorr(n_inst)
    pw = B.Phasewheel(n_inst, 1024); pw.debug_capture(True)
    L = O.load("best"); thr = _threads()
    N = 2048
    hist = np.zeros((2 * n_inst, N), np.float32)                # the last N samples per row = what the ring holds
    fired_at = -1
    for b in range(nb):
        ptr = xd.data_ptr() + 4 * NFRAM * (b % RING)
        co.process_ptr(ptr, xd.stride(0), NFRAM)
        fired = pw.process_ptr(ptr, xd.stride(0), NFRAM)
        L.orc_cor_process(oc.h, _blk_ptr(x, b), RING * NFRAM, NFRAM, thr)
        hist = np.concatenate([hist[:, NFRAM:], x[:, (b % RING) * NFRAM:(b % RING + 1) * NFRAM]], axis=1)
        if fired:
            fired_at = b
            snap = hist.copy()
    assert np.array_equal(u32(co.read()), u32(oc.read()))
    assert np.array_equal(u32(co.state()), u32(oc.peek()))
    assert fired_at >= 0
    # independent pin: numpy's float64 FFT of the Hann-windowed ring (gui/fft.c:69-79,122-161,318-333)
    i = np.arange(N)
    w = (0.5 - 0.5 * np.cos(2.0 * np.pi / (N - 1.0) * i)).astype(np.float32)
    w = (w.astype(np.float64) * (2.0 / w.astype(np.float64).sum())).astype(np.float32)      # float window, normalised to sum 2 in double
    Xr = np.fft.rfft((snap * w[None, :]).astype(np.float64), axis=1)        # [rows, N/2+1]
    P = (Xr.real ** 2 + Xr.imag ** 2)
    worst = 0.0
    for inst in range(0, n_inst, 1):
        pl, pr, fl, fr = pw.raw(inst)
        for ch, got in ((0, pl), (1, pr)):
            ref = P[2 * inst + ch, :1023]
            peak = ref.max()
            sel = ref[1:1022] >= peak * 1e-2                    # bins within 20 dB of the frame peak
            d = np.abs(10 * np.log10(got[1:1022][sel].astype(np.float64) / ref[1:1022][sel])).max()
            worst = max(worst, d)
            assert np.abs(np.sqrt(got[1:1022].astype(np.float64)) - np.sqrt(ref[1:1022])).max() <= 2e-6 * np.sqrt(peak), (inst, ch)
    assert worst <= 1e-4, worst
