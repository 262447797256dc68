"""GPU parity: stereo correlation and 30-band spectrum banks vs the CPU oracle."""
import numpy as np
import pytest

import _oracle as O
import _signals as S

pytestmark = pytest.mark.gpu


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ------------------------------------------------------------------ Stcorr
def test_cor_coeffs_bitwise():
    import meters_lv2_b200 as B
    for fs in (48000, 44100, 96000):
        assert np.array_equal(u32(B.Stcorrdsp(1, fs).coeffs()), u32(O.Stcorr(1, fs).coeffs()))


@pytest.mark.parametrize("n_inst,blocks", [(67, [1024] * 40), (5, [64] * 10 + [480, 8192, 1, 3, 1023, 33] * 2)])
def test_cor_bit_exact(n_inst, blocks):
    import torch
    import meters_lv2_b200 as B
    x = S.white(2 * n_inst, sum(blocks), seed=21)
    x[2] = x[3]                                   # fully correlated pair
    x[5] = -x[4]                                  # anti-correlated pair
    g = B.Stcorrdsp(n_inst); o = O.Stcorr(n_inst)
    xd = torch.from_numpy(x).cuda()
    pos = 0
    for n in blocks:
        g.process(xd[:, pos:pos + n]); o.process(np.ascontiguousarray(x[:, pos:pos + n]), nthreads=4)
        pos += n
        assert np.array_equal(u32(g.read()), u32(o.read()))
    assert np.array_equal(u32(g.state()), u32(o.peek()))
    r = g.read()
    assert r[1] > 0.99 and r[2] < -0.99


def test_cor_host_path_and_nasty_input():
    import meters_lv2_b200 as B
    x = S.nasty(2 * 9, 1000 * 6 + 1)[:, 1:]
    g = B.Stcorrdsp(9); o = O.Stcorr(9)
    for b in range(6):
        blk = x[:, b * 1000:(b + 1) * 1000]
        g.process(blk); o.process(np.ascontiguousarray(blk))
    assert np.array_equal(u32(g.state()), u32(o.peek()))
    assert np.array_equal(u32(g.read()), u32(o.read()))


# ------------------------------------------------------------------ spectr30
def test_spec_coeffs_bitwise():
    import meters_lv2_b200 as B
    for fs in (48000.0, 96000.0, 44100.0):
        W = B.Spectr30(1, 2, fs).coeffs(); R = O.Spectr30(1, 2, fs).coeffs()
        assert np.array_equal(W.view(np.uint64), R.view(np.uint64)), fs


@pytest.mark.parametrize("n_inst,nchan,blocks", [(9, 2, [1024] * 24), (3, 1, [1024] * 8), (2, 2, [64] * 8 + [480, 8192, 1, 3, 1023, 777])])
def test_spec_bit_exact(n_inst, nchan, blocks):
    import torch
    import meters_lv2_b200 as B
    x = S.white(nchan * n_inst, sum(blocks), seed=31)
    g = B.Spectr30(n_inst, nchan); o = O.Spectr30(n_inst, nchan)
    xd = torch.from_numpy(x).cuda()
    pos = 0
    for n in blocks:
        g.process(xd[:, pos:pos + n]); o.process(np.ascontiguousarray(x[:, pos:pos + n]), nthreads=8)
        pos += n
    for inst in (0, n_inst - 1):
        z, v, m = g.state(inst); oz, ov, om = o.state(inst)
        assert np.array_equal(z.view(np.uint64), oz.view(np.uint64)), "biquad state (fp64) must be bit-exact"
        assert np.array_equal(u32(v), u32(ov)) and np.array_equal(u32(m), u32(om))
    gp, op = g.read(), o.read()
    assert np.array_equal(u32(gp), u32(op)), float(np.abs(gp - op).max())      # dB ports, glibc-exact log10f


@pytest.mark.parametrize("n_inst,nchan,blocks", [(9, 2, [1024] * 24), (2, 2, [64] * 8 + [480, 8192, 1, 3, 1023, 777])])
def test_spec_fma_mode_within_tolerance(n_inst, nchan, blocks):
    """B200M_PREC_FMA: fused multiply-adds in the fp64 biquad cascade.  Band levels and maxima (dB ports) within the contract's
    +-1e-4 dB of the reference; the filters run in double precision, so the measured deviation is ~1e-6 dB (float32 port rounding)"""
    import torch
    import meters_lv2_b200 as B
    x = S.white(nchan * n_inst, sum(blocks), seed=35)
    x[1] *= 1e-3
    g = B.Spectr30(n_inst, nchan); g.set_precision(B.PREC_FMA); o = O.Spectr30(n_inst, nchan)
    xd = torch.from_numpy(x).cuda()
    pos, worst = 0, 0.0
    for n in blocks:
        g.process(xd[:, pos:pos + n]); o.process(np.ascontiguousarray(x[:, pos:pos + n]), nthreads=8)
        pos += n
        gp, op = g.read().astype(np.float64), o.read().astype(np.float64)
        live = op > -100
        assert np.array_equal(live, gp > -100)
        worst = max(worst, float(np.abs(gp[live] - op[live]).max()))
    assert worst <= 1e-4, worst
    z, v, m = g.state(0); oz, ov, om = o.state(0)
    assert np.allclose(z, oz, rtol=1e-9, atol=1e-18 + 1e-9 * np.abs(oz).max())


def test_spec_speed_and_peak_reset_controls():
    import torch
    import meters_lv2_b200 as B
    x = S.white(2 * 4, 1024 * 12, seed=33)
    g = B.Spectr30(4, 2); o = O.Spectr30(4, 2)
    xd = torch.from_numpy(x).cuda()
    ctl = [(1.0, -4.0)] * 3 + [(5.0, -4.0)] * 3 + [(5.0, 1.0)] * 2 + [(0.001, 1.0)] * 2 + [(20.0, 2.0)] * 2
    for b, (spd, rst) in enumerate(ctl):
        g.process(xd[:, b * 1024:(b + 1) * 1024], spd, rst)
        o.process(np.ascontiguousarray(x[:, b * 1024:(b + 1) * 1024]), spd, rst)
        gp, op = g.read(), o.read()
        assert np.array_equal(u32(gp[:, :30]), u32(op[:, :30])), b
        pending = op[:, 30:] <= -500                       # reset handshake cycles: reference emits -500 - rand()
        assert np.array_equal(pending, gp[:, 30:] <= -500)
        assert np.array_equal(u32(gp[:, 30:][~pending]), u32(op[:, 30:][~pending])), b
    z, v, m = g.state(1); oz, ov, om = o.state(1)
    assert np.array_equal(u32(m), u32(om)) and np.array_equal(u32(v), u32(ov))


def test_spec_nasty_input_host_path():
    import meters_lv2_b200 as B
    x = S.nasty(2 * 3, 1024 * 4)
    g = B.Spectr30(3, 2); o = O.Spectr30(3, 2)
    for b in range(4):
        blk = np.ascontiguousarray(x[:, b * 1024:(b + 1) * 1024])
        g.process(blk); o.process(blk)
    z, v, m = g.state(2); oz, ov, om = o.state(2)
    assert np.array_equal(z.view(np.uint64), oz.view(np.uint64))
    assert np.array_equal(u32(g.read()), u32(o.read()))


# ------------------------------------------------------------------ Stcorr: time-parallel mode and the fused phasewheel feed
COR_TOL = 1e-5          # correlation is a ratio in [-1, 1]: the contract's +-1e-4 dB is 1.15e-5 relative


@pytest.mark.parametrize("n_inst,blocks", [(70, [1024] * 30), (5, [64] * 10 + [480, 8192, 1, 3, 1023, 33, 2049] * 2), (3, [8192] * 4)])
def test_cor_scan_mode_within_tolerance(n_inst, blocks):
    """B200M_PREC_FMA: a warp owns one pair and scans over time (csrc/cor.cu cor_scan_kernel) instead of one lane per pair"""
    import torch
    import meters_lv2_b200 as B
    x = S.white(2 * n_inst, sum(blocks), seed=22)
    x[2] = x[3]; x[5] = -x[4]
    g = B.Stcorrdsp(n_inst); g.set_precision(B.PREC_FMA); o = O.Stcorr(n_inst)
    xd = torch.from_numpy(x).cuda()
    pos, worst = 0, 0.0
    for n in blocks:
        g.process(xd[:, pos:pos + n]); o.process(np.ascontiguousarray(x[:, pos:pos + n]), nthreads=4)
        pos += n
        worst = max(worst, float(np.abs(g.read().astype(np.float64) - o.read()).max()))
    assert worst <= COR_TOL, worst
    gs, os_ = g.state().astype(np.float64), o.peek().astype(np.float64)
    assert (np.abs(gs - os_) <= 2e-5 * np.abs(os_) + 1e-9).all()      # the serial fp32 recurrence itself random-walks ~4e-6 over its 14400-sample memory
    r = g.read()
    assert r[1] > 0.99 and r[2] < -0.99


def test_cor_scan_mode_host_path_unaligned_and_nasty():
    import meters_lv2_b200 as B
    x = S.nasty(2 * 9, 1000 * 6 + 1)[:, 1:]                     # rows start 4 bytes off a 16-byte boundary
    g = B.Stcorrdsp(9); g.set_precision(B.PREC_FMA); o = O.Stcorr(9)
    for b in range(6):
        blk = x[:, b * 1000:(b + 1) * 1000]
        g.process(blk); o.process(np.ascontiguousarray(blk))
    # NaN / Inf poison a whole segment composite instead of one sample: both end in the per-call scrub (stcorrdsp.cc:65-75)
    assert np.isfinite(g.read()).all() and np.isfinite(g.state()).all()


@pytest.mark.parametrize("prec", ["exact", "scan"])
@pytest.mark.parametrize("bins,blocks", [(1024, [1024] * 9), (256, [480] * 12 + [8192, 33]), (6144, [8192] * 3)])
def test_fused_phasewheel_feed(prec, bins, blocks):
    """b200m_pw_attach_cor: one kernel reads the block once, runs Stcorrdsp and appends to the FFT ring.  The ring, hence every
    spectrum, is identical to the unfused bank's; the correlation is bit-exact in exact mode and within tolerance in scan mode."""
    import torch
    import meters_lv2_b200 as B
    n_inst = 37
    x = S.white(2 * n_inst, sum(blocks), seed=24)
    xd = torch.from_numpy(x).cuda()
    fused = B.Phasewheel(n_inst, bins); co = B.Stcorrdsp(n_inst); fused.attach_cor(co)
    if prec == "scan":
        co.set_precision(B.PREC_FMA)
    plain = B.Phasewheel(n_inst, bins); oc = O.Stcorr(n_inst)
    pos = 0
    for n in blocks:
        f1 = fused.process(xd[:, pos:pos + n]); f2 = plain.process(xd[:, pos:pos + n])
        oc.process(np.ascontiguousarray(x[:, pos:pos + n]), nthreads=4)
        pos += n
        assert f1 == f2
        if f1:
            for a, b in zip(fused.read(), plain.read()):
                assert np.array_equal(u32(a), u32(b))
        if prec == "exact":
            assert np.array_equal(u32(co.read()), u32(oc.read()))
        else:
            assert np.abs(co.read().astype(np.float64) - oc.read()).max() <= COR_TOL
