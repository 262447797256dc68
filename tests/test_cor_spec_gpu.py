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
