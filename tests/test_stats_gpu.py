"""GPU parity: bit-meter and signal-distribution histogram banks vs the reference plugins driven through their own
LV2 run() (oracle/_ref, src/meters.cc compiled unmodified).  Integer tables bit-exact; fp64 statistics bitwise."""
import numpy as np
import pytest

import _oracle as O
import _signals as S

pytestmark = pytest.mark.gpu
BLOCKS = [1024] * 24 + [1, 4, 31, 33, 1000, 8192, 1024, 1024]


def _input(n, total, seed):
    x = S.nasty(n, total, seed=seed)
    x[1] *= np.float32(1e-39)          # denormals
    x[2] *= np.float32(3.0)            # beyond +-1: out-of-range bins for the distribution histogram
    x[3] = 0
    return x


@pytest.mark.parametrize("average", [1, 0])
def test_bitmeter_bit_exact(average):
    import torch
    import meters_lv2_b200 as B
    n = 37
    x = _input(n, sum(BLOCKS), 31)
    g = B.Bitmeter(n); o = O.Bitmeter(n)
    g.control(B.CTL_AVERAGE if average else B.CTL_WINDOWED); o.mode(average)
    xd = torch.from_numpy(x).cuda()
    pos = 0
    for bi, nb in enumerate(BLOCKS):
        g.run(xd[:, pos:pos + nb]); o.process(np.ascontiguousarray(x[:, pos:pos + nb]), nthreads=8)
        pos += nb
        if bi % 4 == 3 or nb != 1024:
            for inst in (0, 1, 2, 3, n - 1):
                gh, gc, gm, gt = g.results(inst); oh, oc, om, ot = o.read(inst)
                assert np.array_equal(gh, oh), (bi, inst, int((gh != oh).sum()))
                assert np.array_equal(gc, oc) and np.array_equal(gm.view(np.uint32), om.view(np.uint32)) and gt == ot, (bi, inst, gc, oc, gm, om, gt, ot)


def test_sigdist_bit_exact():
    import torch
    import meters_lv2_b200 as B
    n = 35
    x = _input(n, sum(BLOCKS), 32)
    g = B.SigDistHist(n); o = O.SigDist(n)
    xd = torch.from_numpy(x).cuda()
    g.run(xd[:, :512]); o.process(np.ascontiguousarray(x[:, :512]))          # not integrating yet: no effect
    g.control(B.CTL_START); o.integrate(True)
    pos = 0
    for bi, nb in enumerate(BLOCKS):
        g.run(xd[:, pos:pos + nb]); o.process(np.ascontiguousarray(x[:, pos:pos + nb]), nthreads=8)
        pos += nb
    for inst in (0, 1, 2, 3, n - 1):
        gh, gp, ga, gt = g.results(inst); oh, op, oa, ot = o.read(inst)
        assert np.array_equal(gh, oh) and np.array_equal(gp, op) and gt == ot, inst
        assert np.array_equal(ga.view(np.uint64), oa.view(np.uint64)), (inst, ga, oa)


def test_stats_host_path_and_reset():
    import meters_lv2_b200 as B
    n = 5
    x = _input(n, 4096 + 3, 33)[:, 3:]
    g = B.Bitmeter(n); o = O.Bitmeter(n); g.control(B.CTL_AVERAGE); o.mode(1)
    s = B.SigDistHist(n); os_ = O.SigDist(n); s.control(B.CTL_START); os_.integrate(True)
    for b in range(4):
        blk = x[:, b * 1024:(b + 1) * 1024]
        g.run(blk); o.process(np.ascontiguousarray(blk)); s.run(blk); os_.process(np.ascontiguousarray(blk))
    gh, gc, gm, gt = g.results(2); oh, oc, om, ot = o.read(2)
    assert np.array_equal(gh, oh) and np.array_equal(gc, oc) and gt == ot
    sh, sp, sa, st = s.results(2); rh, rp, ra, rt = os_.read(2)
    assert np.array_equal(sh, rh) and np.array_equal(sa.view(np.uint64), ra.view(np.uint64))
    g.control(B.CTL_RESET); s.control(B.CTL_RESET)
    assert g.results(2)[0].sum() == 0 and g.results(2)[3] == 0 and s.results(2)[1][1] == -1
