"""GPU parity: needle-meter ballistics bank (VU, IEC I/II PPM, BBC M/S PPM) vs the reference build, bit-exact."""
import numpy as np
import pytest

import _oracle as O
import _signals as S

pytestmark = pytest.mark.gpu


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("fs,n,blocks", [(48000.0, 37, [1024] * 12 + [1, 2, 3, 5, 1027, 8192, 64]), (44100.0, 1, [512] * 9), (96000.0, 130, [1024] * 3)])
def test_needle_meters_bit_exact(kind, fs, n, blocks):
    import torch
    import meters_lv2_b200 as B
    rows = 2 * n if kind == 3 else n
    x = S.white(rows, sum(blocks), seed=17 + kind)
    x *= 4.0                                               # drive the meters towards their clamp range
    g = B.NeedleMeters(n, kind, fs); o = O.Needle(n, kind, fs)
    assert np.array_equal(u32(B.design_ppm(kind, fs)), u32(o.coeffs()))
    if kind == 3:
        g.set_gain(-6, 14); o.set_gain(-6, 14)             # bbcm_run: S meter +14 dB when port 7 > 0.5 (src/meters.cc:561-562)
    xd = torch.from_numpy(x).cuda()
    pos = 0
    for bi, nb in enumerate(blocks):
        g.process(xd[:, pos:pos + nb]); o.process(np.ascontiguousarray(x[:, pos:pos + nb]), nthreads=8)
        pos += nb
        if bi % 3 != 2:
            assert np.array_equal(u32(g.read()), u32(o.read())), (kind, bi)
    assert np.array_equal(u32(g.state()), u32(o.peek()))


@pytest.mark.parametrize("kind", [0, 1, 3])
def test_needle_meters_nasty_input_host_path(kind):
    import meters_lv2_b200 as B
    n = 9
    rows = 2 * n if kind == 3 else n
    x = S.nasty(rows, 1000 * 5 + 1, seed=23)[:, 1:]
    g = B.NeedleMeters(n, kind); o = O.Needle(n, kind)
    for b in range(5):
        blk = x[:, b * 1000:(b + 1) * 1000]
        g.process(blk); o.process(np.ascontiguousarray(blk))
        assert np.array_equal(u32(g.read()), u32(o.read()))
    assert np.array_equal(u32(g.state()), u32(o.peek()))
