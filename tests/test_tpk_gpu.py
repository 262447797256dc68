"""GPU parity: b200m_tpk_* (true peak 4x + K-meter) vs the CPU oracle; bit-exact on every output."""
import numpy as np
import pytest

import _oracle as O
import _signals as S

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["shortcut", "full", "slabs", "stagger", "fused", "roles_full"])
def phase0_mode(request, monkeypatch):
    """every test runs six times: with the exact phase-0 shortcut of the FIR (default) and with phase 0 always evaluated; with
    process() as the opt-in FIR / ballistics slab pipeline cut into 128-sample slabs; with the opt-in phase stagger of co-resident
    process() CTAs (every second CTA of an SM shortens its first chunk); with process() forced onto the fused kernel (the default is
    the decoupled-role kernel, csrc/tpk.cu tpdec_kernel); and with the decoupled roles evaluating phase 0 always."""
    monkeypatch.delenv("B200M_TPK_ELIDE0", raising=False)
    monkeypatch.delenv("B200M_TPK_DEC", raising=False)
    if request.param == "full":
        monkeypatch.setenv("B200M_TPK_ELIDE0", "0"); monkeypatch.setenv("B200M_TPK_DEC", "0")
    elif request.param == "fused":
        monkeypatch.setenv("B200M_TPK_DEC", "0")
    elif request.param == "roles_full":
        monkeypatch.setenv("B200M_TPK_ELIDE0", "0")
    elif request.param == "slabs":
        monkeypatch.setenv("B200M_TPK_SLAB", "128"); monkeypatch.setenv("B200M_TPK_SPLIT", "2")
    elif request.param == "stagger":
        monkeypatch.setenv("B200M_TPK_STAGGER", "1"); monkeypatch.setenv("B200M_TPK_DEC", "0")
    return request.param


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_coeffs_bitwise():
    import meters_lv2_b200 as B
    for fs in (48000.0, 44100.0, 96000.0):
        g = B.TruePeakKmeter(1, fs)
        w, t, k = g.coeffs()
        ow, ot = O.TruePeak(1, fs).coeffs()
        om, oh = O.Kmeter(1, fs).coeffs()
        assert np.array_equal(u32(w), u32(ow)) and np.array_equal(u32(t), u32(ot))
        assert u32(k[:1])[0] == u32(np.float32(om))[()] and int(k[1]) == oh


@pytest.mark.parametrize("n,block", [(4096, 1024), (3000, 1000), (777, 777), (8192, 8192), (130, 65)])
def test_fir_stream_bit_exact(n, block):
    """the raw 4x oversampled stream equals zita-resampler's output bit for bit."""
    import torch
    import meters_lv2_b200 as B
    x = S.white(11, n, seed=5)
    x[3] *= 1e-12; x[4] = 0; x[5, ::7] = 1e-40          # tiny / zero / denormal rows exercise the 1e-20f bias
    g = B.TruePeakKmeter(11); g.debug_capture(True)
    xd = torch.from_numpy(x).cuda()
    for ch in (0, 3, 4, 5, 10):
        pass
    outs = {ch: [] for ch in (0, 3, 4, 5, 10)}
    for o in range(0, n, block):
        k = min(block, n - o)
        g.process(xd[:, o:o + k])
        for ch in outs:
            outs[ch].append(g.debug_upsampled(ch, 4 * k))
    for ch in outs:
        ref = O.tp_upsample(x[ch], block=block)
        got = np.concatenate(outs[ch])
        assert np.array_equal(u32(got), u32(ref)), (ch, int((u32(got) != u32(ref)).sum()))


def _drive(x, blocks, read_every=1, mode=0, flags=3, host=False):
    import torch
    import meters_lv2_b200 as B
    C = x.shape[0]
    g = B.TruePeakKmeter(C, flags=flags)
    ot = O.TruePeak(C); ok = O.Kmeter(C)
    xd = None if host else torch.from_numpy(x).cuda()
    pos = 0
    reads = []
    for bi, n in enumerate(blocks):
        blk = np.ascontiguousarray(x[:, pos:pos + n])
        if flags & 1:
            ot.process(blk, mode=mode, nthreads=8)
        if flags & 2:
            ok.process(blk, nthreads=8)
        g.process(blk if host else xd[:, pos:pos + n], tp_mode=mode)
        pos += n
        if read_every and (bi + 1) % read_every == 0:
            r = g.read()
            if flags & 1:
                m, p = ot.read()
                assert np.array_equal(u32(r["tp_m"]), u32(m)), ("tp_m", bi)
                assert np.array_equal(u32(r["tp_p"]), u32(p)), ("tp_p", bi)
            if flags & 2:
                rms, pk = ok.read()
                assert np.array_equal(u32(r["km_rms"]), u32(rms)), ("km_rms", bi)
                assert np.array_equal(u32(r["km_peak"]), u32(pk)), ("km_peak", bi)
    s = g.state()
    if flags & 1:
        m, p, z1, z2, res = ot.peek()
        for k, v in (("m", m), ("p", p), ("z1", z1), ("z2", z2)):
            assert np.array_equal(u32(s[k]), u32(v)), k
        assert np.array_equal(s["res"], res)
    if flags & 2:
        km = ok.peek()
        assert np.array_equal(u32(s["km"]), u32(km)), "kmeter state"
    return g


@pytest.mark.parametrize("C,blocks,read_every", [
    (70, [1024] * 50, 1),                 # TPnRMS cadence: read after every run()
    (19, [1024] * 30, 0),                 # never read: max-merge branch, m *= g quirk
    (9, [64] * 20 + [480] * 10 + [8192] * 2 + [1, 2, 3, 5, 1023, 4097], 3),   # ragged blocks, n % 4 != 0
])
def test_process_bit_exact(C, blocks, read_every):
    x = S.white(C, sum(blocks), seed=11)
    _drive(x, blocks, read_every)


@pytest.mark.parametrize("flags", [3, 1])
def test_wide_ctas_bit_exact(flags, monkeypatch):
    """process() with 64-channel CTAs (the form banks of >= 9472 channels use; forced here with B200M_TPK_WIDE=2): 70 and 150
    channels (partial last CTA), ragged blocks, reads every block or never"""
    monkeypatch.setenv("B200M_TPK_WIDE", "2"); monkeypatch.setenv("B200M_TPK_SPLIT", "0")
    blocks = [1024] * 20 + [64] * 6 + [480, 8192, 1, 3, 1023, 33, 4097]
    _drive(S.white(70, sum(blocks), seed=12), blocks, read_every=1, flags=flags)
    _drive(S.white(150, 1024 * 10, seed=13), [1024] * 10, read_every=0, flags=flags)
    _drive(S.nasty(66, 1024 * 4), [1024] * 4, read_every=2, flags=flags)


def test_process_max_mode():
    x = S.white(21, 1024 * 12, seed=3)
    _drive(x, [1024] * 12, read_every=1, mode=1, flags=1)
    _drive(x, [1024] * 12, read_every=0, mode=1, flags=1)


def test_single_meter_banks_and_host_path():
    x = S.white(10, 1000 * 8 + 1, seed=9)[:, 1:]           # unaligned rows
    _drive(x, [1000] * 8, flags=1, host=True)
    _drive(x, [1000] * 8, flags=2, host=True)


def test_nan_inf_denormal():
    x = S.nasty(16, 1024 * 6)
    _drive(x, [1024] * 6, read_every=2)


def test_true_peak_of_fs4_sine():
    # fs/4 sine at 45 deg: sample peaks 0.7071, true peak 1.0 (+3 dB) -- SURVEY App. C: p -> 1.0000032
    n = 1024 * 47
    x = np.ascontiguousarray(S.sine(n, 12000.0, phase=np.pi / 4)[None, :])
    g = _drive(x, [1024] * 47, read_every=0, flags=1)
    s = g.state()
    assert abs(s["p"][0] - 1.0) < 0.02 and s["p"][0] > 0.99


def test_reset():
    import torch
    import meters_lv2_b200 as B
    x = S.white(6, 2048, seed=2)
    g = B.TruePeakKmeter(6); ot = O.TruePeak(6); ok = O.Kmeter(6)
    xd = torch.from_numpy(x).cuda()
    g.process(xd[:, :1024]); ot.process(np.ascontiguousarray(x[:, :1024])); ok.process(np.ascontiguousarray(x[:, :1024]))
    g.reset(2); ot.reset(2); ok.reset(2)
    g.process(xd[:, 1024:]); ot.process(np.ascontiguousarray(x[:, 1024:])); ok.process(np.ascontiguousarray(x[:, 1024:]))
    s = g.state(); m, p, z1, z2, res = ot.peek()
    assert np.array_equal(u32(s["m"]), u32(m)) and np.array_equal(u32(s["p"]), u32(p))
    assert np.array_equal(u32(s["km"]), u32(ok.peek()))



def _guard_signals(n, seed=3):
    """rows built to sit on both sides of the phase-0 guard (csrc/tpk.cu: |x| > 1.25e-7*M + 1e-12)."""
    rng = np.random.default_rng(seed)
    rows = []
    base = rng.uniform(-1, 1, n).astype(np.float32)
    rows.append(base.copy())                                                  # plain noise: fast path
    r = base.copy(); r[rng.random(n) < 0.9] = 0; rows.append(r)               # sparse impulses: x == 0 next to large neighbours
    e = rng.integers(-40, 1, n); rows.append((base * np.exp2(e)).astype(np.float32))     # 40 octaves of dynamic range
    e = rng.integers(-30, -18, n); r = base.copy(); k = rng.random(n) < 0.3
    r[k] = (base[k] * np.exp2(e[k])).astype(np.float32); rows.append(r)       # main taps right around 2^-23 * M
    r = base.copy(); r[::5] *= np.float32(1.25e-7); r[1::5] *= np.float32(1.19e-7); r[2::5] *= np.float32(1.3e-7); rows.append(r)
    rows.append((base * np.float32(1e-12)).astype(np.float32))                # around the absolute floor of the guard
    rows.append((base * np.float32(4e-13)).astype(np.float32))
    r = (base * np.float32(1e-38)).astype(np.float32); rows.append(r)         # denormal products
    r = base.copy(); r[100] = np.inf; r[400] = np.nan; r[700] = -np.inf; rows.append(r)
    r = np.zeros(n, np.float32); r[n // 2] = 1.0; rows.append(r)              # unit impulse: reads the table itself
    r = np.full(n, 0.5, np.float32); rows.append(r)                           # DC
    rows.append(np.zeros(n, np.float32)); rows.append(np.full(n, -0.0, np.float32))          # digital silence, either sign
    r = np.zeros(n, np.float32); r[n // 3:] = base[n // 3:]; r[2 * n // 3:] = -0.0; rows.append(r)   # silence -> signal -> silence
    r = np.where(np.arange(n) % 2 == 0, 1.0, -1.0).astype(np.float32) * 0.9; rows.append(r)   # fs/2
    return np.stack(rows)


@pytest.mark.parametrize("block", [1024, 333])
def test_phase0_guard_fir_stream(block):
    """the phase-0 shortcut never changes a bit of the oversampled stream, whichever side of the guard a sample is on."""
    import torch
    import meters_lv2_b200 as B
    n = 4096
    x = _guard_signals(n)
    g = B.TruePeakKmeter(x.shape[0]); g.debug_capture(True)
    xd = torch.from_numpy(x).cuda()
    outs = [[] for _ in range(x.shape[0])]
    for o in range(0, n, block):
        k = min(block, n - o)
        g.process(xd[:, o:o + k])
        for ch in range(x.shape[0]):
            outs[ch].append(g.debug_upsampled(ch, 4 * k))
    for ch in range(x.shape[0]):
        ref = O.tp_upsample(x[ch], block=block)
        got = np.concatenate(outs[ch])
        nan = np.isnan(ref)                                                   # NaN payloads are not part of the contract
        assert np.array_equal(np.isnan(got), nan), ch
        assert np.array_equal(u32(got)[~nan], u32(ref)[~nan]), (ch, int((u32(got)[~nan] != u32(ref)[~nan]).sum()))


@pytest.mark.parametrize("mode", [0, 1])
def test_phase0_guard_meters(mode):
    x = _guard_signals(6000, seed=9)
    x = np.concatenate([x, _guard_signals(6000, seed=10)], axis=0)           # 24 rows: spans two process() tiles
    _drive(x, [1024, 1024, 1000, 24, 1024, 1904], read_every=2, mode=mode, flags=1 if mode else 3)


def test_phase0_guard_off_matches(monkeypatch):
    """B200M_TPK_ELIDE0=0 (full evaluation) and the default produce identical state on noise."""
    import torch
    import meters_lv2_b200 as B
    x = S.white(40, 4096, seed=21)
    xd = torch.from_numpy(x).cuda()
    res = []
    for v in ("1", "0"):
        monkeypatch.setenv("B200M_TPK_ELIDE0", v)
        g = B.TruePeakKmeter(40)
        for o in range(0, 4096, 1024):
            g.process(xd[:, o:o + 1024])
        res.append(g.state())
    for k in ("m", "p", "z1", "z2"):
        assert np.array_equal(u32(res[0][k]), u32(res[1][k])), k
