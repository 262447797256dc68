"""Generates tests/golden/golden_v1.npz from the REFERENCE build (oracle/_ref, i.e. the unmodified
x42/meters.lv2 sources compiled by oracle/Makefile) on small seeded streams.

    python tests/golden/make_golden.py            # rewrite the fixture (needs /root/reference -> oracle/_ref)

The reference ships no golden vectors or tests, so these fixtures — outputs of the
reference itself, run here — are the pin for the oracle port on machines where /root/reference is absent.
`compute(kind)` is also what tests/test_oracle_port.py::test_golden_vectors replays.
The phasewheel entries come from the port (FFTW3 is absent: that path has no reference build).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as O          # noqa: E402
import _signals as S         # noqa: E402

BLOCKS = [1024] * 130 + [480, 64, 8192, 1, 3, 1023]


def compute(kind):
    out = {}
    kind_o = kind
    x = S.white(8, sum(BLOCKS), seed=0xC0FFEE)
    x[6, 5000:5010] = [np.nan, np.inf, -np.inf, 1e-42, 0, 1, -1, 1e-30, 3e38, -3e38]
    e = O.Ebu(4, 2, kind=kind); e.integr("start")
    tp = O.TruePeak(8, kind=kind); km = O.Kmeter(8, kind=kind); co = O.Stcorr(4, kind=kind)
    sp = O.Spectr30(2, 2, kind=kind)
    tpm, kmr, cor = [], [], []
    pos = 0
    for i, n in enumerate(BLOCKS):
        blk = np.ascontiguousarray(x[:, pos:pos + n]); pos += n
        e.process(blk); tp.process(blk); km.process(blk); co.process(blk)
        if i < 24:
            sp.process(np.ascontiguousarray(blk[:4]))
        m, p = tp.read(); r, pk = km.read()
        tpm.append(np.stack([m, p])); kmr.append(np.stack([r, pk])); cor.append(co.read())
    out["ebu_results"] = e.read()
    for i in range(4):
        hm, hs, c = e.hist(i)
        out["ebu_histM_%d" % i], out["ebu_histS_%d" % i], out["ebu_counts_%d" % i] = hm, hs, c
    out["ebu_coeffs"] = e.coeffs()
    out["tp_reads"] = np.stack(tpm); out["km_reads"] = np.stack(kmr); out["cor_reads"] = np.stack(cor)
    w, t = tp.coeffs(); out["tp_w"], out["tp_ctab"] = w, t
    out["tp_upsampled"] = O.tp_upsample(x[0, :2048], kind=kind)
    ports = sp.read()
    out["spec_ports"] = ports[:, :30].copy(); out["spec_maxports"] = ports[:, 30:].copy()
    out["spec_coeffs"] = sp.coeffs()
    # needle-meter ballistics (VU, IEC I, IEC II, M/S PPM)
    for nk in range(4):
        nm = O.Needle(4, nk, oracle=kind_o)
        reads = []
        pos2 = 0
        for i, n in enumerate(BLOCKS[:40]):
            blk = np.ascontiguousarray((x[:8 if nk == 3 else 4, pos2:pos2 + n] * np.float32(3.0))); pos2 += n
            nm.process(blk); reads.append(nm.read())
        out["needle_reads_%d" % nk] = np.stack(reads)
        out["needle_state_%d" % nk] = nm.peek()
    # bit-meter (cumulative mode) and signal distribution histogram
    bm = O.Bitmeter(4, oracle=kind_o); bm.mode(1)
    sd = O.SigDist(4, oracle=kind_o); sd.integrate(True)
    pos3 = 0
    for n in BLOCKS[:30]:
        blk = np.ascontiguousarray(x[4:8, pos3:pos3 + n]); pos3 += n
        bm.process(blk); sd.process(blk)
    for i in (0, 2):
        h, c, mm, it = bm.read(i)
        out["bim_hist_%d" % i], out["bim_cnt_%d" % i], out["bim_minmax_%d" % i] = h, c, mm
        h, mp, av, it = sd.read(i)
        out["sdh_hist_%d" % i], out["sdh_maxpeak_%d" % i], out["sdh_stats_%d" % i] = h, mp, av
    # 997 Hz / -23 dBFS tone (EBU Tech 3341 case 1)
    s = S.sine(1024 * 300, 997.0, amp=10 ** (-23 / 20)); y = np.ascontiguousarray(np.stack([s, s]))
    e2 = O.Ebu(1, 2, kind=kind); e2.integr("start")
    for b in range(300):
        e2.process(np.ascontiguousarray(y[:, b * 1024:(b + 1) * 1024]))
    out["ebu_tone_results"] = e2.read()
    # DR-14 (stereo, 36 s in 8192-frame blocks: 12 scored windows) and TPnRMS (mono)
    t = np.arange(8192 * 212) / 48000.0
    env = (0.2 + 0.8 * np.abs(np.sin(2 * np.pi * t / 6.1))).astype(np.float32)
    xd = (S.white(4, 8192 * 212, seed=0xD214) * env * np.float32(2.5)).astype(np.float32)
    xd[2:4, 8192 * 60:8192 * 110] = 0.0                                    # instance 1: 3 s windows of silence are not scored
    dr = O.Dr14(2, 2, 48000.0, True, kind=kind)
    reads = []
    for b in range(212):
        dr.process(np.ascontiguousarray(xd[:, b * 8192:(b + 1) * 8192]))
        if b % 4 == 3:
            reads.append(dr.read())
        if b == 150:
            dr.reset()
    out["dr14_ports"] = np.stack(reads)
    tn = O.Dr14(3, 1, 44100.0, False, kind=kind)
    reads = []
    for b in range(40):
        tn.process(np.ascontiguousarray(x[:3, b * 1000:(b + 1) * 1000]))
        reads.append(tn.read())
    out["tpnrms_ports"] = np.stack(reads)
    # phasewheel (port only)
    pw = O.Phasewheel(2, 1024, kind="port")
    for b in range(4):
        pw.process(np.ascontiguousarray(x[:4, b * 1024:(b + 1) * 1024]))
    ph, lv, pk = pw.read()
    out["pw_phase"], out["pw_level"], out["pw_peak"] = ph, lv, pk
    return out


if __name__ == "__main__":
    assert O.available("reference"), "build oracle/_ref first (make -C oracle ref)"
    d = compute("reference")
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **d)
    print("wrote golden_v1.npz:", {k: v.shape for k, v in d.items()})
