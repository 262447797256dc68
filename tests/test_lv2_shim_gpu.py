"""GPU: the per-instance LV2 façade (lv2_descriptor / instantiate / connect_port / run / cleanup) driven the way
an LV2 host drives meters.so (robtk/jackwrap.c:531-544), compared with the oracle DSP objects plus the reference's
port glue (src/meters.cc:333-536, src/spectrumlv2.c:159-257) restated inline."""
import ctypes as C

import numpy as np
import pytest

import _oracle as O
import _signals as S

pytestmark = pytest.mark.gpu
URI = b"http://gareus.org/oss/lv2/meters#"


class Desc(C.Structure):
    pass


Desc._fields_ = [("URI", C.c_char_p),
                 ("instantiate", C.CFUNCTYPE(C.c_void_p, C.POINTER(Desc), C.c_double, C.c_char_p, C.c_void_p)),
                 ("connect_port", C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_void_p)),
                 ("activate", C.c_void_p),
                 ("run", C.CFUNCTYPE(None, C.c_void_p, C.c_uint32)),
                 ("deactivate", C.c_void_p),
                 ("cleanup", C.CFUNCTYPE(None, C.c_void_p)),
                 ("extension_data", C.c_void_p)]


def descriptors():
    import meters_lv2_b200 as B
    L = C.CDLL(B.LIB_PATH)
    L.lv2_descriptor.restype = C.POINTER(Desc); L.lv2_descriptor.argtypes = [C.c_uint32]
    out, i = {}, 0
    while True:
        d = L.lv2_descriptor(i)
        if not d:
            break
        out[d.contents.URI[len(URI):].decode()] = d
        i += 1
    return out


class Plugin:
    def __init__(self, d, rate=48000.0):
        self.d = d.contents
        self.h = self.d.instantiate(d, rate, b"", None)
        assert self.h
        self.bufs = {}

    def port(self, idx, arr):
        self.bufs[idx] = arr
        self.d.connect_port(self.h, idx, arr.ctypes.data_as(C.c_void_p))

    def run(self, n):
        self.d.run(self.h, n)

    def close(self):
        self.d.cleanup(self.h)


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_descriptor_table():
    d = descriptors()
    assert set(d) == {"COR", "spectr30mono", "spectr30stereo", "dBTPmono", "dBTPstereo", "K12mono", "K14mono", "K20mono",
                      "K12stereo", "K14stereo", "K20stereo"}


def test_k20stereo_and_dbtp_and_cor_ports():
    d = descriptors()
    x = S.white(2, 1024 * 20, seed=91)
    ctl = {k: np.zeros(1, np.float32) for k in range(10)}
    # K20stereo: 0 ref, 1 in0, 2 out0, 3 level0, 4 in1, 5 out1, 6 level1, 7 peak0, 8 peak1, 9 hold
    k = Plugin(d["K20stereo"]); t = Plugin(d["dBTPstereo"]); c = Plugin(d["COR"])
    ports = {}
    for name, p in (("k", k), ("t", t), ("c", c)):
        ports[name] = {i: np.zeros(1, np.float32) for i in (0, 3, 6, 7, 8, 9)}
        for i, a in ports[name].items():
            p.port(i, a)
    ok = O.Kmeter(2); ot = O.TruePeak(2); oc = O.Stcorr(1)
    hold = 0.0; pmax = np.zeros(2, np.float32)
    for b in range(20):
        l = np.ascontiguousarray(x[0, b * 1024:(b + 1) * 1024]); r = np.ascontiguousarray(x[1, b * 1024:(b + 1) * 1024])
        for p in (k, t, c):
            p.port(1, l); p.port(2, l); p.port(4, r); p.port(5, r)       # in-place: out == in
            p.run(1024)
        blk = np.ascontiguousarray(x[:, b * 1024:(b + 1) * 1024])
        ok.process(blk); ot.process(blk); oc.process(blk)
        assert u32(ports["c"][3])[0] == u32(oc.read())[0]
        if b == 0:
            # p_refl starts at -9999 != *ref (0): the first run() of the dBTP / K-meter plugins is a peak-reset handshake
            # cycle: reset, process, force a port change, return WITHOUT read() (src/meters.cc:339-357,381-389,444-489)
            assert ports["k"][9][0] <= -1.0 and ports["t"][3][0] <= -500.0
            continue
        rms, pk = ok.read(); m, pp = ot.read()
        hold = max(hold, float(pk.max())); pmax = np.maximum(pmax, pp)
        assert u32(ports["k"][3])[0] == u32(rms[:1])[0] and u32(ports["k"][6])[0] == u32(rms[1:])[0]
        assert u32(ports["k"][7])[0] == u32(pk[:1])[0] and u32(ports["k"][8])[0] == u32(pk[1:])[0]
        assert ports["k"][9][0] == np.float32(hold)
        assert u32(ports["t"][3])[0] == u32(m[:1])[0] and u32(ports["t"][6])[0] == u32(m[1:])[0]
        assert u32(ports["t"][7])[0] == u32(pmax[:1])[0] and u32(ports["t"][8])[0] == u32(pmax[1:])[0]
    # peak-reset handshake (port 0 re-used, src/meters.cc:339-357): |ref| < 3 resets, ports get a forced change
    ports["k"][0][0] = 1.0
    k.run(1024)
    assert ports["k"][9][0] <= -1.0
    for p in (k, t, c):
        p.close()


def test_spectr30stereo_ports():
    d = descriptors()
    x = S.white(2, 1024 * 6, seed=92)
    p = Plugin(d["spectr30stereo"])
    out = np.zeros(60, np.float32); spd = np.ones(1, np.float32); rst = np.full(1, -4.0, np.float32); amp = np.zeros(1, np.float32)
    for i in range(60):
        p.d.connect_port(p.h, i, out[i:i + 1].ctypes.data_as(C.c_void_p))
    p.port(60, spd); p.port(61, rst); p.port(62, amp)
    o = O.Spectr30(1, 2)
    for b in range(6):
        l = np.ascontiguousarray(x[0, b * 1024:(b + 1) * 1024]); r = np.ascontiguousarray(x[1, b * 1024:(b + 1) * 1024])
        p.port(64, l); p.port(65, l); p.port(66, r); p.port(67, r)
        p.run(1024)
        o.process(np.ascontiguousarray(x[:, b * 1024:(b + 1) * 1024]))
        assert np.array_equal(u32(out), u32(o.read()[0]))
    p.close()
