"""LV2 facade, EBUr128: N plugin instances driven like a host does (run() of each instance once per 1024-frame cycle),
default synchronous banks of one vs B200M_LV2_BATCH (one shared bank, one cycle of latency).  Run under gpurun."""
import ctypes as C
import os
import struct
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import meters_lv2_b200 as B
from test_lv2_shim_gpu import descriptors, Plugin
from test_lv2_ebur128_gpu import MTR, cfg, obj, sequence

BLK, CAP = 1024, 8192
for n in (16, 64, 256):
    for mode in ("sync", "batch"):
        if mode == "batch":
            os.environ["B200M_LV2_BATCH"] = str(n)
        else:
            os.environ.pop("B200M_LV2_BATCH", None)
        mine, lib = descriptors(B.LIB_PATH)
        ps = [Plugin(mine["EBUr128"]) for _ in range(n)]
        rng = np.random.default_rng(1)
        bufs = [[rng.uniform(-0.5, 0.5, BLK).astype(np.float32) for _ in range(2)] for _ in range(n)]
        notes = [np.zeros(CAP, np.uint8) for _ in range(n)]
        start = sequence([cfg("UISETTINGS", 72.0), cfg("START", 0)]); empty = sequence([])
        for i, p in enumerate(ps):
            p.port(1, notes[i])
            for c in range(2):
                p.port(2 + 2 * c, bufs[i][c]); p.port(3 + 2 * c, bufs[i][c])

        def cycle(ctl):
            for i, p in enumerate(ps):
                notes[i][:8] = np.frombuffer(struct.pack("<II", CAP - 8, 0), np.uint8)
                p.port(0, ctl)
                p.run(BLK)
        cycle(start)
        for _ in range(5):
            cycle(empty)
        t0 = time.perf_counter(); k = 30
        for _ in range(k):
            cycle(empty)
        dt = (time.perf_counter() - t0) / k
        print("%4d instances %-5s : %8.3f ms per cycle = %7.1f us per instance  (real time budget of a 1024-frame cycle at 48 kHz: 21.3 ms)"
              % (n, mode, dt * 1e3, dt / n * 1e6), flush=True)
        for p in ps:
            p.close()
