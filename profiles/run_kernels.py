"""Tiny driver for ncu captures: a few full-size launches of each hot kernel (used by the profiling recipe in
profiles/README.md; never used for bench numbers)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import meters_lv2_b200 as B

which = sys.argv[1] if len(sys.argv) > 1 else "all"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
torch.manual_seed(1)
x = (torch.rand((16384, 4 * 1024), device="cuda") * 2 - 1) * 0.25
p, s = x.data_ptr(), x.stride(0)
if which in ("all", "ebu"):
    e = B.Ebu_r128_proc(8192, 2); e.integr_start()
    for i in range(n):
        e.process_ptr(p + 4096 * (i % 4), s, 1024)
if which in ("all", "tpmax"):
    t = B.TruePeakKmeter(16384, flags=B.TPK_TRUEPEAK)
    for i in range(n):
        t.process_ptr(p + 4096 * (i % 4), s, 1024, B.TP_MODE_MAX)
if which in ("all", "tpk"):
    t2 = B.TruePeakKmeter(16384)
    for i in range(n):
        t2.process_ptr(p + 4096 * (i % 4), s, 1024)
if which in ("all", "spec"):
    sp = B.Spectr30(4096, 2)
    for i in range(max(2, n // 3)):
        sp.process_ptr(p + 4096 * (i % 4), s, 1024)
if which in ("all", "pw"):
    pw = B.Phasewheel(2048, 1024); co = B.Stcorrdsp(2048)
    for i in range(n):
        co.process_ptr(p + 4096 * (i % 4), s, 1024); pw.process_ptr(p + 4096 * (i % 4), s, 1024)
if which in ("all", "pwf"):
    # C5 as bench.py runs it: fused feed (Stcorrdsp in its time-parallel mode + FFT ring append) + the 25 Hz analysis
    pwf = B.Phasewheel(2048, 1024); cof = B.Stcorrdsp(2048); cof.set_precision(B.PREC_FMA); pwf.attach_cor(cof)
    for i in range(n):
        pwf.process_ptr(p + 4096 * (i % 4), s, 1024)
torch.cuda.synchronize()
print("done", B.launch_count())
