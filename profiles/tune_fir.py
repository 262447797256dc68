"""Timing probe for the true-peak FIR kernel: phase-0 shortcut on/off x input kind (run under gpurun)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import meters_lv2_b200 as B

C, N, RING = 16384, 1024, 8
g = torch.Generator(device="cuda"); g.manual_seed(1)
inputs = {
    "uniform": (torch.rand((C, RING * N), generator=g, device="cuda") * 2 - 1) * 0.25,
    "normal": torch.randn((C, RING * N), generator=g, device="cuda") * 0.1,
    "zeros": torch.zeros((C, RING * N), device="cuda"),
}
for mode, mname in ((B.TP_MODE_MAX, "max"), (0, "process")):
    for kind, x in inputs.items():
        for el in ("1", "0", "2"):
            os.environ["B200M_TPK_ELIDE0"] = el
            t = B.TruePeakKmeter(C, 48000.0, flags=B.TPK_TRUEPEAK)
            base, stride = x.data_ptr(), x.stride(0)
            for s in range(5):
                t.process_ptr(base + 4 * N * (s % RING), stride, N, mode)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for s in range(50):
                t.process_ptr(base + 4 * N * (s % RING), stride, N, mode)
            e1.record(); torch.cuda.synchronize()
            print("%-8s %-8s elide0=%s  %.1f us/launch" % (mname, kind, el, e0.elapsed_time(e1) / 50 * 1e3), flush=True)
            del t
