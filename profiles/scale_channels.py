"""How the K-weighting kernel and the true-peak FIR kernel scale with the number of channels per GPU (run under gpurun):
at the BASELINE batch (8192 stereo = 16384 channels) the K-weighting kernel has 512 warps for 592 SM sub-partitions."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import meters_lv2_b200 as B

NF, RING = 1024, 4


def t(fn, k=60):
    for s in range(10):
        fn(s)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
    for s in range(k):
        fn(s)
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / k * 1e3


for n in (2048, 4096, 8192, 16384, 32768):
    x = (torch.rand((2 * n, RING * NF), device="cuda") * 2 - 1) * 0.25
    base, stride = x.data_ptr(), x.stride(0)
    e = B.Ebu_r128_proc(n, 2, 48000.0); e.integr_start()
    us_e = t(lambda s: e.process_ptr(base + 4 * NF * (s % RING), stride, NF))
    tp = B.TruePeakKmeter(2 * n, 48000.0, flags=B.TPK_TRUEPEAK)
    us_t = t(lambda s: tp.process_ptr(base + 4 * NF * (s % RING), stride, NF, B.TP_MODE_MAX))
    b = B.EBUr128(n, 48000.0, True); b.control(B.EBUr128.START)
    us_r = t(lambda s: b.run_ptr(base + 4 * NF * (s % RING), stride, NF))
    samples = 2 * n * NF
    print("%6d stereo: EBU %7.1f us = %6.1f G samples/s (%.2f of 6486 GB/s) | FIR max %7.1f us = %6.1f G | EBUr128 cycle %7.1f us = %6.1f G samples/s"
          % (n, us_e, samples / us_e / 1e3, samples * 4 / us_e / 1e3 / 6486.5, us_t, samples / us_t / 1e3, us_r, samples / us_r / 1e3), flush=True)
    del e, tp, b, x
