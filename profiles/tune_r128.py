"""Tuning probe (not a bench number): device-resident and host-path step time of the EBUr128 cycle for the
B200M_R128_CONCURRENT / B200M_R128_SLICES knobs, plus scalar vs packed FIR and the ALU probes."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import meters_lv2_b200 as B

N, NF, RING = 8192, 1024, 8
torch.manual_seed(3)
x = (torch.rand((2 * N, RING * NF), device="cuda") * 2 - 1) * 0.25
base, stride = x.data_ptr(), x.stride(0)
host = torch.empty((2 * N, 2 * NF), dtype=torch.float32).pin_memory(); host.copy_(x[:, :2 * NF].cpu())
print("fp32 probe %.0f  fp32x2 probe %.0f  fp64 probe %.0f  (1e9 lane-ops/s)" % (B.peak_probe(0), B.peak_probe(2), B.peak_probe(1)))


def dev_ms(bank, k=60):
    for s in range(10):
        bank.run_ptr(base + 4 * NF * (s % RING), stride, NF)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(k):
        bank.run_ptr(base + 4 * NF * (s % RING), stride, NF)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def host_ms(bank, k=30):
    res = np.empty(N, B.EBU_RESULT_DTYPE); tp = np.empty(N, np.float32)
    for s in range(3):
        bank.run_ptr(host.data_ptr() + 4 * NF * (s % 2), host.stride(0), NF, host=True); bank.results(out=res, tp=tp)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for s in range(k):
        bank.run_ptr(host.data_ptr() + 4 * NF * (s % 2), host.stride(0), NF, host=True); bank.results(out=res, tp=tp)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3


for conc in (0, 2):
    for kw in (4, 2, 1):
        os.environ["B200M_R128_CONCURRENT"] = str(conc); os.environ["B200M_R128_SLICES"] = "4"; os.environ["B200M_R128_K1_WARPS"] = str(kw)
        b = B.EBUr128(N, 48000.0, True); b.control(B.EBUr128.START)
        d = [dev_ms(b) for _ in range(3)]
        h = host_ms(b)
        print("concurrent=%d k1_warps=%d : device %s ms/step   host %.4f ms/step (%.2f Gsamples/s)" % (conc, kw, " ".join("%.4f" % v for v in d), h, 2 * N * NF / h / 1e6))
        b.close()
