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


for conc in (0, 1):
    for sl in (1, 2, 4, 8):
        os.environ["B200M_R128_CONCURRENT"] = str(conc); os.environ["B200M_R128_SLICES"] = str(sl)
        b = B.EBUr128(N, 48000.0, True); b.control(B.EBUr128.START)
        d = dev_ms(b) if sl == 1 else float("nan")
        h = host_ms(b)
        print("concurrent=%d slices=%d : device %.4f ms/step   host %.4f ms/step (%.2f Gsamples/s)" % (conc, sl, d, h, 2 * N * NF / h / 1e6))
        b.close()

# plain pinned H2D copy rate of one 64 MiB block, for reference
dst = torch.empty((2 * N, NF), device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    dst.copy_(host[:, :NF], non_blocking=True)
torch.cuda.synchronize()
print("torch pinned H2D of [16384 x 1024] strided rows: %.3f ms per block" % ((time.perf_counter() - t0) / 20 * 1e3))
hc = torch.empty((2 * N, NF), dtype=torch.float32).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    dst.copy_(hc, non_blocking=True)
torch.cuda.synchronize()
print("torch pinned H2D contiguous 64 MiB: %.3f ms per block" % ((time.perf_counter() - t0) / 20 * 1e3))


# true-peak kernels: process_max (FIR only) and process + K-meter (lock-step vs warp-specialised pipeline)
def tp_ms(bank, mode, read, k=40):
    for s in range(5):
        bank.process_ptr(base + 4 * NF * (s % RING), stride, NF, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(k):
        bank.process_ptr(base + 4 * NF * (s % RING), stride, NF, mode)
        if read:
            bank.read_device()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


for imm in (0, 1):
    os.environ["B200M_TPK_IMM"] = str(imm)
    t = B.TruePeakKmeter(2 * N, flags=B.TPK_TRUEPEAK)
    print("FIR process_max imm=%d : %.4f ms/block" % (imm, tp_ms(t, B.TP_MODE_MAX, False)))
    t2 = B.TruePeakKmeter(2 * N)
    print("TP+K20 process imm=%d : %.4f ms/block" % (imm, tp_ms(t2, B.TP_MODE_PROCESS, True)))
    del t, t2
e = B.Ebu_r128_proc(N, 2); e.integr_start()
for s in range(300):
    e.process_ptr(base + 4 * NF * (s % RING), stride, NF)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(117):
    e.process_ptr(base + 4 * NF * (s % RING), stride, NF)
e1.record(); torch.cuda.synchronize()
print("EBU R128 only : %.4f ms/block" % (e0.elapsed_time(e1) / 117))
