#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native metering engine (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Workload (BASELINE.json `metric`: "audio-samples/sec/GPU (48 kHz stereo, 8192-ch batch) EBU R128 + true-peak"):
8192 stereo instances PER GPU (weak scaling), 48 kHz, 1024-frame blocks, the EBUr128 plugin's audio cycle with
integration running and dBTP enabled (src/ebulv2.cc:341-367: Ebu_r128_proc::process + TruePeakdsp::process_max on
both channels + getters).  One "step" = one such cycle over the whole batch = 16 777 216 mono samples per GPU.
`value` = samples/s with the input resident in HBM (ring of 8 distinct 64 MiB blocks > L2); `e2e` = the same
cycle through b200m_r128_run_host with pinned HOST buffers, H2D copy and D2H result read inside the timed region.
`--impl reference` times the reference's own CPU code (oracle/_ref, else the oracle port) on all host threads.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "audio-samples/sec/GPU (48 kHz stereo, 8192-ch batch) EBU R128 + true-peak"
FS = 48000.0
N_INST = 8192          # stereo instances per GPU
NFRAM = 1024
RING = 8               # distinct device-resident blocks: 8 x 64 MiB = 512 MiB > 126 MB L2
PRIME = 480            # untimed blocks (10.2 s of audio) so that S, I (>=50 M-points) and LRA (>=20 S-points) are live
SAMPLES_PER_STEP = N_INST * 2 * NFRAM


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_for(kernel):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(kernel)
    return None


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons of one GPU through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.maxc = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.maxc = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        names = {"GpuIdle": 0x1, "ApplicationsClocksSetting": 0x2, "sw_power_cap": 0x4, "hw_slowdown": 0x8,
                 "SyncBoost": 0x10, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80}
        while not self.stop_flag:
            try:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, v in names.items():
                    if r & v and k != "GpuIdle":
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.maxc, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.maxc, "reasons": sorted(self.reasons), "samples": len(self.sm)}


def dist_env():
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), ws


def make_ring(torch, dev, rank):
    g = torch.Generator(device=dev); g.manual_seed(0x42B200 + rank)
    c = torch.arange(2 * N_INST, device=dev)
    gain = torch.pow(10.0, -(6.0 + 30.0 * (c % 97).float() / 96.0) / 20.0)
    x = (torch.rand((2 * N_INST, RING * NFRAM), generator=g, device=dev, dtype=torch.float32) * 2 - 1) * gain[:, None]
    return x.contiguous()


def timed_loop(torch, dist, ws, fn, steps):
    """barrier + synchronize on both sides, CUDA events on the current stream, max over ranks (ms)"""
    if ws > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        fn(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if ws > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    return ms


def run_b200(args):
    import torch
    import torch.distributed as dist
    import meters_lv2_b200 as B
    rank, local, ws = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if ws > 1:
        os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    hbm_peak, peak_src = peaks()
    x = make_ring(torch, dev, rank)
    stride = x.stride(0)
    base = x.data_ptr()

    bank = B.EBUr128(N_INST, FS, dbtp_enable=True, device=local)
    bank.control(B.EBUr128.START)
    pos = [0]

    def step(_):
        b = pos[0] % RING
        bank.run_ptr(base + 4 * NFRAM * b, stride, NFRAM)
        pos[0] += 1

    for s in range(PRIME + W):
        step(s)
    torch.cuda.synchronize()
    sampler = ClockSampler(local); sampler.start()
    l0 = B.launch_count()
    ms = timed_loop(torch, dist, ws, step, K)
    launches = B.launch_count() - l0
    sampler.stop_flag = True; sampler.join()
    value = ws * SAMPLES_PER_STEP * K / (ms * 1e-3)

    out = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": ws, "steps": K, "warmup": W,
           "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "impl": "b200",
           "config": {"workload": "8192 stereo EBU R128 M+S+I (integrating) + dBTP true-peak 4x (ebur128_run audio cycle) per GPU",
                      "instances_per_gpu": N_INST, "channels_per_instance": 2, "block": NFRAM, "fs": FS,
                      "input": "device ring of %d distinct 64 MiB blocks (512 MiB > L2), no L2 flush needed" % RING,
                      "prime_blocks": PRIME, "parallelism": "channel-shard x%d, no data-path collective" % ws},
           "value_per_gpu": value / ws, "gpu_launches": int(launches), "clocks": sampler.summary()}

    # ---- whole-mix gated loudness (the path's one exchange): per-GPU histogram sum -> ONE int32[1508] all-reduce
    # (NCCL when N > 1) -> calc_integ / calc_range on the sum.  Not part of the timed cycle: it is due once per 0.5 s of audio.
    try:
        from meters_lv2_b200 import shard
        mixv = torch.zeros(B.MIX_WORDS, dtype=torch.int32, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):                                   # first calls set up the communicator
            bank.ebu.mix_reduce(mixv); shard.allreduce_mix(mixv)
        torch.cuda.synchronize(); e0.record()
        bank.ebu.mix_reduce(mixv); shard.allreduce_mix(mixv)
        e1.record(); torch.cuda.synchronize()
        mo = bank.ebu.mix_finish(mixv)
        cm = int(mixv[2 * 752].item()); res0, _tp0 = bank.results()
        out["whole_mix"] = {"integrated": float(mo[0]), "range_min": float(mo[2]), "range_max": float(mo[3]),
                            "hist_M_points": cm, "hist_M_points_rank0_times_n": int(res0["hist_M_count"].astype(np.int64).sum()) * ws,
                            "reduce_plus_allreduce_us": e0.elapsed_time(e1) * 1e3, "collective": "nccl all_reduce int32[%d]" % B.MIX_WORDS if ws > 1 else "none (N = 1)"}
    except Exception as e:
        out["whole_mix"] = {"error": repr(e)}

    # ---- parity spot check against the CPU oracle on the first instances (same block sequence) ----------------
    if rank == 0:
        try:
            import _oracle as O
            ni = 2
            xs = x[:2 * ni].cpu().numpy()
            oe = O.Ebu(ni, 2, FS); ot = O.TruePeak(2 * ni, FS); oe.integr("start")
            tpmax = np.full(ni, -np.inf, np.float32)
            for s in range(PRIME + W + K):
                b = s % RING
                blk = np.ascontiguousarray(xs[:, b * NFRAM:(b + 1) * NFRAM])
                oe.process(blk); ot.process(blk, mode=1)
                m, _ = ot.read()
                v = np.maximum(m[0::2], m[1::2])
                with np.errstate(divide="ignore"):
                    tp = np.where(v == 0, -np.inf, (20.0 * np.log10(v.astype(np.float32)).astype(np.float64)).astype(np.float32))
                tpmax = np.maximum(tpmax, tp)
            res, tpg = bank.results()
            orr = oe.read()
            names = ("loudness_M", "maxloudn_M", "loudness_S", "maxloudn_S", "integrated", "integ_thr", "range_min", "range_max", "range_thr")
            exact = all(np.array_equal(res[n][:ni].view(np.uint32), orr[:, i].view(np.uint32)) for i, n in enumerate(names))
            hm, hs = bank.ebu.histogram(0); om, os_, _ = oe.hist(0)
            out["parity"] = {"oracle": O.load().orc_kind().decode(), "instances_checked": ni,
                             "ebu_bit_exact": bool(exact), "hist_bit_exact": bool(np.array_equal(hm, om) and np.array_equal(hs, os_)),
                             "dbtp_max_abs_diff_db": float(np.max(np.abs(tpg[:ni] - tpmax)))}
        except Exception as e:  # the bench number stands on its own; tests are the parity gate
            out["parity"] = {"error": repr(e)}

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ---------------------------------
    hbank = B.EBUr128(N_INST, FS, dbtp_enable=True, device=local)
    hbank.control(B.EBUr128.START)
    HR = 2
    # HR separate dense [channels][1024] blocks from the library's pinned allocator (GPU-local NUMA node): what a host
    # that double-buffers its capture hands over each cycle; dense blocks go over PCIe as one DMA per slice
    hosts = [B.host_alloc(2 * N_INST, NFRAM) for _ in range(HR)]
    for i, hb in enumerate(hosts):
        hb[:] = x[:, i * NFRAM:(i + 1) * NFRAM].cpu().numpy()
    res_buf = np.empty(N_INST, B.EBU_RESULT_DTYPE); tp_buf = np.empty(N_INST, np.float32)
    hptrs = [hb.ctypes.data for hb in hosts]
    ke = max(3, min(K, args.e2e_steps))

    def estep(s):
        hbank.run_ptr(hptrs[s % HR], NFRAM, NFRAM, host=True)
        hbank.results(out=res_buf, tp=tp_buf)

    for s in range(3):
        estep(s)
    if ws > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(ke):
        estep(s)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if ws > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    out["e2e"] = {"value": ws * SAMPLES_PER_STEP * ke / dt, "unit": "samples/s", "steps": ke,
                  "h2d_bytes_per_step": 2 * N_INST * NFRAM * 4, "d2h_bytes_per_step": int(res_buf.nbytes + tp_buf.nbytes),
                  "api": "b200m_r128_run_host + b200m_r128_results (pinned host buffers from b200m_host_alloc)"}
    del hbank

    # ---- per-kernel timings for the roofline (kernel alone, same ring, CUDA events) -----------------------------
    tpb = B.TruePeakKmeter(2 * N_INST, FS, flags=B.TPK_TRUEPEAK, device=local)
    ebb = B.Ebu_r128_proc(N_INST, 2, FS, device=local); ebb.integr_start()
    for s in range(W):
        tpb.process_ptr(base + 4 * NFRAM * (s % RING), stride, NFRAM, B.TP_MODE_MAX)
        ebb.process_ptr(base + 4 * NFRAM * (s % RING), stride, NFRAM)
    ms_tp = timed_loop(torch, dist, ws, lambda s: tpb.process_ptr(base + 4 * NFRAM * (s % RING), stride, NFRAM, B.TP_MODE_MAX), K)
    l1 = B.launch_count()
    ms_eb = timed_loop(torch, dist, ws, lambda s: ebb.process_ptr(base + 4 * NFRAM * (s % RING), stride, NFRAM), K)
    eb_launch = B.launch_count() - l1
    fp32_peak = B.peak_probe(0, local)
    alg_bytes = SAMPLES_PER_STEP * 4.0                      # 4 B per mono sample read once; outputs ~0
    tp_gbs = alg_bytes / (ms_tp / K * 1e-3) / 1e9
    eb_gbs = alg_bytes / (ms_eb / K * 1e-3) / 1e9
    # fp32 ops the FIR executes per input sample, unfused (bit-exact): phases 1-3 always (3 x 24 x (2 FMUL + 2 FADD) = 288);
    # phase 0 (96 more) only where the exact-delay guard of csrc/tpk.cu fails (never on this noise input).  The
    # guard's own ~9 ops/sample are not counted.  `full_eval` = the same kernel forced to evaluate all 384 ops/sample
    # (B200M_TPK_ELIDE0=0): the data-independent worst case (digital silence, sparse impulses).
    os.environ["B200M_TPK_ELIDE0"] = "0"
    tpf = B.TruePeakKmeter(2 * N_INST, FS, flags=B.TPK_TRUEPEAK, device=local)
    del os.environ["B200M_TPK_ELIDE0"]
    for s in range(W):
        tpf.process_ptr(base + 4 * NFRAM * (s % RING), stride, NFRAM, B.TP_MODE_MAX)
    ms_tpf = timed_loop(torch, dist, ws, lambda s: tpf.process_ptr(base + 4 * NFRAM * (s % RING), stride, NFRAM, B.TP_MODE_MAX), K)
    del tpf
    fir_ops = SAMPLES_PER_STEP * 288.0
    out["roofline"] = {"kernel": "tpk_kernel<TP,MAX> (4x polyphase FIR + max)", "bound": "hbm", "achieved": tp_gbs, "peak": hbm_peak,
                       "unit": "GB/s", "frac": tp_gbs / hbm_peak, "traffic": traffic_for("tpk_kernel"), "peak_source": peak_src,
                       "ms_per_launch": ms_tp / K,
                       "note": "this kernel is fp32-issue bound (288-384 unfused FMUL/FADD per sample), see roofline_alu"}
    out["roofline_alu"] = {"kernel": "tpk_kernel<TP,MAX>", "bound": "fp32 issue (unfused mul+add)", "achieved": fir_ops / (ms_tp / K * 1e-3) / 1e9,
                           "peak": fp32_peak, "unit": "1e9 lane-ops/s", "frac": fir_ops / (ms_tp / K * 1e-3) / 1e9 / fp32_peak,
                           "ops_per_sample": 288, "peak_source": "b200m_peak_probe(0) measured in this run",
                           "full_eval": {"ops_per_sample": 384, "ms_per_launch": ms_tpf / K,
                                         "frac": SAMPLES_PER_STEP * 384.0 / (ms_tpf / K * 1e-3) / 1e9 / fp32_peak}}
    out["roofline_kernels"] = [
        {"kernel": "ebu_kweight_frag (+ebu_loudness_hist every 2400 frames)", "bound": "hbm", "achieved": eb_gbs, "peak": hbm_peak, "unit": "GB/s",
         "frac": eb_gbs / hbm_peak, "traffic": traffic_for("ebu_kweight_frag"), "ms_per_block": ms_eb / K, "launches_per_block": eb_launch / K,
         "samples_per_s": SAMPLES_PER_STEP * K / (ms_eb * 1e-3)}]
    del tpb, ebb

    # ---- the other BASELINE configs, single GPU only (reported, not the headline) -------------------------------
    if ws == 1 and not args.headline_only:
        out["configs"] = other_configs(torch, dist, B, x, K, W, hbm_peak)

    # ---- CPU baseline beside it (rank 0, N = 1) -----------------------------------------------------------------
    if rank == 0 and ws == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(budget_s=12.0)
    if rank == 0:
        print(json.dumps(out))
    if ws > 1:
        dist.destroy_process_group()


def other_configs(torch, dist, B, x, K, W, hbm_peak):
    base, stride = x.data_ptr(), x.stride(0)
    cfg = {}
    fp64_peak = B.peak_probe(1)

    def blk(s):
        return base + 4 * NFRAM * (s % RING)

    # C2 pure: EBU R128 only
    e = B.Ebu_r128_proc(N_INST, 2, FS); e.integr_start()
    for s in range(PRIME // 4 + W):
        e.process_ptr(blk(s), stride, NFRAM)
    ms = timed_loop(torch, dist, 1, lambda s: e.process_ptr(blk(s), stride, NFRAM), K)
    n = N_INST * 2 * NFRAM
    cfg["C2_ebu_r128_8192st"] = {"samples_per_s": n * K / (ms * 1e-3), "ms_per_block": ms / K, "hbm_frac": n * 4 * K / (ms * 1e-3) / 1e9 / hbm_peak}
    del e
    # the same bank at 4x the BASELINE batch: 8192 stereo instances are 512 warps for 592 SM sub-partitions (latency bound, the
    # block time is flat from 2048 to 8192 instances); with 32768 the kernel streams (profiles/r1_scale_channels.txt)
    try:
        n4 = 4 * N_INST
        x4 = (torch.rand((2 * n4, 4 * NFRAM), device=x.device, dtype=torch.float32) * 2 - 1) * 0.25
        e4 = B.Ebu_r128_proc(n4, 2, FS); e4.integr_start()
        b4, s4 = x4.data_ptr(), x4.stride(0)
        for s in range(W + 4):
            e4.process_ptr(b4 + 4 * NFRAM * (s % 4), s4, NFRAM)
        k4x = max(10, K // 4)
        ms = timed_loop(torch, dist, 1, lambda s: e4.process_ptr(b4 + 4 * NFRAM * (s % 4), s4, NFRAM), k4x)
        cfg["C2x4_ebu_r128_32768st"] = {"samples_per_s": 4 * n * k4x / (ms * 1e-3), "ms_per_block": ms / k4x,
                                        "hbm_frac": 4 * n * 4 * k4x / (ms * 1e-3) / 1e9 / hbm_peak, "note": "not a BASELINE config: shows the kernel's bandwidth when the chip is filled"}
        del e4, x4
    except Exception as ex:  # an extra, never fatal
        cfg["C2x4_ebu_r128_32768st"] = {"error": repr(ex)}
    # C3: true peak (process) + K-meter, read every block (TPnRMS, src/dr14.c:391-450)
    t = B.TruePeakKmeter(2 * N_INST, FS)

    def c3(s):
        t.process_ptr(blk(s), stride, NFRAM); t.read_device()
    for s in range(W):
        c3(s)
    ms = timed_loop(torch, dist, 1, c3, K)
    cfg["C3_truepeak_k20_8192st"] = {"samples_per_s": n * K / (ms * 1e-3), "ms_per_block": ms / K, "hbm_frac": n * 4 * K / (ms * 1e-3) / 1e9 / hbm_peak,
                                     "fp32_issue_frac": n * 336.0 * K / (ms * 1e-3) / 1e9 / B.peak_probe(0)}
    del t
    # C4: 4096 stereo 30-band spectrum (unit: stereo frames)
    sp = B.Spectr30(4096, 2, FS)
    k4 = max(3, K // 10)
    for s in range(2):
        sp.process_ptr(blk(s), stride, NFRAM)
    ms = timed_loop(torch, dist, 1, lambda s: sp.process_ptr(blk(s), stride, NFRAM), k4)
    fr = 4096 * NFRAM
    cfg["C4_spectr30_4096st"] = {"frames_per_s": fr * k4 / (ms * 1e-3), "ms_per_block": ms / k4, "hbm_frac": fr * 8 * k4 / (ms * 1e-3) / 1e9 / hbm_peak,
                                 "fp64_frac": fr * 30 * 39.0 * k4 / (ms * 1e-3) / 1e9 / fp64_peak, "fp64_peak_glops": fp64_peak}
    del sp
    # C5: 2048 stereo phasewheel 2048-pt FFT + Stcorr (unit: stereo frames)
    pw = B.Phasewheel(2048, 1024, FS); co = B.Stcorrdsp(2048, int(FS))

    # two independent banks over the same input ring, each on its own stream (the latency-bound correlation kernel, 64 warps,
    # runs beside the FFT kernels); the streams fork at the first block of a loop and join at its last
    side = torch.cuda.Stream()
    k5 = K - (K % 2)

    def c5(s, last=None):
        cur = torch.cuda.current_stream()
        if s == 0:
            side.wait_stream(cur)
        co.process_ptr(blk(s), stride, NFRAM, stream=side)
        pw.process_ptr(blk(s), stride, NFRAM)
        if s == (k5 - 1 if last is None else last):
            cur.wait_stream(side)
    for s in range(W + 1):
        c5(s, last=W)
    ms = timed_loop(torch, dist, 1, c5, k5)
    fr = 2048 * NFRAM
    cfg["C5_phasewheel_stcorr_2048st"] = {"frames_per_s": fr * k5 / (ms * 1e-3), "ms_per_block": ms / k5, "hbm_frac": fr * 12 * k5 / (ms * 1e-3) / 1e9 / hbm_peak}
    return cfg


def cpu_baseline(budget_s=12.0, steps=None, warmup=1):
    """the reference's CPU code (oracle/_ref if present, else the port) on all host threads, bounded sample"""
    import _oracle as O
    L = O.load("best")
    kind = L.orc_kind().decode()
    threads = max(1, L.orc_hw_threads())
    n_s = int(min(N_INST, 64 * threads))
    blocks = 4
    rng = np.random.Generator(np.random.Philox(key=0x42B200))
    x = (rng.random((2 * n_s, blocks * NFRAM), dtype=np.float32) * 2 - 1)
    c = np.arange(2 * n_s)
    x *= (10.0 ** (-(6.0 + 30.0 * (c % 97) / 96.0) / 20.0)).astype(np.float32)[:, None]
    e = O.Ebu(n_s, 2, FS, kind="best"); t = O.TruePeak(2 * n_s, FS, kind="best"); e.integr("start")
    per_step = n_s * 2 * NFRAM * blocks
    for _ in range(warmup):
        O.r128_cycle(e, t, x, NFRAM, blocks, threads)
    t0 = time.perf_counter(); k = 0
    while True:
        O.r128_cycle(e, t, x, NFRAM, blocks, threads); k += 1
        dt = time.perf_counter() - t0
        if (steps is not None and k >= steps) or (steps is None and dt > budget_s):
            break
    return {"value": per_step * k / dt, "unit": "samples/s", "cores": threads, "kind": "reference" if kind == "reference" else "port",
            "sample": "%d of %d stereo instances x %d blocks of %d frames per step, %d steps, %.1f s" % (n_s, N_INST, blocks, NFRAM, k, dt),
            "ms_per_step": dt / k * 1e3, "steps": k}


def run_reference(args):
    rank, local, ws = dist_env()
    if rank != 0:
        return
    K, W = args.steps, max(args.warmup, 1)
    cb = cpu_baseline(steps=K, warmup=W)
    out = {"metric": METRIC, "value": cb["value"], "unit": "samples/s", "n_gpus": ws, "steps": K, "warmup": W,
           "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "impl": "reference",
           "config": {"workload": "8192 stereo EBU R128 M+S+I (integrating) + dBTP true-peak 4x (ebur128_run audio cycle), CPU: " + cb["sample"],
                      "block": NFRAM, "fs": FS},
           "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
           "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--headline-only", action="store_true", help="skip the other BASELINE configs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
