"""Turns an .ncu-rep (ncu --set full) into the short text summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof3_ebu.ncu-rep [more.ncu-rep ...] > profiles/r1_xxx.txt

Reads the report here (no GPU needed) with `ncu -i ... --page raw --csv`; prints per kernel: duration, DRAM
bytes read/written per launch (= `traffic` in bench.py's roofline object), DRAM/SM throughput %, issue slot
utilisation, pipe utilisation, occupancy, registers, and the warp-stall breakdown (cycles per issued instruction).
"""
import csv
import subprocess
import sys

KEYS = [
    ("duration", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("regs/thread", "launch__registers_per_thread"),
    ("smem/block dyn", "launch__shared_mem_per_block_dynamic"), ("smem/block static", "launch__shared_mem_per_block_static"),
    ("waves/SM", "launch__waves_per_multiprocessor"), ("achieved occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("dram read", "dram__bytes_read.sum"), ("dram write", "dram__bytes_write.sum"),
    ("dram throughput % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue slots busy % (active)", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warp cycles per issued inst", "smsp__average_warp_latency_per_inst_issued.ratio"),
    ("inst executed (warp)", "smsp__inst_executed.sum"),
    ("pipe fma %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("pipe fmaheavy %", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active"),
    ("pipe alu %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("pipe fp64 %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("pipe lsu %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("pipe tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("smem bank conflicts (ld)", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
    ("smem bank conflicts (st)", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"),
]
STALL = "smsp__average_warps_issue_stalled_"


def summarize(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    for r in data:
        print("=" * 100)
        print("report :", path)
        print("kernel :", r[ki][:160])
        for label, k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-32s %s %s" % (label, r[i], units[i]))
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i].replace(",", "")), h[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        print("  stall reasons (warp-cycles per issued instruction, top 8):")
        for v, n in sorted(stalls, reverse=True)[:8]:
            print("      %-28s %.3f" % (n, v))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        summarize(p)
