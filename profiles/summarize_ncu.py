"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept in profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_launches.md
    python profiles/summarize_ncu.py full gpurun_out/prof.ncu-rep      > profiles/rNN_full.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]

EXTRA_PATTERNS = ["pipe_tensor", "pipe_tc", "tmem", "utcmma", "utchmma", "l1tex__data_pipe_lsu_wavefronts_mem_shared",
                  "l1tex__data_bank", "smsp__inst_executed_op_shared", "sm__mio", "lts__t_sectors_op_red",
                  "lts__t_sectors_op_atom", "l1tex__m_xbar2l1tex", "smsp__warp_issue_stalled_barrier",
                  "sm__pipe_shared", "stalled_membar", "stalled_sleeping", "stalled_no_instruction", "stalled_branch", "stalled_drain", "stalled_imc", "stalled_tex", "stalled_selected"]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("b2s::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ns | share | avg us |\n|---|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {c} | {t:.0f} | {100 * t / tot:.1f}% | {t / c / 1000:.1f} |")
    print(f"\ntotal {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare shares)")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("### " + r[idx["Kernel Name"]].split("(")[0])
        for k in KEYS:
            if k in idx:
                print(f"- {k} = {r[idx[k]]} {units[idx[k]]}")
        # tensor pipe, TMEM and shared-memory pipe: whatever the capture holds under these names
        for h in hdr:
            v = r[idx[h]] if h in idx else ""
            if any(t in h for t in (".min", ".max", "peak_sustained", "per_second", "per_cycle")) and "pct_of_peak" not in h:
                continue
            if any(t in h for t in (".min.", ".max.")):
                continue
            try:
                if float(v.replace(",", "")) == 0.0:
                    continue
            except ValueError:
                continue
            if h not in KEYS and any(p in h for p in EXTRA_PATTERNS):
                print(f"- {h} = {r[idx[h]]} {units[idx[h]]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
