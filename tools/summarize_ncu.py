#!/usr/bin/env python
"""Summaries for profiles/: (1) a per-kernel table from an `ncu --metrics gpu__time_duration.sum --csv` launch
list, (2) key metrics + stall breakdown from an `ncu --set full` report (read with `ncu -i ... --page raw/source`).

  python tools/summarize_ncu.py launches <launches.csv>
  python tools/summarize_ncu.py report <file.ncu-rep>
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_active.avg",
        "sm__cycles_elapsed.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size"]


def launches(path, n_sms=148):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    agg, smt = collections.defaultdict(list), collections.defaultdict(float)
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        name = r[ki].split("(")[0]
        agg[name].append(v)
        grid = 1
        for g in r[gi].strip("()").split(","):
            grid *= int(g)
        smt[name] += v * min(1.0, grid / n_sms)   # device time weighted by the fraction of SMs the grid can occupy
    tot, tot_sm = sum(sum(v) for v in agg.values()), sum(smt.values())
    print("# per-kernel device time from %s (ncu replays each launch alone and cold: compare SHARES)" % path)
    print("# sm_share weights every launch by min(1, CTAs / %d SMs): a 1-CTA kernel occupies 1/%d of the GPU" % (n_sms, n_sms))
    print("%-44s %7s %11s %7s %9s %9s %9s %9s" % ("kernel", "n", "total_ms", "share", "sm_share", "avg_us", "min_us", "max_us"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-44s %7d %11.3f %6.1f%% %8.1f%% %9.2f %9.2f %9.2f" % (k[:44], len(v), sum(v) / 1000, 100 * sum(v) / tot,
                                                                    100 * smt[k] / tot_sm, sum(v) / len(v), min(v), max(v)))
    print("%-44s %7d %11.3f" % ("TOTAL", sum(len(v) for v in agg.values()), tot / 1000))


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("# %s : %s" % (path, name.split("(")[0]))
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                print("%-82s %-10s %s" % (h, u, v))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) < 3:
        return
    hdr, data = rows[1], rows[2:]
    try:
        s0 = hdr.index("stall_barrier")
    except ValueError:
        return
    names = hdr[s0:s0 + 17]
    tot = sum(int(r[2]) for r in data) or 1
    byop, bystall = collections.Counter(), collections.Counter()
    for r in data:
        t = r[1].strip().split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        byop[op] += int(r[2])
        for n, v in zip(names, r[s0:s0 + 17]):
            bystall[n] += int(v)
    print("warp-state samples: %d" % tot)
    print("by opcode  : " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in byop.most_common(10)))
    print("by reason  : " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in bystall.most_common(9)))


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
