#!/usr/bin/env python3
"""Condense ncu output brought back in gpurun_out/ into the small tracked files under profiles/.

  launches  <launches.csv> <out.md>       per-kernel totals and shares of a `--metrics gpu__time_duration.sum` pass
  kernels   <raw.csv> <out.md> [regex]    key counters of every captured launch of a `--set full` report
                                          (raw.csv = `ncu -i X.ncu-rep --page raw --csv`)
  source    <src.csv[.gz]> <out.md> [N]   top N source lines of one launch by warp-stall samples, with their share of the warp
                                          instructions and the average number of active lanes
                                          (src.csv = `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`)
"""
import collections
import csv
import re
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instr (of 32)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe ALU %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe FMA %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe LSU %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local loads (stack)"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local stores (stack)"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp latency / instr issued (cycles)"),
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    tot = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1e-3)
        t = tot.setdefault(k, [0, 0.0, 0.0])
        t[0] += 1; t[1] += v; t[2] = max(t[2], v)
    s = sum(v[1] for v in tot.values())
    with open(dst, "w") as f:
        f.write(f"ncu launch list `{src}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n\n")
        f.write("| kernel | launches | total us | share | longest us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1][1]):
            f.write(f"| {k} | {v[0]} | {v[1]:.1f} | {100 * v[1] / s:.1f}% | {v[2]:.1f} |\n")
        f.write(f"\ntotal {s / 1e3:.3f} ms over {sum(v[0] for v in tot.values())} launches\n")


def kernels(src, dst, pattern=None):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"ncu --set full capture `{src}` (one block per captured launch; per launch values)\n")
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            if pattern and not re.search(pattern, name):
                continue
            f.write(f"\n## {name} (launch id {r[col['ID']]})\n\n| counter | value |\n|---|---|\n")
            for key, label in KEYS:
                if key in col:
                    f.write(f"| {label} (`{key}`) | {r[col[key]]} {units[col[key]]} |\n")
            st = []
            for h, i in col.items():
                m = STALL.match(h)
                if m:
                    try:
                        st.append((float(r[i]), m.group(1)))
                    except ValueError:
                        pass
            st.sort(reverse=True)
            f.write("| top stall reasons (warps stalled per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in st[:6]) + " |\n")


def source(src, dst, top="40"):
    import gzip
    import io
    text = gzip.open(src, "rt").read() if src.endswith(".gz") else open(src).read()
    hdr, fname, kernel, items, tot, tot_inst, tot_thr = None, "", "", [], 0, 0, 0
    for r in csv.reader(io.StringIO(text)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            kernel = r[1]
        elif r[0] == "Line No":
            hdr = r
            si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        elif hdr is not None and r[0] != "":
            try:
                n, ie, te = int(r[si]), int(r[ii]), int(r[ti])
            except (ValueError, IndexError):
                continue
            tot += n; tot_inst += ie; tot_thr += te
            items.append((n, ie, te, fname, r[0], r[1].strip()[:110].replace("|", "\\|")))
    items.sort(reverse=True)
    with open(dst, "w") as f:
        f.write(f"ncu source page `{src}`: `{kernel}`\n\n{tot} stall samples, {tot_inst} warp instructions, {tot_thr / max(tot_inst, 1):.1f} active lanes per warp "
                f"instruction on average (inlined callee lines are counted under every inlining frame, so shares can add up to more than 100 %)\n\n")
        f.write("| stall samples | warp instr | lanes | line | source |\n|---:|---:|---:|---|---|\n")
        for n, ie, te, fn, ln, text_ in items[:int(top)]:
            f.write(f"| {100 * n / max(tot, 1):.1f}% | {100 * ie / max(tot_inst, 1):.1f}% | {te / max(ie, 1):.1f} | {fn}:{ln} | `{text_}` |\n")


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels, "source": source}[sys.argv[1]](*sys.argv[2:])
