#!/usr/bin/env python3
"""profiles/roofline_inputs.json from `ncu --set full` raw CSVs of the build being benchmarked (tools/profile_pass.sh).

usage: roofline_inputs.py <trace_raw.csv> <shade_raw.csv> <rays of the captured k_trace_closest launch> [source note]

The oracle's per-ray node / primitive counts (the algorithmic-bytes definition of SURVEY.md s8d) are kept from the existing file:
they depend on the scene and the reference algorithm only.  The captured launches are the FIRST k_trace_closest launch of the capture
(iteration 0 of a 16-spp S3 step: every camera ray of the step, so the ray count is known) and the longest k_shade launch of the same step.
"""
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles" / "roofline_inputs.json"


def longest(path, needle, first=False):
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    ki, di = h.index("Kernel Name"), h.index("gpu__time_duration.sum")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[di]]
    best = None
    for r in rows[2:]:
        if needle in r[ki] and (best is None or (not first and float(r[di]) > float(best[di]))):
            best = r
    if best is None:
        raise SystemExit(f"no {needle} launch in {path}")

    def val(name, mult=None):
        i = h.index(name)
        v = float(best[i])
        if mult:
            v *= mult[units[i]]
        return v

    byte = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return {"kernel": best[ki], "ms": float(best[di]) * scale,
            "dram_bytes": int(val("dram__bytes_read.sum", byte) + val("dram__bytes_write.sum", byte)),
            "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "lanes_per_inst": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "occupancy_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "l1_hit_pct": val("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": val("lts__t_sector_hit_rate.pct"),
            "dram_pct_of_ncu_peak": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "registers": int(val("launch__registers_per_thread")), "warp_instructions": int(val("smsp__inst_executed.sum"))}


def main():
    trace_csv, shade_csv, rays = sys.argv[1], sys.argv[2], int(sys.argv[3])
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    old = json.loads(OUT.read_text()) if OUT.exists() else {}
    tc = longest(trace_csv, "k_trace_closest", first=True)  # the FIRST captured launch: iteration 0, whose ray count is known (every camera ray of the step)
    tc["rays"] = rays
    out = {"comment": "inputs of bench.py's roofline line that do not change between runs of one build: oracle counters on S3 (reference binary BVH, maxnodeprims 4, "
                      "SAH) and ncu --set full --clock-control none counters of the longest k_trace_closest / k_shade launch of a 16-spp 1080p step",
           "source": note,
           "nodes_per_closest_ray": old.get("nodes_per_closest_ray"), "prims_per_closest_ray": old.get("prims_per_closest_ray"),
           "trace_closest": tc, "shade": longest(shade_csv, "k_shade")}
    OUT.write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
