#!/usr/bin/env python3
"""Summarise `ncu --set full` captures of the persistent kernel into profiles/.

  python tools/ncu_extract.py <K> <I> <individuals> <svi-iterations-per-launch> <capture.ncu-rep> [tag]

Reads the report with `ncu -i ... --page raw --csv` (no GPU needed) and merges one entry, keyed
"K,I,individuals", into profiles/r2_ncu_metrics.json -- the file bench.py takes `roofline.traffic` and
`fp64_pipe_pct` from.  Also writes profiles/r2_ncu_<tag>.json with the full list of metrics quoted in
profiles/r2_summary.md and, when the report holds source-level counters, the stall-reason shares."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
KEEP = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "lts__t_bytes.sum", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def main():
    k, ipt, n, iters, rep = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    tag = sys.argv[6] if len(sys.argv) > 6 else "k%d_i%d_n%d" % (k, ipt, n)
    m = raw_page(rep)

    def num(name):
        v, u = m[name]
        return float(v.replace(",", "")) * UNIT.get(u, 1.0)
    dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    full = {h: " ".join(x for x in m[h] if x) for h in KEEP if h in m}
    full["svi_iterations_per_launch"] = iters
    full["source"] = os.path.basename(rep)
    json.dump(full, open(os.path.join(ROOT, "profiles", "r2_ncu_%s.json" % tag), "w"), indent=1)
    path = os.path.join(ROOT, "profiles", "r2_ncu_metrics.json")
    allm = json.load(open(path)) if os.path.exists(path) else {}
    dur_ms = num("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(m["gpu__time_duration.sum"][1], 1.0)
    allm["%d,%d,%d" % (k, ipt, n)] = {
        "kernel": m["Kernel Name"][0], "capture": os.path.basename(rep), "svi_iterations_per_launch": iters,
        "duration_ms_under_ncu": dur_ms,
        "dram_bytes_per_svi_iteration": dram / iters,
        "algorithmic_bytes_per_svi_iteration": (32.0 * k + 0.25) * n,
        "fp64_pipe_pct": float(m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0]),
        "registers_per_thread": int(float(m["launch__registers_per_thread"][0])),
        "dram_throughput_pct": float(m["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0]),
    }
    json.dump(allm, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(allm["%d,%d,%d" % (k, ipt, n)], indent=1))


if __name__ == "__main__":
    main()
