#!/usr/bin/env python
"""Turn the artefacts a GPU round brought back (gpurun_out/) into the tracked
summaries under profiles/:  launches.csv -> per-kernel time shares of one eager
step;  *.ncu-rep (ncu --set full) -> the key raw metrics per captured launch."""
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_requests_srcunit_tex.sum",
        "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_sectors.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "sm__cycles_active.avg"]


def launches(tag, fname="launches.csv", out="launches_one_step", what="DeepFM (c2)", pick=(-3, -2),
             cmd="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline"):
    path = os.path.join(OUT, fname)
    if not os.path.exists(path):
        return
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
    names = [r["Kernel Name"] for r in rows]
    starts = [i for i, n in enumerate(names) if "gather_fm_fwd" in n]
    if len(starts) < 3:
        return
    # one train step = the longest run of launches between two consecutive gather launches (the roofline
    # section of bench.py launches the gather kernel back to back: those windows have length 1)
    wins = [(x, y) for x, y in zip(starts[:-1], starts[1:]) if y - x > 5]
    if not wins:
        return
    a, b = wins[-1]                   # the last full step (the first one also zero-fills the Adam slots)
    step = [r for r in rows[a:b] if "FillFunctor<unsigned char>" not in r["Kernel Name"]]     # bench.py's L2 flush
    tot = sum(float(r["Metric Value"]) for r in step) / 1e3
    with open(os.path.join(ROOT, "profiles", f"{tag}_{out}.md"), "w") as fh:
        fh.write(f"# {tag}: every kernel of ONE eager {what} train step, `ncu --metrics gpu__time_duration.sum "
                 f"--clock-control none`\n\nCold-cache, serialised per-launch times: compare SHARES, not absolutes. "
                 f"Sum = {tot:.1f} us over {len(step)} launches "
                 f"(command: `{cmd}`).\n\n"
                 f"| # | us | share | kernel |\n|---|---|---|---|\n")
        for i, r in enumerate(step):
            v = float(r["Metric Value"]) / 1e3
            kname = r["Kernel Name"].split("(")[0][:110]
            fh.write(f"| {i} | {v:.1f} | {100 * v / tot:.1f}% | `{kname}` |\n")


def rep(name, tag):
    path = os.path.join(OUT, name + ".ncu-rep")
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    out_name = name if name.startswith(tag) else f"{tag}_{name}"
    with open(os.path.join(ROOT, "profiles", f"{out_name}.md"), "w") as fh:
        fh.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` -> {name}.ncu-rep\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            fh.write(f"## {d.get('Kernel Name', '?')[:120]}  (launch id {d.get('ID')})\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in hdr:
                if any(k == kk or k.startswith(kk) for kk in KEYS):
                    fh.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
            fh.write("\n")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    launches(tag)
    launches(tag, "launches_c3.csv", "launches_c3_step", "DCN-matrix bf16 (c3)", (-2, -1),
             "python bench.py --config c3 --steps 1 --warmup 3 --no-graph")
    names = sys.argv[2:] or ["prof_gather_fwd", "prof_tcgemm", "prof_cross", "prof_fused_short"]
    for n in names:
        rep(n, tag)
    print(os.listdir(os.path.join(ROOT, "profiles")))
