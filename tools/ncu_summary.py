"""Summarise ncu reports brought back in gpurun_out/ into small tracked JSON files under profiles/ (the .ncu-rep files themselves are scratch).
usage: python tools/ncu_summary.py gpurun_out/r2_g6_fold.ncu-rep [more.ncu-rep ...]   ->  profiles/<stem>_ncu_summary.json
Reads `ncu -i <rep> --page raw --csv` (first captured launch) and `--page source --csv` (stall reasons aggregated over the kernel)."""
import csv, io, json, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:                                                     # the capture's kernel filter matched no launch: nothing to summarise
        raise SystemExit(f"{rep}: no kernel in this report (check the -k filter of the capture: it matches the base name, not the template arguments)")
    hdr = rows[0]; units = rows[1]; vals = rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") or h == "Kernel Name":
            try:
                d[h] = {"value": float(v.replace(",", "")), "unit": u} if h != "Kernel Name" else v
            except ValueError:
                d[h] = v
    return d


def stalls(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    try:
        h = next(r for r in rows if "# Samples" in r)
    except StopIteration:
        return None
    ix = {n: i for i, n in enumerate(h)}
    data = rows[rows.index(h) + 1:]
    tot = sum(int(r[ix["# Samples"]]) for r in data if len(r) > ix["# Samples"])
    reasons = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    agg = {n: sum(int(r[ix[n]]) for r in data if len(r) > ix[n]) for n in reasons}
    ops = {}
    for r in data:
        if len(r) <= ix["Instructions Executed"]:
            continue
        t = r[ix["Source"]].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[ix["Instructions Executed"]])
    return {"samples": tot, "by_reason_pct": {k: round(100.0 * v / max(tot, 1), 1) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]},
            "warp_instructions_executed_by_opcode": dict(sorted(ops.items(), key=lambda x: -x[1])[:12])}


OUT_DIR = ROOT / "profiles"
args = sys.argv[1:]
if args and args[0] == "--out":                                          # on the GPU box: summaries into gpurun_out/ (the .ncu-rep files are too big to bring back)
    OUT_DIR = Path(args[1]); args = args[2:]
for arg in args:
    rep = Path(arg)
    d = raw(rep)
    t = d.get("gpu__time_duration.sum", {}).get("value")
    rd, wr = d.get("dram__bytes_read.sum", {}), d.get("dram__bytes_write.sum", {})
    def to_bytes(x):
        if not isinstance(x, dict):
            return None
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(x.get("unit"), 1)
        return x["value"] * mult
    summary = {"report": rep.name, "kernel": d.get("Kernel Name"), "duration_us": t if d.get("gpu__time_duration.sum", {}).get("unit") in ("us", "usecond") else (t / 1e3 if t and d["gpu__time_duration.sum"].get("unit") in ("ns", "nsecond") else t),
               "duration_unit_in_report": d.get("gpu__time_duration.sum", {}).get("unit"),
               "dram_bytes_per_launch": (to_bytes(rd) or 0) + (to_bytes(wr) or 0) if rd and wr else None, "metrics": {k: v for k, v in d.items() if k != "Kernel Name"}, "stall_sampling": stalls(rep),
               "how": "ncu --set full --clock-control none --import-source on, one launch after warm-up (cold caches, serialised: compare shares and ratios, not absolutes)"}
    import re
    out = OUT_DIR / (re.sub(r"^r2_g\d+_", "r02_", rep.stem) + "_ncu_summary.json")
    out.write_text(json.dumps(summary, indent=1))
    print(out.name, summary["kernel"][:60] if summary["kernel"] else None, summary["duration_us"], summary["duration_unit_in_report"])
