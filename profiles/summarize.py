"""Turn ncu outputs (gpurun_out/) into the small text summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv  > profiles/rN_launches.md
  python profiles/summarize.py report   gpurun_out/prof.ncu-rep  > profiles/rN_kernel.md
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__cycles_active.avg", "launch__grid_size", "launch__block_size"]


def launches(path):
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    idx = [i for i, x in enumerate(rows) if "to_channels_last" in x["Kernel Name"]]
    fw = rows[idx[0]:] if idx else rows
    agg = collections.defaultdict(lambda: [0, 0.0])
    for x in fw:
        name = re.sub(r"\(.*", "", x["Kernel Name"]).replace("mudg::<unnamed>::", "").replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += float(x["Metric Value"]) / 1e3
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list: one UNet forward ({len(fw)} launches, sum of gpu__time_duration = {tot / 1e3:.2f} ms)\n")
    print("Per-launch times are cold-cache and serialised (ncu); compare SHARES.\n")
    print("| share | total ms | launches | avg us | kernel |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {v[1] / tot * 100:.2f}% | {v[1] / 1e3:.2f} | {v[0]} | {v[1] / v[0]:.1f} | `{k[:70]}` |")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of `{path}`\n")
    for r in rows[2:]:
        d = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
        print(f"## {d['Kernel Name'][0][:80]}  grid {d.get('Grid Size', ('', ''))[0]} block {d.get('Block Size', ('', ''))[0]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d:
                print(f"| {k} | {d[k][0]} | {d[k][1]} |")
        print()


def table(paths, out_json=None):
    """One row per distinct kernel (first captured launch of each name + grid) of `ncu --set full` reports: duration, DRAM
    bytes, achieved DRAM GB/s against the measured copy peak, tensor-pipe %, issue %, registers.  Also written as JSON
    (bench.py attaches it to roofline.ncu)."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        hbm = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm = 6650.0
    seen, rows_out = set(), []
    for path in paths:
        if path.endswith(".csv"):        # already exported on the GPU box: ncu -i x.ncu-rep --page raw --csv > x.csv
            out = open(path).read()
        else:
            out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(l for l in out.splitlines() if not l.startswith("==")))
        if len(rows) < 3:
            continue
        hdr = rows[0]
        for r in rows[2:]:
            d = {h: r[i] for i, h in enumerate(hdr)}
            name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("mudg::<unnamed>::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("void ", "")
            key = (name, d.get("Grid Size", ""))
            if key in seen:
                continue
            seen.add(key)
            f = lambda k: float(d[k].replace(",", "")) if d.get(k, "") not in ("", "n/a") else float("nan")
            tunit = rows[1][hdr.index("gpu__time_duration.sum")]
            dur_us = f("gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(tunit, 1e-3)
            rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
            # ncu reports dram bytes in the unit of the second header row; normalise through the throughput metric when needed
            rows_out.append(dict(kernel=name, grid=d.get("Grid Size", ""), block=d.get("Block Size", ""), duration_us=dur_us,
                                 dram_read=rd, dram_write=wr, dram_unit=rows[1][hdr.index("dram__bytes_read.sum")] if "dram__bytes_read.sum" in hdr else "",
                                 dram_pct=f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                                 tensor_pct=f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                                 xu_pct=f("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                                 issue_pct=f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                 regs=f("launch__registers_per_thread"), source=os.path.basename(path)))
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    print("# ncu --set full, one row per kernel (first captured launch of each name / grid)\n")
    print(f"Achieved DRAM GB/s = (dram__bytes_read + dram__bytes_write) / gpu__time_duration; peak = {hbm:.0f} GB/s (MEASURED_PEAKS.json). "
          "Times under ncu are cold-cache, serialised and at unlocked clocks: they explain a kernel, they are not bench values.\n")
    print("| kernel | grid x block | us | DRAM MB (r + w) | GB/s | of peak | tensor pipe % | XU % | issue % | regs |\n|---|---|---|---|---|---|---|---|---|---|")
    for r in rows_out:
        m = mult.get(r["dram_unit"], 1.0)
        mb = (r["dram_read"] + r["dram_write"]) * m / 1e6
        gbs = mb / 1e3 / (r["duration_us"] / 1e6) if r["duration_us"] > 0 else float("nan")
        r["dram_bytes_per_launch"] = mb * 1e6
        r["dram_gbs"] = gbs
        r["dram_frac_of_measured_peak"] = gbs / hbm
        print(f"| `{r['kernel'][:60]}` | {r['grid']} x {r['block']} | {r['duration_us']:.1f} | {mb:.1f} | {gbs:.0f} | {gbs / hbm:.2f} | "
              f"{r['tensor_pct']:.1f} | {r['xu_pct']:.1f} | {r['issue_pct']:.1f} | {r['regs']:.0f} |")
    if out_json:
        with open(out_json, "w") as fjson:
            json.dump(rows_out, fjson, indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "table":
        args = sys.argv[2:]
        oj = None
        if "--json" in args:
            oj = args[args.index("--json") + 1]
            args = [a for a in args if a not in ("--json", oj)]
        table(args, oj)
    else:
        {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
