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


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
