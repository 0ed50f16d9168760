"""Summarise ncu outputs brought back in gpurun_out/ into committed markdown under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
  python profiles/summarize.py report   gpurun_out/gather_fast_r1.ncu-rep profiles/r1_gather_fast.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__cycles_active.avg",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"krs::\(anonymous namespace\)::", "", name)
    return name[:90]


def launches(src, dst):
    rows = []
    with open(src, newline="") as f:
        text = f.read()
    start = text.find('"ID"')
    rd = csv.DictReader(io.StringIO(text[start:]))
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            mult = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1e-3)
            rows.append((short(r["Kernel Name"]), v * mult))
    agg = collections.OrderedDict()
    for k, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}) — per-kernel device time (cold-cache, serialised: compare SHARES)\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {us:.1f} | {us / n:.1f} | {100 * us / total:.1f}% |\n")
        f.write(f"\ntotal {total:.1f} us over {len(rows)} launches\n")
    print(open(dst).read())


def report(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr = rd[0]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src}\n\n")
        for row in rd[2:]:
            d = dict(zip(hdr, row))
            f.write(f"## {short(d.get('Kernel Name', '?'))}  (id {d.get('ID')})\n\n| metric | value |\n|---|---|\n")
            for k in hdr:
                if any(k.startswith(x) for x in KEYS):
                    f.write(f"| {k} | {d[k]} |\n")
            f.write("\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
