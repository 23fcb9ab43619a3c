"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.
  python tools/ncu_summary.py rep <file.ncu-rep> <out.txt>      key metrics of every captured launch
  python tools/ncu_summary.py launches <launches.csv> <out.tsv> per-kernel totals and shares of a launch list"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x"]


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none, from {path.split('/')[-1]} (one block per captured launch)\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"\n{name[:160]}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"  {k:75s} {r[i]:>16s} {units[i]}\n")


def launches(path, out):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    tot = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        if r[ui] == "ns":
            v /= 1e3
        elif r[ui] == "ms":
            v *= 1e3
        name = r[ki].split("(")[0][:110]
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += v
    total = sum(t[1] for t in tot.values())
    with open(out, "w") as f:
        f.write(f"# per-kernel totals of {path.split('/')[-1]} (gpu__time_duration.sum, ncu replay: cold-cache, serialised — compare SHARES)\n")
        f.write("kernel\tlaunches\ttotal_us\tshare\n")
        for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name}\t{n}\t{us:.1f}\t{us / total:.4f}\n")
        f.write(f"TOTAL\t{sum(t[0] for t in tot.values())}\t{total:.1f}\t1.0\n")


def traffic(path, out):
    """per-kernel duration + DRAM bytes of ONE decode: the launches between the last two latent_to_nhwc launches of a
    `--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` CSV (graphs off)."""
    import json
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    idi, ki, mi, ui, vi = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    per = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        d = per.setdefault(int(r[idi]), {"name": r[ki]})
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1.0)
        d[r[mi]] = v * scale
    ids = sorted(per)
    starts = [i for i in ids if per[i]["name"].startswith("latent_to_nhwc")]
    lo, hi = starts[-2], starts[-1]
    summ = OrderedDict()
    for i in ids:
        if not (lo <= i < hi):
            continue
        d = per[i]
        name = d["name"].split("<")[0].split("(")[0].replace("void ", "")
        t = summ.setdefault(name, {"launches": 0, "ms": 0.0, "dram_read_gb": 0.0, "dram_write_gb": 0.0})
        t["launches"] += 1
        t["ms"] += d.get("gpu__time_duration.sum", 0.0)
        t["dram_read_gb"] += d.get("dram__bytes_read.sum", 0.0)
        t["dram_write_gb"] += d.get("dram__bytes_write.sum", 0.0)
    json.dump(summ, open(out, "w"), indent=1)


if __name__ == "__main__":
    {"rep": rep, "launches": launches, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
