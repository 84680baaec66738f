"""Turn gpurun_out/*.ncu-rep / launch-list CSVs into the markdown summaries committed under profiles/."""
import collections
import csv
import re
import subprocess
import sys


def launch_list(csv_path, out_path, title, note):
    lines = [l for l in open(csv_path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"rgrg::", "", name)[:100]
        agg[name][0] += 1
        agg[name][1] += v
        agg[name][2] = max(agg[name][2], v)
        n += 1
    tot = sum(a[1] for a in agg.values())
    with open(out_path, "w") as f:
        f.write("# %s\n\n%s\n\n%d launches, %.1f ms of kernel time.  Per-launch times are cold-cache and serialised by ncu: compare SHARES.\n\n" % (title, note, n, tot / 1e3))
        f.write("| kernel | launches | total us | avg us | max us | share |\n|---|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.2f | %.1f | %.3f |\n" % (k, a[0], a[1], a[1] / a[0], a[2], a[1] / tot))


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum"]


def full_report(rep_path, out_path, title, note):
    out = subprocess.run(["ncu", "-i", rep_path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out_path, "w") as f:
        f.write("# %s\n\n%s\n\n" % (title, note))
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write("## `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % re.sub(r"rgrg::", "", name)[:160])
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write("| %s | %s | %s |\n" % (w, r[i], units[i]))
            f.write("\n")


WANT2 = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
         ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
         ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
         ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor % (elapsed)"),
         ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor % (active)"),
         ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
         ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM read")]


def wide_table(rep_path, out_path, title, note, skip=()):
    """one row per captured launch, one column per metric"""
    if rep_path.endswith(".csv"):  # already exported on the GPU box: `ncu -i rep --page raw --csv`
        out = open(rep_path).read()
    else:
        out = subprocess.run(["ncu", "-i", rep_path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out_path, "w") as f:
        f.write("# %s\n\n%s\n\n" % (title, note))
        f.write("| kernel | " + " | ".join(w[1] for w in WANT2) + " |\n|---|" + "---:|" * len(WANT2) + "\n")
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", re.sub(r"rgrg::", "", r[hdr.index("Kernel Name")])).replace("void ", "")[:120]
            if any(s_ in name for s_ in skip):
                continue
            cells = []
            for w, _ in WANT2:
                if w not in hdr:
                    cells.append("-")
                    continue
                i = hdr.index(w)
                v, u = r[i], units[i]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                u = {"register/thread": "", "": "", "%": " %"}.get(u, " " + u)
                cells.append(v + u)
            f.write("| `%s` | " % name + " | ".join(cells) + " |\n")


if __name__ == "__main__":
    kind = sys.argv[1]
    if kind == "wide":
        wide_table(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5], tuple(sys.argv[6:]))
    else:
        (launch_list if kind == "list" else full_report)(*sys.argv[2:6])
