"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_ncu.py launches <launches.csv> <out.md> [title]
    python scripts/summarize_ncu.py full <prof.ncu-rep> <out.md> [title]

`launches`: the `--metrics gpu__time_duration.sum` launch list of ONE training step (scripts/profile_step.py inside a
cudaProfilerStart/Stop range): per-kernel totals, shares of the step and the launches above 50 us.
`full`: key metrics of every kernel captured with `--set full` (DRAM bytes, DRAM %, tensor-pipe %, occupancy, registers).
"""
import collections
import csv
import re
import subprocess
import sys


def short_name(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("ctu::", "")
    return re.sub(r"\(.*", "", name)


def launches(path, out, title):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        rows.append((short_name(row["Kernel Name"]), v, row["Grid Size"], row["Block Size"]))
    total = sum(r[1] for r in rows)
    agg = collections.OrderedDict()
    for n, v, _, _ in rows:
        a = agg.setdefault(n, [0.0, 0])
        a[0] += v
        a[1] += 1
    with open(out, "w") as f:
        f.write("# %s\n\n" % title)
        f.write("Source: `%s` (ncu `--metrics gpu__time_duration.sum --clock-control none`, one training step; per-launch "
                "times are cold-cache and serialised -- compare SHARES).\n\n" % path)
        f.write("Launches: %d, sum of kernel time: %.1f us\n\n" % (len(rows), total))
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            if v / total < 0.001:
                continue
            f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (n, c, v, 100 * v / total))
        f.write("\n## Launches above 50 us, in launch order\n\n| # | us | kernel | grid | block |\n|---:|---:|---|---|---|\n")
        for i, (n, v, g, b) in enumerate(rows):
            if v > 50:
                f.write("| %d | %.1f | `%s` | %s | %s |\n" % (i, v, n, g, b))


WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
    ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "tc pipe %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem->TC %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
]


def full(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# %s\n\nSource: `%s` (ncu `--set full --clock-control none --import-source on`).\n\n" % (title, path))
        f.write("| # | kernel | grid | " + " | ".join(lbl for _, lbl in WANT) + " |\n")
        f.write("|---:|---|---|" + "---:|" * len(WANT) + "\n")
        for k, r in enumerate(rows[2:]):
            cells = []
            for m, _ in WANT:
                i = idx.get(m)
                if i is None:
                    cells.append("n/a")
                    continue
                v = r[i]
                try:
                    v = "%.1f" % float(v.replace(",", ""))
                except ValueError:
                    pass
                cells.append("%s %s" % (v, units[i].replace("byte", "B").replace("register/thread", "")))
            f.write("| %d | `%s` | %s | %s |\n" % (k, short_name(r[idx["Kernel Name"]]), r[idx["Grid Size"]], " | ".join(cells)))


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    ttl = sys.argv[4] if len(sys.argv) > 4 else src
    (launches if mode == "launches" else full)(src, dst, ttl)
