"""Turn ncu artefacts brought back from the GPU box (gpurun_out/) into the small markdown summaries kept under profiles/.

    python profiles/summarize_ncu.py rep  gpurun_out/X.ncu-rep  > profiles/r1_X.md     # --set full capture(s)
    python profiles/summarize_ncu.py list gpurun_out/launches.csv > profiles/r1_launches.md  # gpu__time_duration list
    python profiles/summarize_ncu.py table gpurun_out/X_raw.csv [label ...] > profiles/r2_X.md  # CSV exported on the box
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"# ncu --set full summary of `{path}`\n")
    print("(read with `ncu -i <rep> --page raw --csv`; values are per launch, cold-cache, under the profiler — "
          "shares and ratios are meaningful, absolute times are not bench numbers)\n")
    for r in rows[2:]:
        print(f"## `{r[ki][:110]}`\n")
        print("| metric | value |\n|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {r[i]} {units[i]} |")
        print()


TABLE_KEYS = [
    ("time us", "gpu__time_duration.sum"), ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("L2->SM", "l1tex__m_xbar2l1tex_read_bytes.sum"),
    ("DRAM rd", "dram__bytes_read.sum"), ("DRAM wr", "dram__bytes_write.sum"),
    ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("grid", "launch__grid_size"),
    ("regs", "launch__registers_per_thread"), ("waves", "launch__waves_per_multiprocessor"),
]


def table(path, labels=None):
    """Raw-metrics CSV exported ON THE GPU BOX (`ncu -i X.ncu-rep --page raw --csv`) -> one markdown row per launch."""
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"# ncu --set full, one row per launch, from `{path}`\n")
    print("(per launch, cold-cache, serialised under the profiler: ratios and shares are meaningful, absolute times are "
          "not bench numbers)\n")
    print("| # | kernel | " + " | ".join(k for k, _ in TABLE_KEYS) + " |")
    print("|---|---|" + "---|" * len(TABLE_KEYS))
    for n, r in enumerate(rows[2:]):
        name = r[ki].split("(")[0].replace("void ", "").replace("ddpm::", "")
        if labels and n < len(labels):
            name += f" -- {labels[n]}"
        cells = []
        for _, k in TABLE_KEYS:
            if k in hdr:
                i = hdr.index(k)
                v = r[i].replace(",", "")
                try:
                    cells.append(f"{float(v):.4g} {units[i]}".strip())
                except ValueError:
                    cells.append(v)
            else:
                cells.append("-")
        print(f"| {n} | `{name}` | " + " | ".join(cells) + " |")


def lst(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ui], 1e-3) * v
        name = r[ki].split("(")[0].replace("void ", "").replace("ddpm::", "")
        tot[name] += v
        cnt[name] += 1
    s = sum(tot.values())
    print(f"# launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`) of `{path}`\n")
    print(f"{sum(cnt.values())} launches, {s / 1e3:.2f} ms summed device time (serialised, cold caches: compare SHARES)\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"| `{k}` | {cnt[k]} | {v:.0f} | {100 * v / s:.1f} % |")


if __name__ == "__main__":
    if sys.argv[1] == "table":
        table(sys.argv[2], sys.argv[3:])
    else:
        (rep if sys.argv[1] == "rep" else lst)(sys.argv[2])
