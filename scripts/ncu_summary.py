"""Summarise ncu artefacts into the small text files kept under profiles/.

  python scripts/ncu_summary.py launches <launches.csv>      # per-kernel time shares
  python scripts/ncu_summary.py full <report.ncu-rep>        # key metrics of each captured launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__inst_executed_op_ldgsts.sum", "smsp__inst_executed_op_tma_ld.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1)
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# total {total / 1e6:.1f} ms over {sum(cnt.values())} launches; per-launch times are cold-cache and "
          "serialised under ncu: compare SHARES")
    for n, v in tot.most_common(28):
        print(f"{v / 1e6:9.3f} ms {100 * v / total:5.1f}% {cnt[n]:5d}  {n}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("----", r[hdr.index("Kernel Name")][:100])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:78s} {r[hdr.index(k)]} {units[hdr.index(k)]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
