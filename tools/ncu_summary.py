"""Turn ncu outputs into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r1b.csv > profiles/r1_launches.md
    python tools/ncu_summary.py full gpurun_out/prof_r1c.ncu-rep > profiles/r1_hoptraj_hist_full.md
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        per.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in per.values())
    print("| kernel | launches | total us | mean us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        print("| %s | %d | %.1f | %.1f | %.1f%% |" % (k[:60], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))


METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    # the L2 side of the scattered reductions (binning): tag pipeline, atomic units, RED sectors by lookup result
    "lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed", "lts__t_tag_requests.max.pct_of_peak_sustained_elapsed",
    "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_red.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum",
    "lts__t_sectors.sum", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed_op_global_red.sum",
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("### %s" % r[idx["Kernel Name"]].split("(")[0].replace("void ", ""))
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for m in METRICS:
            if m in idx:
                print("| %s | %s | %s |" % (m, r[idx[m]], units[idx[m]]))
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
