"""Turn ncu output into the small summaries kept under profiles/.

  python tools/ncu_summary.py launches <launches.csv> "<command line>"      -> per-kernel shares (text)
  python tools/ncu_summary.py full <raw.csv>                                -> selected metrics (csv)

<launches.csv> is the --csv log of `ncu --metrics gpu__time_duration.sum --clock-control none`,
<raw.csv> is `ncu -i report.ncu-rep --page raw --csv`."""
import collections
import csv
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(path, command):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].split("(")[0]
        tot[name] += float(r[vi].replace(",", "")) / 1e3
        cnt[name] += 1
    total = sum(tot.values())
    print("ncu launch list of `%s`" % command)
    for name, us in tot.most_common():
        print("%-62s n=%4d total=%9.1f us avg=%8.1f us share=%5.1f%%" % (name[:62], cnt[name], us, us / cnt[name],
                                                                           100 * us / total))


def full(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = csv.writer(sys.stdout)
    out.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(rows) - 2)])
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    for name in ["Kernel Name"] + KEEP + stalls:
        if name in hdr:
            i = hdr.index(name)
            out.writerow([name, units[i]] + [r[i] for r in rows[2:]])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2])
