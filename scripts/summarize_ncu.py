"""Turn ncu outputs into the small text summaries committed under profiles/.
  python scripts/summarize_ncu.py launches <launches.csv> <out.txt> "<title>"
  python scripts/summarize_ncu.py full <report.ncu-rep> <out.txt> "<title>"
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path, out, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1]
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault((name, row["Grid Size"]), [0, 0.0, []])
        a[0] += 1; a[1] += v; a[2].append(v)
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(title + "\n")
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised: "
                "compare SHARES, not absolutes\n")
        f.write(f"{'kernel':46s} {'grid':>14s} {'n':>5s} {'total_us':>10s} {'med_us':>8s} {'share':>7s}\n")
        for (name, grid), (n, t, vs) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            vs = sorted(vs)
            f.write(f"{name[:46]:46s} {grid:>14s} {n:5d} {t / 1e3:10.1f} {vs[len(vs) // 2] / 1e3:8.2f} {t / tot:7.1%}\n")
        by_name = collections.Counter()
        for (name, _), (n, t, vs) in agg.items():
            by_name[name] += t
        f.write("\nshare by kernel function:\n")
        for name, t in by_name.most_common():
            f.write(f"  {name[:60]:60s} {t / tot:7.1%}\n")


WANT = ["Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
        "smsp__warp_issue_stalled_sleeping_per_warp_active.pct", "smsp__warp_issue_stalled_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
        "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_drain_per_warp_active.pct", "smsp__warp_issue_stalled_imc_miss_per_warp_active.pct"]


def opcode_mix(rep, f):
    """instruction mix per kernel from the source page (needs -lineinfo + --import-source on)"""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    cur, hdr, agg, samp = None, None, None, None

    def flush():
        if cur and agg:
            tot, ts = sum(agg.values()), max(1, sum(samp.values()))
            f.write(f"\n--- opcode mix (all captured launches): {cur[:100]}\n    total warp instructions {tot}\n")
            for k, v in agg.most_common(14):
                f.write(f"    {k:10s} {100 * v / tot:6.2f} % of instructions   {100 * samp[k] / ts:6.2f} % of stall samples\n")
    for r in csv.reader(raw.splitlines()):
        if r and r[0] == "Kernel Name":
            flush()
            cur, hdr, agg, samp = r[1], None, collections.Counter(), collections.Counter()
        elif r and r[0] == "Address":
            hdr = r
        elif hdr and len(r) == len(hdr):
            try:
                toks = r[hdr.index("Source")].split()
                op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
                agg[op] += int(r[hdr.index("Instructions Executed")])
                samp[op] += int(r[hdr.index("# Samples")])
            except (ValueError, IndexError):
                pass
    flush()


def full(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write(title + "\nncu --set full --clock-control none --import-source on (cold caches; one replay set per launch)\n")
        for r in data:
            f.write(f"\n--- launch {r[0]}: {r[hdr.index('Kernel Name')]}\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"    {w:72s} {r[i]:>16s} {units[i]}\n")
        opcode_mix(rep, f)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:5])
