"""Developer tool: condensed text summary of one `ncu --set full --import-source on` report.

  python tools/ncu_summary.py gpurun_out/X.ncu-rep [n_units] > profiles/rNN_ncu_X_summary.txt

Prints duration, DRAM traffic, pipe utilisation, issue / stall picture and the dynamic opcode
histogram (per unit when n_units, e.g. the number of 2 KiB blocks of the launch, is given).
"""
import collections
import csv
import io
import re
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = page(rep, "raw")
    hdr, unit, val = raw[0], raw[1], raw[2]
    m = {h: (v, u) for h, u, v in zip(hdr, unit, val)}
    print("kernel:", m.get("Kernel Name", ("?",))[0])
    want = [
        "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
    ]
    for k in want:
        if k in m:
            print("%-82s %-10s %s" % (k, m[k][1], m[k][0]))
    print("\nstall reasons (warps per issue slot):")
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            v = float(m[h][0].replace(",", ""))
            if v >= 0.05:
                print("  %-28s %.3f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
    src = page(rep, "source")
    h2 = src[1]
    i_s, i_e, i_n = h2.index("Source"), h2.index("Instructions Executed"), h2.index("# Samples")
    cnt, samp, tot, stot = collections.Counter(), collections.Counter(), 0, 0
    for r in src[2:]:
        if len(r) <= i_e:
            continue
        s = re.sub(r"^@!?U?P\d+\s+", "", r[i_s].strip())
        op = s.split()[0].rstrip(";") if s else "?"
        n = int(r[i_e])
        cnt[op] += n
        tot += n
        samp[op] += int(r[i_n])
        stot += int(r[i_n])
    print("\ndynamic warp-instructions: %d%s" % (tot, "  (%.1f per unit)" % (tot / units) if units else ""))
    for op, n in cnt.most_common(32):
        print("  %-22s %6.2f %%  %s samples %5.1f %%" % (op, 100.0 * n / tot,
              ("%8.1f per unit " % (n / units)) if units else "", 100.0 * samp[op] / max(stot, 1)))


if __name__ == "__main__":
    main()
