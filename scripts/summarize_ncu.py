"""scripts/summarize_ncu.py REP.ncu-rep OUT.txt -- the metrics quoted in DESIGN.md / profiles/README.md, one block per launch.
Reads the report here (no GPU needed): `ncu -i REP --page raw --csv`."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
    "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    with open(out, "w") as f:
        f.write("# %s  (ncu --set full --clock-control none; per-launch values, cold cache, serialised)\n" % rep)
        for r in rows[2:]:
            f.write("\n== %s\n" % r[idx["Kernel Name"]][:160])
            for w in WANT:
                if w in idx:
                    f.write("%-72s %s %s\n" % (w, r[idx[w]], units[idx[w]]))
            vals = sorted([(float(r[idx[h]]), h) for h in stalls if r[idx[h]] not in ("", "n/a")], reverse=True)[:6]
            f.write("top stall reasons (warps per issue-active cycle): " + ", ".join(
                "%s=%.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, h in vals) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
