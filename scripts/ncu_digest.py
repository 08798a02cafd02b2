"""Turn `ncu --set full` reports into the text summaries kept under profiles/ and into profiles/r02_traffic.json
(DRAM bytes per launch, read by bench.py for `roofline.traffic`).

    python scripts/ncu_digest.py <tag> <report.ncu-rep> [<label> <algorithmic bytes> <positions>]

Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU), keeps the metrics the profiling recipe
names (B200_PROFILING.md) and writes profiles/<tag>.txt.  With a label the DRAM traffic of the LAST kernel in the report
is added to profiles/r02_traffic.json under that label (bench.py key: "<class> <layer>")."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def rows(rep):
    if rep.endswith(".csv"):          # already exported on the GPU box (`ncu -i <rep> --page raw --csv`)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith('"')]
    r = list(csv.reader(io.StringIO("\n".join(lines))))
    names, units = r[0], r[1]
    return names, units, r[2:]


def main():
    tag, rep = sys.argv[1], sys.argv[2]
    names, units, data = rows(rep)
    idx = {n: i for i, n in enumerate(names)}
    lines = ["ncu --set full --clock-control none  (%s)" % os.path.basename(rep)]
    last = None
    for row in data:
        lines.append("Kernel Name".ljust(90) + row[idx["Kernel Name"]])
        rec = {}
        for k in KEEP:
            if k in idx:
                lines.append("%-90s %-16s %s" % (k, units[idx[k]], row[idx[k]]))
                rec[k] = (row[idx[k]], units[idx[k]])
        last = rec
        lines.append("")
    with open(os.path.join(ROOT, "profiles", tag + ".txt"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))
    if len(sys.argv) > 3 and last:
        label, alg, pos = sys.argv[3], float(sys.argv[4]), int(sys.argv[5])

        def tobytes(key):
            v, u = last[key]
            return float(v.replace(",", "")) * UNIT.get(u, 1.0)

        tj = os.path.join(ROOT, "profiles", "r02_traffic.json")
        d = json.load(open(tj)) if os.path.exists(tj) else {
            "_source": "ncu --set full --clock-control none, one launch at the shape named per entry (scripts/gpu_evidence.sh); "
                       "dram__bytes_read.sum + dram__bytes_write.sum", "per_launch": {}}
        d["per_launch"][label] = {"dram_bytes": tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"),
                                  "algorithmic_bytes": alg, "positions": pos, "report": tag}
        json.dump(d, open(tj, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
