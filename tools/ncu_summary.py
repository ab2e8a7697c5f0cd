"""Write the profiles/ summary of one ncu report: key raw metrics + hot SASS instructions by stall samples.
usage: python tools/ncu_summary.py X.ncu-rep OUT.txt   (needs `ncu` on PATH; runs without a GPU)"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg.per_second"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, d = rows[0], rows[1], rows[2]
    idx = {k: i for i, k in enumerate(h)}
    lines = ["%-75s %s" % ("Kernel Name", d[idx["Kernel Name"]])]
    for k in KEYS:
        if k in idx:
            lines.append("%-75s %s %s" % (k, d[idx[k]], u[idx[k]]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    open("/tmp/_ncu_src.csv", "w").write(src)
    hot = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_hot.py"), "/tmp/_ncu_src.csv", "30"],
                         capture_output=True, text=True).stdout
    lines.append("")
    lines.append("--- warp-stall sampling (ncu --page source), top SASS instructions")
    lines.append(hot)
    open(out, "w").write("\n".join(lines))
    print("\n".join(lines[:22]))


if __name__ == "__main__":
    main()
