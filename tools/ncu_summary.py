"""Selected metrics of every launch in an ncu report (ncu -i REPORT --page raw --csv piped through this): duration, DRAM bytes,
L2 throughput, tensor-pipe activity, occupancy facts.  Usage: python tools/ncu_summary.py report.ncu-rep > summary.txt"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__t_bytes.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    cols = [(m, h.index(m)) for m in METRICS if m in h]
    print("# " + path)
    print("# columns: kernel | " + " | ".join("%s [%s]" % (m, units[i]) for m, i in cols))
    tot_t = tot_b = 0.0
    n = 0
    for r in rows[2:]:
        vals = [r[i] for _, i in cols]
        print("%-34s | %s" % (r[ki].split("(")[0].replace("<unnamed>::", "")[:34], " | ".join(vals)))
        try:
            tot_t += float(r[h.index("gpu__time_duration.sum")].replace(",", ""))
            tot_b += float(r[h.index("dram__bytes_read.sum")].replace(",", "")) + float(r[h.index("dram__bytes_write.sum")].replace(",", ""))
            n += 1
        except Exception:
            pass
    print("# launches %d, total duration %.1f %s, mean DRAM traffic per launch %.3f %s" % (
        n, tot_t, units[h.index("gpu__time_duration.sum")], tot_b / max(1, n), units[h.index("dram__bytes_read.sum")]))


if __name__ == "__main__":
    main(sys.argv[1])
