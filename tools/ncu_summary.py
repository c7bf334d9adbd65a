"""One-line-per-kernel summary of an ncu report (raw page): duration, occupancy, issue rate, memory, top stalls.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep
"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__inst_executed.sum", "winst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("l1tex__t_sector_hit_rate.pct", "L1hit%"), ("lts__t_sector_hit_rate.pct", "L2hit%"),
        ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%")]


def main():
    text = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        out = [name[:48]]
        for k, label in KEYS:
            if k in hdr:
                v = r[hdr.index(k)]
                u = units[hdr.index(k)]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                out.append(f"{label}={v}{u if label.startswith('dram') and label != 'dram%' else ''}")
        stalls = [(float(r[i]), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[i]]
        stalls.sort(reverse=True)
        out.append("stalls: " + ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for v, h in stalls[:5]))
        print("  ".join(out))


if __name__ == "__main__":
    main()
