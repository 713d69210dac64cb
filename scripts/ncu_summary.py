"""Condense an `ncu --page raw --csv` export into one line per kernel launch."""
import csv
import sys

WANT = [('gpu__time_duration.sum', 'ms'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps%'), ('launch__registers_per_thread', 'regs'),
        ('dram__bytes_read.sum', 'rdMB'), ('dram__bytes_write.sum', 'wrMB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'fp64%'),
        ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma%'),
        ('smsp__inst_executed.sum', 'Minst'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'st_long'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'st_short'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'st_wait'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'st_bar'),
        ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'st_math'),
        ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'st_mio'),
        ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'st_lg'),
        ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'st_nsel')]
SCALE = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3, 'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3, 'inst': 1e-6}


def main(path, header=""):
    rows = [r for r in csv.reader(open(path)) if len(r) > 20]
    hdr, units = rows[0], rows[1]
    out = []
    if header:
        out.append("# " + header)
    out.append('%-26s' % 'kernel' + ' '.join('%8s' % w[1] for w in WANT))
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '').replace('nele::', '')[:26]
        vals = []
        for w, _ in WANT:
            try:
                i = hdr.index(w)
                v = float(r[i]) * (1 if _.startswith('st_') else SCALE.get(units[i], 1))   # stall ratios carry the unit 'inst'
                vals.append('%8.2f' % v)
            except Exception:
                vals.append('%8s' % '-')
        out.append('%-26s' % name + ' '.join(vals))
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    sys.stdout.write(main(sys.argv[1], " ".join(sys.argv[2:])))
