"""Per-launch issue / stall summary from an `ncu --page raw --csv` export (warp-state breakdown of profiles/).
Usage: python scripts/ncu_stalls.py raw.csv [kernel-substring]"""
import csv
import sys


def main(path, sub=''):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
    cols = [('us', 'gpu__time_duration.sum'), ('regs', 'launch__registers_per_thread'), ('grid', 'launch__grid_size'),
            ('issue%', 'smsp__issue_active.avg.pct_of_peak_sustained_active'), ('elig', 'smsp__warps_eligible.avg.per_cycle_active'),
            ('warps', 'smsp__warps_active.avg.per_cycle_active'), ('Minst', 'smsp__inst_executed.sum'),
            ('dramRdMB', 'dram__bytes_read.sum'), ('dramWrMB', 'dram__bytes_write.sum'), ('l2%', 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
            ('alu%', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'), ('fmaH%', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'),
            ('smemConfl', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')]
    for r in rows[2:]:
        if sub not in r[ki]:
            continue
        out = []
        for name, col in cols:
            if col not in hdr:
                continue
            i = hdr.index(col)
            v = float(r[i].replace(',', '') or 0)
            u = units[i]
            if name == 'us' and u.startswith('ms'):
                v *= 1e3
            if name == 'us' and u.startswith('ns'):
                v /= 1e3
            if name == 'Minst':
                v /= 1e6
            if name.startswith('dram'):
                v = v / 1e6 if u in ('byte', 'B') else (v * 1e3 if u.startswith('G') else (v / 1e3 if u.startswith('K') else v))
            out.append(f'{name}={v:.2f}')
        st = sorted(((float(r[hdr.index(h)].replace(',', '') or 0), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
                     for h in stalls), reverse=True)
        print(r[ki].replace('void gs::', '')[:60])
        print('   ', ' '.join(out))
        print('    stalls/issue:', ', '.join(f'{n}={v:.2f}' for v, n in st[:9]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '')
