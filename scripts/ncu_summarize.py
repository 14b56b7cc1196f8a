"""Summarise an `ncu --set full` report into profiles/: per-launch table (markdown) + per-class DRAM traffic (JSON).
Usage (CPU box): ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv; python scripts/ncu_summarize.py raw.csv OUT_PREFIX [LAST_N]
LAST_N: only the last N launches of the capture (the launches of the last prove)."""
import csv
import json
import sys

CLASSES = [('hash_columns', 'hash_columns'), ('merkle_level', 'merkle_build'), ('merkle_top', 'merkle_build'), ('merkle_sub', 'merkle_build'),
           ('merkle_tail', 'merkle_build'), ('compose', 'compose'), ('ntt_pass', 'ntt'), ('ntt2_', 'ntt'), ('fri_fold', 'fri_fold'),
           ('fri_challenge', 'fri_fold'), ('fri_tail', 'fri_tail'), ('derive_coeffs', 'derive_coeffs'), ('gather_chunks', 'gather_queries')]
COLS = [('time_us', 'gpu__time_duration.sum'), ('regs', 'launch__registers_per_thread'), ('grid', 'launch__grid_size'),
        ('dram_rd_MB', 'dram__bytes_read.sum'), ('dram_wr_MB', 'dram__bytes_write.sum'),
        ('sm_thr_pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed'), ('dram_thr_pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
        ('issue_active_pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active'), ('alu_pipe_pct', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
        ('fmaheavy_pipe_pct', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'), ('warps_active_pct', 'sm__warps_active.avg.pct_of_peak_sustained_active')]


def main(raw, prefix, last_n=0):
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    if last_n:
        rows = rows[:2] + rows[2:][-last_n:]
    ki = hdr.index('Kernel Name')
    # fused single-GPU commit (merkle_span_kernel + the leaf-hashing tree top): one class, as in bench.py's kernels_ms_per_step
    fused = any('merkle_span' in r[ki] for r in rows[2:])
    classes_of = ([('merkle_span', 'merkle_commit'), ('merkle_top', 'merkle_commit')] if fused else []) + CLASSES
    table, classes = [], {}
    for r in rows[2:]:
        name = r[ki]
        rec = {'kernel': name.split('(')[0].replace('void gs::', '').replace('gs::', '')[:44]}
        for key, col in COLS:
            i = hdr.index(col)
            v = float(r[i].replace(',', '')) if r[i] else 0.0
            u = units[i]
            if key == 'time_us' and u in ('ns', 'nsecond'):
                v /= 1e3
            elif key == 'time_us' and u in ('ms', 'msecond'):
                v *= 1e3
            if key.endswith('_MB'):
                v = v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1.0)
            rec[key] = round(v, 2)
        table.append(rec)
        for pat, cls in classes_of:
            if pat in name:
                c = classes.setdefault(cls, {'launches_captured': 0, 'time_us': 0.0, 'dram_read_MB': 0.0, 'dram_write_MB': 0.0})
                c['launches_captured'] += 1; c['time_us'] += rec['time_us']
                c['dram_read_MB'] += rec['dram_rd_MB']; c['dram_write_MB'] += rec['dram_wr_MB']
                break
    for c in classes.values():
        c['dram_bytes_captured'] = int((c['dram_read_MB'] + c['dram_write_MB']) * 1e6)
    with open(prefix + '_launch_table.md', 'w') as f:
        keys = ['kernel'] + [k for k, _ in COLS]
        f.write('| ' + ' | '.join(keys) + ' |\n|' + '---|' * len(keys) + '\n')
        for rec in table:
            f.write('| ' + ' | '.join(str(rec[k]) for k in keys) + ' |\n')
    json.dump(classes, open(prefix + '_classes.json', 'w'), indent=1)
    print(json.dumps(classes, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
