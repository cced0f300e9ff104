"""Summarise an `ncu --set full --import-source on` report as text: key metrics of every captured launch plus the SASS
instructions that collected the most warp-stall samples (what the evidence files under profiles/ are made of).

    python scripts/ncu_summary.py report.ncu-rep [top_n] > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed',
        'l1tex__m_l1tex2xbar_write_bytes.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def ncu(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    print('# %s' % rep)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('\n## launch %s  %s' % (d.get('ID', '?'), d.get('Kernel Name', '')[:110]))
        for k in KEYS:
            if d.get(k) not in (None, ''):
                print('  %-78s %s %s' % (k, d[k], units[hdr.index(k)]))
    src = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'source', '--csv', '--print-source', 'sass']))))
    starts = [i for i, r in enumerate(src) if r and r[0] == 'Kernel Name']
    if not starts:
        return
    s0 = starts[0]
    e0 = starts[1] if len(starts) > 1 else len(src)
    h = src[s0 + 1]
    body = src[s0 + 2:e0]
    ix = {n: i for i, n in enumerate(h)}
    stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
    tot = sum(int(r[ix['# Samples']]) for r in body)
    print('\n## first launch: SASS instructions with the most warp-stall samples (%d samples in total, %d instructions)' % (tot, len(body)))
    print('# index  samples  executed  top stall reasons  |  instruction')
    top = sorted(range(len(body)), key=lambda i: -int(body[i][ix['# Samples']]))[:top_n]
    for i in sorted(top):
        r = body[i]
        st = sorted(((int(r[ix[n]]), n[6:]) for n in stalls), reverse=True)[:2]
        print('%6d %8s %9s  %-44s | %s' % (i, r[ix['# Samples']], r[ix['Instructions Executed']],
                                         ', '.join('%s %d' % (n, v) for v, n in st if v), r[ix['Source']].strip()[:90]))


main()
