"""Summarise an `ncu --set full` capture of the tensor-core conv launches of one 8-image forward.

    ncu --set full --clock-control none -k regex:tapgemm_tc --launch-skip <first step's launches> -o gpurun_out/X \
        python scripts/profile_step.py infer                               # on the GPU box
    python scripts/ncu_conv_summary.py gpurun_out/X.ncu-rep <launches per step> profiles/r1_conv_kernels_ncu.txt \
        profiles/r1_conv_traffic.json                                          # here (ncu -i works without a GPU)

Takes the LAST <launches per step> rows of the report (one whole forward), writes the per-launch table and the DRAM
traffic total that bench.py reports as roofline.traffic.
"""
import csv
import io
import json
import subprocess
import sys

rep, per_step, out_txt, out_json = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
data = data[-per_step:]


def col(name):
    return hdr.index(name)


def val(r, name, scale_by_unit=True):
    i = col(name)
    v = float(r[i].replace(',', '')) if r[i] not in ('', 'n/a') else 0.0
    u = units[i]
    if not scale_by_unit:
        return v
    return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1e3, 'us': 1.0, 'ns': 1e-3, 's': 1e6}.get(u, 1.0)


lines, tot_t, tot_r, tot_w = [], 0.0, 0.0, 0.0
for i, r in enumerate(data):
    name = r[col('Kernel Name')]
    kind = 'strip' if 'strip' in name else ('wgrad' if 'wgrad' in name else 'fwd')
    t = val(r, 'gpu__time_duration.sum')
    rd, wr = val(r, 'dram__bytes_read.sum'), val(r, 'dram__bytes_write.sum')
    lts = val(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed', False) if 'lts__throughput.avg.pct_of_peak_sustained_elapsed' in hdr else 0.0
    regs = r[col('launch__registers_per_thread')] if 'launch__registers_per_thread' in hdr else '?'
    grid = r[col('Grid Size')] if 'Grid Size' in hdr else '?'
    tot_t += t; tot_r += rd; tot_w += wr
    lines.append('%2d %-5s grid=%s time=%.1f read=%.1f write=%.1f lts=%.1f regs=%s  %s' %
                 (i, kind, grid.replace(' ', ''), t, rd / 1e6, wr / 1e6, lts, regs, name.split('(')[0][-48:]))
with open(out_txt, 'w') as f:
    f.write('# ncu --set full --clock-control none, the %d tensor-core conv launches of ONE 8-image 512x512 forward '
            '(profile_step.py infer)\n# idx kernel grid time_us dram_read_MB dram_write_MB lts_throughput_%% regs name\n' % per_step)
    f.write('\n'.join(lines) + '\n')
    f.write('# total: time %.1f us (cold-cache, serialised), dram read %.1f MB, write %.1f MB\n' % (tot_t, tot_r / 1e6, tot_w / 1e6))
with open(out_json, 'w') as f:
    json.dump({'dram_bytes_per_step': tot_r + tot_w, 'dram_read': tot_r, 'dram_write': tot_w, 'ncu_time_us': tot_t,
               'launches': per_step,
               'source': '%s (ncu --set full, %d conv launches of one 8x512x512 forward)' % (out_txt, per_step)}, f, indent=1)
print(open(out_txt).read())
