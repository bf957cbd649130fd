"""Summarise `ncu --set full` reports into the JSON bench.py reads its roofline `traffic` / tensor-pipe numbers from.

    python tools/ncu_to_metrics.py profiles/r2_kernel_metrics.json report1.ncu-rep [report2.ncu-rep ...]

Per kernel (the last launch of each name wins): duration, DRAM bytes read + written, tensor-pipe and issue utilisation,
executed warp instructions, registers, and the report it came from.
"""
import csv
import json
import os
import re
import subprocess
import sys

WANT = {
    'gpu__time_duration.sum': 'duration_us',
    'dram__bytes_read.sum': 'dram_read',
    'dram__bytes_write.sum': 'dram_write',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pipe_pct',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active': 'alu_pipe_pct',
    'smsp__inst_executed.sum': 'warp_instructions',
    'launch__registers_per_thread': 'registers',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed': 'smem_wavefront_pct',
}
SCALE = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1.0, 'us': 1.0, 'ms': 1e3, 'ns': 1e-3}


def short(name):
    m = re.match(r'(?:void )?(?:p2v::)?(\w+)(<[^>(]*>)?', name)
    base = m.group(1) if m else name
    args = m.group(2) if m and m.group(2) else ''
    return base + args.replace('(unsigned int)', '').replace(' ', '')


def main():
    out_path, reports = sys.argv[1], sys.argv[2:]
    result = {}
    for rep in reports:
        text = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(text.splitlines()))
        head, units = rows[0], rows[1]
        ki = head.index('Kernel Name')
        for r in rows[2:]:
            entry = {'report': os.path.basename(rep)}
            for i, h in enumerate(head):
                if h in WANT and r[i] != '':
                    v = float(r[i].replace(',', ''))
                    entry[WANT[h]] = round(v * SCALE.get(units[i], 1.0), 3)
            entry['dram_bytes'] = entry.pop('dram_read', 0.0) + entry.pop('dram_write', 0.0)
            result[short(r[ki])] = entry
    with open(out_path, 'w') as f:
        json.dump(result, f, indent=1, sort_keys=True)
    for k, v in sorted(result.items()):
        print('%-44s %8.1f us  dram %7.1f MB  tensor %5.1f %%  issue %5.1f %%' % (
            k, v.get('duration_us', 0), v.get('dram_bytes', 0) / 1e6, v.get('tensor_pipe_pct', 0), v.get('issue_active_pct', 0)))


if __name__ == '__main__':
    main()
