"""Turn ncu outputs brought back in gpurun_out/ into the small summaries committed under profiles/.

    python tools/summarize_ncu.py full  gpurun_out/X.ncu-rep  profiles/NAME.json  "<how it was captured>"
    python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/NAME_summary.csv

`full`: one JSON object per captured kernel launch (the first launch goes to the top level, so
bench.py can read dram__bytes_* from it), metrics from `ncu --page raw`.
`launches`: per-kernel launch count, total device time and share from the
`--metrics gpu__time_duration.sum` launch list.
"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict

KEEP = [
    'Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum',
    'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum',
]


def full(report, out, source):
    raw = subprocess.run(['ncu', '-i', report, '--page', 'raw', '--csv'], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    launches = []
    for row in rows[2:]:
        entry = {}
        for name in KEEP:
            if name in header:
                i = header.index(name)
                entry[name] = {'value': row[i], 'unit': units[i]}
        stalls = {}
        for i, name in enumerate(header):
            if name.startswith('smsp__average_warps_issue_stalled_') and name.endswith('_per_issue_active.ratio'):
                try:
                    value = float(row[i])
                except ValueError:
                    continue
                if value >= 0.02:
                    stalls[name[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]] = round(value, 3)
        entry['stall_per_issue'] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        launches.append(entry)
    result = {'_source': source}
    result.update(launches[0])
    if len(launches) > 1:
        result['other_launches'] = launches[1:]
    with open(out, 'w') as f:
        json.dump(result, f, indent=1)
    print(f'{out}: {len(launches)} launch(es)')


def launches(log, out):
    lines = [line for line in open(log) if not line.startswith('==')]
    rows = list(csv.DictReader(io.StringIO(''.join(lines))))
    total = defaultdict(float)
    count = defaultdict(int)
    for row in rows:
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        value = float(row['Metric Value'].replace(',', ''))
        scale = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'nsecond': 1e-6, 'ms': 1.0, 'msecond': 1.0,
                 'second': 1e3, 's': 1e3}[row['Metric Unit']]
        name = row['Kernel Name'].split('(')[0][:90]
        total[name] += value * scale
        count[name] += 1
    grand = sum(total.values())
    with open(out, 'w') as f:
        writer = csv.writer(f)
        writer.writerow(['kernel', 'launches', 'total_ms', 'share'])
        for name in sorted(total, key=lambda k: -total[k]):
            writer.writerow([name, count[name], f'{total[name]:.3f}', f'{total[name] / grand:.4f}'])
    print(f'{out}: {len(total)} kernels, {grand:.1f} ms')


if __name__ == '__main__':
    if sys.argv[1] == 'full':
        full(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        launches(sys.argv[2], sys.argv[3])
