#!/usr/bin/env python
"""SASS opcode histogram of the shipped kernels (cuobjdump -sass of the object files built by
torbi_b200/csrc/Makefile): per kernel the instruction count by mnemonic, the FADD2 / FMNMX3 / LDS / LDGSTS
counts, how many FADD2 carry an operand-reuse flag, and the cluster / DSMEM / async-copy / strong-access
instructions that identify the decompositions.  No GPU needed.

    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt
"""
import collections
import pathlib
import re
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
BUILD = ROOT / 'torbi_b200' / 'csrc' / '_build'
# object file -> substrings selecting the instantiations that ship on the measured paths
KERNELS = {
    'forward_tiled_hist.o': ['forward_tiled_hist_kernelILi4ELj2ELb0E', 'forward_tiled_hist_kernelILi4ELj2ELb1E'],
    'forward_tiled.o': ['forward_tiled_kernelILi4ELj2ELb0E'],
    'forward_split.o': ['forward_split_kernel'],
    'forward_cluster.o': ['forward_cluster_kernelILj2E'],
    'forward_wide.o': ['forward_wide_kernelILi1ELj2E'],
    'forward_flow.o': ['forward_flow_kernelILi1ELi10E', 'forward_flow_kernelILi2ELi8E'],
    'backtrace.o': ['backtrace_recompute_kernelILi12E', 'backtrace_compose_kernelILi4E'],
    'entropy.o': ['entropy_kernelILi0E'],
}
MARKERS = ['FADD2', 'FMNMX3', 'FMNMX', 'FADD', 'LDS', 'LDGSTS', 'LDG', 'STG', 'SHFL', 'REDUX', 'BAR', 'SYNCS', 'UCGABAR',
           'MEMBAR', 'ATOM', 'RED', 'MUFU', 'CCTL', 'MAPA', 'ST.E', 'STS']


def functions(obj):
    out = subprocess.run(['cuobjdump', '-sass', str(obj)], capture_output=True, text=True, check=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        elif name and re.match(r'\s*/\*[0-9a-f]{4}\*/', line):
            body.append(line)
    if name:
        yield name, body


def main():
    for obj, wanted in KERNELS.items():
        path = BUILD / obj
        if not path.exists():
            print(f'# {obj}: not built')
            continue
        seen = set()
        for name, body in functions(path):
            key = next((w for w in wanted if w in name), None)
            if key is None or key in seen:
                continue
            seen.add(key)
            ops = collections.Counter()
            reuse = collections.Counter()
            special = collections.Counter()
            for line in body:
                m = re.search(r'\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
                if not m:
                    continue
                op = m.group(1)
                base = op.split('.')[0]
                ops[base] += 1
                if '.reuse' in line:
                    reuse[base] += 1
                for tag in ('STRONG.GPU', 'STRONG.SYS', 'LDGSTS', 'UCGABAR', 'SYNCS', 'MAPA', 'REDUX', 'CREDUX', 'UTMA', 'shared::cluster'):
                    if tag in line:
                        special[tag] += 1
                if base in ('ST', 'LD') and '.E' in op and 'desc' not in line:
                    special['generic/cluster ' + op] += 1
            total = sum(ops.values())
            print(f'## {obj}: {name}\n   {total} instructions')
            print('   ' + ', '.join(f'{op} {n}' for op, n in ops.most_common(18)))
            print('   operand-reuse flags: ' + (', '.join(f'{op} {n} of {ops[op]}' for op, n in reuse.most_common(6)) or 'none'))
            print('   ' + (', '.join(f'{tag} {n}' for tag, n in sorted(special.items())) or '-'))
            print()


if __name__ == '__main__':
    main()
