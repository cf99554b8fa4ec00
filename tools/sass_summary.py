#!/usr/bin/env python
"""Per-kernel SASS opcode counts of csrc/librgnn.so (`cuobjdump -sass`): the mnemonics that prove which hardware path a kernel uses
-- UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UBLKCP / UBLKRED (cp.async.bulk / cp.reduce.async.bulk),
LDGSTS (cp.async), SYNCS (mbarrier), FFMA / FFMA2 -- plus registers are left to `-Xptxas -v`.

    python tools/sass_summary.py > profiles/sass_opcodes_rNN.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'graph_neural_network_for_radar_perception_b200', 'csrc', 'librgnn.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTCBAR', 'UBLKCP', 'UBLKRED', 'UTMALDG', 'LDGSTS', 'SYNCS', 'FFMA2', 'FFMA', 'HFMA2', 'F2FP',
        'LDG', 'STG', 'LDS', 'STS', 'ATOMG', 'REDG', 'RED', 'SHFL', 'BAR', 'R2UR']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    name, counts, total = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
            counts[name] = collections.Counter()
            total[name] = 0
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and name is not None:
            op = m.group(1)
            total[name] += 1
            for k in KEYS:
                if op == k or (k in ('LDG', 'STG', 'LDS', 'STS', 'BAR', 'SHFL', 'ATOMG', 'REDG', 'RED') and op.startswith(k) and not (k == 'RED' and op.startswith('REDG'))):
                    counts[name][k] += 1
                    break
    print(f'{"kernel":58s} {"instr":>7s} ' + ' '.join(f'{k:>7s}' for k in KEYS))
    for n, c in counts.items():
        print(f'{n[:58]:58s} {total[n]:7d} ' + ' '.join(f'{c.get(k, 0):7d}' for k in KEYS))


if __name__ == '__main__':
    main()
