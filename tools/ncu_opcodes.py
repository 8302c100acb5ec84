"""Opcode histogram and execution-count profile of one kernel from an ncu report captured with --import-source on.
  python tools/ncu_opcodes.py <report.ncu-rep> <kernel regex>
"""
import collections
import csv
import subprocess
import sys

raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--kernel-name', 'regex:' + sys.argv[2]],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, ii, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = 0
by, samp = collections.Counter(), collections.Counter()
lines = []
for r in rows[2:]:
    if len(r) <= ii:
        continue
    try:
        n, s = int(r[ii]), int(r[isamp])
    except ValueError:
        continue
    parts = r[ia].split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    op = op.split('.')[0]
    by[op] += n
    samp[op] += s
    tot += n
    lines.append((n, s, r[ia].strip()))
print('static instructions', len(lines), 'executed (warp level)', tot)
for op, n in by.most_common(28):
    print('%-10s %11d %5.1f%%  stall samples %d' % (op, n, 100 * n / tot, samp[op]))
cnt = collections.Counter(n for n, _, _ in lines)
print('execution-count classes (count x static instructions):')
for n, c in sorted(cnt.items(), key=lambda kv: -kv[0] * kv[1])[:10]:
    print('   executed %9d times: %4d instructions = %5.1f%% of all' % (n, c, 100.0 * n * c / tot))
