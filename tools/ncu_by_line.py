"""Attribute executed instructions / stall samples of one kernel (ncu --page source --print-source sass CSV) to the
source lines of a chosen file, following nvdisasm -gi inline chains.
usage: ncu_by_line.py <sass.csv> <nvdisasm_gi.txt> <kernel-substring> <file-to-attribute-to> [top]"""
import collections
import csv
import re
import sys

sass_csv, dis, kern, attr_file = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
import os
skip_from = int(os.environ.get('SKIP_FROM', '0'))   # ignore frames at or after this line (a thin wrapper at the end of the file)
func, chain, addr2chain = None, [], {}
pending = []
for ln in open(dis):
    m = re.match(r'\s*\.text\.(\S+):', ln)
    if m:
        func = m.group(1)
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        pending.append((m.group(1).split('/')[-1], int(m.group(2))))
        if m.group(3):
            pending.append((m.group(3).split('/')[-1], int(m.group(4))))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/\s+(\S.*?);', ln)
    if m:
        if pending:
            chain, pending = pending, []
        if func and kern in func:
            addr2chain[int(m.group(1), 16)] = chain
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
inst, samp = collections.Counter(), collections.Counter()
base = None
for r in rows[2:]:
    if r and r[0].startswith('Kernel Name'):
        break
    try:
        a, e, s = int(r[0], 16), int(r[ix['Instructions Executed']]), int(r[ix['# Samples']])
    except (ValueError, IndexError):
        continue
    base = a if base is None else base
    ch = addr2chain.get(a - base, [])
    key = None
    for f, l in reversed(ch):        # outermost frame first
        if f == attr_file and not (skip_from and l >= skip_from):
            key = l
            break
    if key is None:
        key = ('other', ch[-1] if ch else None)
    inst[key] += e
    samp[key] += s
ti, ts = sum(inst.values()), sum(samp.values())
src = open(sys.argv[6]).read().split('\n') if len(sys.argv) > 6 else None
print(f"total warp-instructions {ti}, stall samples {ts}")
for k, v in inst.most_common(top):
    text = src[k - 1].strip()[:100] if src and isinstance(k, int) else ''
    print(f"{100 * v / ti:5.1f}% inst {100 * samp[k] / ts:5.1f}% time  line {k}: {text}")
