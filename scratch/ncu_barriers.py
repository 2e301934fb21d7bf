"""Samples spent waiting on each mbarrier (by shared-memory offset) in an ncu report's SASS view."""
import csv, subprocess, sys, io, re, collections
rep = sys.argv[1]; kre = sys.argv[2]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name-base", "mangled", "-k", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if r and r[0] == 'Address'); ix = {h: i for i, h in enumerate(hdr)}
data = []; seen = set()
for r in rows:
    if r and r[0].startswith('0x'):
        if r[0] in seen: break
        seen.add(r[0]); data.append(r)
tot = sum(int(r[ix['# Samples']]) for r in data)
# map branch targets of spin loops: a TRYWAIT followed by '@!P BRA target' where target is another TRYWAIT loop
addr_ix = {r[0]: i for i, r in enumerate(data)}
wait = collections.Counter(); spins = collections.Counter()
for i, r in enumerate(data):
    m = re.search(r'TRYWAIT P\d, \[.*\+0x([0-9a-f]+)\]', r[1])
    if m:
        off = int(m.group(1), 16)
        n = int(r[ix['# Samples']])
        for j in range(i + 1, min(i + 4, len(data))):   # the spin branch follows within a few instructions
            n += int(data[j][ix['# Samples']])
            if 'BRA' in data[j][1]: break
        wait[off] += n; spins[off] += int(r[ix['Instructions Executed']])
print('total samples', tot)
for off, n in sorted(wait.items()):
    print(hex(off), n, 'spins', spins[off])
