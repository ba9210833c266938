"""Per-kernel SASS instruction counts of libstb200.so (what proves a Blackwell-native kernel, B200_PROFILING.md):
  UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UTCBAR = tcgen05.commit,
  SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be 0).      python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(sys.argv[1]) if len(sys.argv) > 1 else Path(__file__).resolve().parent.parent / 'style-transfer-pytorch_b200' / 'libstb200.so'
out = subprocess.run(['cuobjdump', '-sass', str(lib)], capture_output=True, text=True, check=True).stdout
WANT = ['UTCHMMA', 'UTCQMMA', 'UTCMMA', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'HMMA', 'ELECT',
        'LDG', 'STG', 'ATOM', 'RED', 'BAR']
kern, counts, order = None, collections.defaultdict(collections.Counter), []
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        kern = m.group(1)
        order.append(kern)
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and kern:
        op = m.group(1)
        counts[kern]['_total'] += 1
        for w in WANT:
            if op.startswith(w):
                counts[kern][w] += 1
                break


def demangle(n):
    r = subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
    r = r.replace('stb::(anonymous namespace)::', '').replace('(anonymous namespace)::', '')
    r = re.sub(r'\(.*', '', r)
    return r.replace('void ', '')


print(f'# {lib.name}: SASS instruction counts per kernel (cuobjdump -sass, sm_100a)')
cols = [w for w in WANT if any(counts[k][w] for k in order)]
print(f'{"kernel":58s} {"instrs":>7s} ' + ' '.join(f'{c:>8s}' for c in cols))
tot = collections.Counter()
for k in order:
    name = demangle(k)
    print(f'{name[:58]:58s} {counts[k]["_total"]:7d} ' + ' '.join(f'{counts[k][c]:8d}' for c in cols))
    tot.update(counts[k])
print(f'{"TOTAL":58s} {tot["_total"]:7d} ' + ' '.join(f'{tot[c]:8d}' for c in cols))
