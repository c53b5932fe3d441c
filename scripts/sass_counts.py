"""Static SASS evidence for profiles/: per kernel of libfibb200.so the instruction mix (packed fp32 FFMA2 / FMUL2 /
FADD2, scalar FFMA / FMUL / FADD, MUFU, local-memory LDL / STL) and the TMA / mbarrier mnemonics.
    python scripts/sass_counts.py [regex] > profiles/r2_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'fib_tf_b200', 'libfibb200.so')
pat = sys.argv[1] if len(sys.argv) > 1 else ''
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
cur, stats = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        stats[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        op = m.group(1)
        base = op.split('.')[0]
        stats[cur][base] += 1
        stats[cur]['_total'] += 1
        if base in ('UTMALDG', 'UTMASTG', 'SYNCS', 'UBLKCP'):
            stats[cur]['tma:' + op] += 1
print('# cuobjdump -sass fib_tf_b200/libfibb200.so, static instruction counts per kernel (whole kernel body,')
print('# all unrolled rows / cells); TMA = UTMALDG / UTMASTG (cp.async.bulk.tensor), SYNCS = mbarrier ops')
print('%-118s %6s %6s %6s %6s %5s %5s %5s %5s %4s %4s %s' % ('kernel', 'total', 'FFMA2', 'FMUL2', 'FADD2', 'FFMA', 'FMUL', 'FADD',
                                                              'MUFU', 'LDL', 'STL', 'TMA / mbarrier'))
for k, c in stats.items():
    if pat and not re.search(pat, k):
        continue
    name = re.sub(r'fib::', '', k)
    name = re.sub(r'\(.*', '', name)[:118]
    tma = ' '.join('%s=%d' % (n[4:], v) for n, v in sorted(c.items()) if n.startswith('tma:'))
    print('%-118s %6d %6d %6d %6d %5d %5d %5d %5d %4d %4d %s' % (name, c['_total'], c['FFMA2'], c['FMUL2'], c['FADD2'], c['FFMA'],
                                                                 c['FMUL'], c['FADD'], c['MUFU'], c['LDL'], c['STL'], tma))
