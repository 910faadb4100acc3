"""SASS evidence per kernel: counts of the mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG/UBLKCP (TMA / bulk copies), SYNCS (mbarrier),
plus legacy HMMA (mma.sync) which must be absent, registers and shared memory from the ELF.
    python tools/sass_summary.py [kernel-name-regex] > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'mpqe_b200', '_C', 'libmpqe_b200.so')
WATCH = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'HMMA', 'LDGSTS', 'FFMA',
         'LDG', 'STG', 'LDS', 'STS', 'ATOM', 'RED', 'BAR', 'MEMBAR', 'CCTL']


def main():
    pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r'Function (\S+):', line)
        if m:
            cur = m.group(1)
        m = re.search(r'REG:(\d+).*SHARED:(\d+)', line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
    counts = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = m.group(1)
            counts[name] = collections.Counter()
            continue
        m = re.search(r'/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P[T\d]+\s+)?([A-Z0-9_.]+)', line)
        if m and name:
            op = m.group(1)
            counts[name]['total'] += 1
            for w in WATCH:
                if op == w or op.startswith(w + '.'):
                    counts[name][w] += 1
    print('SASS summary of %s (cuobjdump -sass; sm_100a)' % os.path.relpath(LIB, ROOT))
    print('columns: instructions | registers, static shared bytes | watched mnemonics (count)')
    for name, c in counts.items():
        m = re.search(r'_cu_[0-9a-f]{8}(\d\d)([A-Za-z_0-9]+)', name)
        short = m.group(2)[:int(m.group(1))] if m else name
        if pat and not pat.search(short):
            continue
        reg, shm = usage.get(name, (None, None))
        marks = '  '.join('%s %d' % (w, c[w]) for w in WATCH if c[w])
        print('%-34s %6d | regs %s smem %s | %s' % (short, c['total'], reg, shm, marks))


if __name__ == '__main__':
    main()
