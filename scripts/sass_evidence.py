"""Static evidence for profiles/: per-kernel SASS mnemonic counts (tcgen05 = UTCHMMA/UTCBAR/LDTM, bulk copies =
UBLKCP, mbarriers = SYNCS.*, cp.async = LDGSTS) and ptxas resource usage of the in-tree library.

    python scripts/sass_evidence.py > profiles/rNN_sass_static.md     (no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'cutie_b200', 'lib', 'libcutie_b200.so')
WATCH = ('UTCHMMA', 'UTCBAR', 'LDTM', 'UBLKCP', 'SYNCS', 'LDGSTS', 'HMMA', 'FFMA', 'LDG', 'STG', 'ATOMG', 'RED', 'LDS', 'STS',
         'SHFL', 'BAR', 'FMNMX3')


def demangle(s):
    return subprocess.run(['c++filt', s], capture_output=True, text=True).stdout.strip()


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.search(r'Function (\S+):', line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'REG:(\d+).*?SHARED:(\d+)', line)
        if m and fn:
            usage[fn] = (int(m.group(1)), int(m.group(2)))
    counts = collections.OrderedDict()
    fn = None
    op = re.compile(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Za-z0-9_]+)*)')
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            continue
        m = op.match(line)
        if m and fn:
            counts[fn][m.group(1)] += 1
            counts[fn]['_total'] += 1
    print('# Static SASS evidence (cuobjdump -sass / -res-usage of cutie_b200/lib/libcutie_b200.so, sm_100a)\n')
    print('Counts are static instructions in the kernel body (loops execute them many times).  `UTCHMMA` = tcgen05.mma,')
    print('`UTCBAR` = tcgen05.commit, `LDTM` = tcgen05.ld, `UBLKCP` = cp.async.bulk, `SYNCS` = mbarrier ops, `LDGSTS` = cp.async.\n')
    print('| kernel | regs | static smem | instrs | ' + ' | '.join(WATCH) + ' |')
    print('|---|---:|---:|---:|' + '---:|' * len(WATCH))
    for fn, c in counts.items():
        name = demangle(fn)
        name = re.sub(r'\(.*', '', name).replace('void ', '')
        r, s = usage.get(fn, ('?', '?'))
        print(f'| `{name}` | {r} | {s} | {c["_total"]} | ' + ' | '.join(str(c.get(w, 0)) for w in WATCH) + ' |')


if __name__ == '__main__':
    sys.exit(main())
