#!/usr/bin/env python
"""SASS / PTX census of the shipped library: per kernel, how many tensor-core (UTCHMMA/UTCQMMA...), tensor-memory
(LDTM/STTM), TMA (UTMALDG, UBLKCP, UBLKPF) and barrier (UTCBAR, SYNCS) instructions the sm_100a cubin holds.

    python scripts/sass_census.py [iodine_b200/lib/libiodine_b200.so] > profiles/r2_sass_census.md
"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else 'iodine_b200/lib/libiodine_b200.so'
OPS = ['UTCHMMA', 'UTCQMMA', 'UTCMMA', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'UBLKPF', 'UTCBAR', 'SYNCS', 'HMMA', 'FFMA', 'MUFU']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.split('\n')
    return dict(zip(names, out))


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    archs = sorted(set(re.findall(r'arch = (sm_\w+)', sass)))
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            op = m.group(1).split('.')[0]
            for o in OPS:
                if op == o or (o in ('UTCHMMA', 'UTCQMMA', 'UTCMMA') and op.startswith(o)):
                    per[cur][o] += 1
    names = demangle(list(per))
    ptx = subprocess.run(['cuobjdump', '-ptx', LIB], capture_output=True, text=True).stdout
    print('# SASS census of `%s`' % LIB)
    print()
    print('`cuobjdump -sass`: cubin architectures = %s.  Counts are static instructions per kernel.' % ', '.join(archs))
    print()
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print('Library totals: ' + ', '.join('%s x%d' % (o, tot[o]) for o in OPS if tot[o]))
    print()
    cols = [o for o in OPS if tot[o] and o not in ('FFMA', 'MUFU')]
    print('| kernel | ' + ' | '.join(cols) + ' | FFMA |')
    print('|---|' + '---|' * (len(cols) + 1))
    rows = []
    for k, c in per.items():
        if not any(c[o] for o in cols) and c['FFMA'] < 64:
            continue
        short = re.sub(r'^void iod::', '', names.get(k, k))
        short = re.sub(r'\(.*$', '', short)
        rows.append((-(c['UTCHMMA'] + c['UTCQMMA'] + c['UTCMMA']), short, c))
    for _, short, c in sorted(rows, key=lambda r: (r[0], r[1])):
        print('| `%s` | ' % short[:90] + ' | '.join(str(c[o]) for o in cols) + ' | %d |' % c['FFMA'])
    print()
    n_tc = len(re.findall(r'tcgen05\.mma', ptx))
    print('`cuobjdump -ptx | grep -c`: tcgen05.mma x%d, tcgen05.ld x%d, tcgen05.commit x%d, cp.async.bulk x%d, '
          'mbarrier x%d' % (n_tc, len(re.findall(r'tcgen05\.ld', ptx)), len(re.findall(r'tcgen05\.commit', ptx)),
                            len(re.findall(r'cp\.async\.bulk', ptx)), len(re.findall(r'mbarrier\.', ptx))))
    if not ptx.strip():
        print('(the library embeds no PTX: `-gencode arch=compute_100a,code=sm_100a` keeps SASS only)')


if __name__ == '__main__':
    main()
