"""Top stall sites of one kernel of an `ncu --set full --import-source on` capture (no GPU needed):
    python scripts/ncu_hotspots.py gpurun_out/prof_top.ncu-rep <launch index> [top N]
Prints the SASS instructions with the most warp-stall samples, the dominant stall reason of each, and the share of
all samples they hold, plus the totals per stall reason."""
import csv
import subprocess
import sys


def main(path, launch, top=25):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass',
                          '--launch-skip', str(launch), '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][1])
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    seen, body = set(), []
    for r in rows[2:]:                                  # the page repeats the listing per source view: keep one
        if len(r) == len(hdr) and r[col['# Samples']].isdigit() and r[col['Address']] not in seen:
            seen.add(r[col['Address']])
            body.append(r)
    tot = sum(int(r[col['# Samples']]) for r in body) or 1
    per = {k: sum(int(r[col[k]]) for r in body) for k in reasons}
    print('samples', tot, ' by reason:', ', '.join('%s %.1f%%' % (k[6:], 100.0 * v / tot)
                                                   for k, v in sorted(per.items(), key=lambda kv: -kv[1]) if v * 200 > tot))
    # shares per code range, split at the first bulk/TMA copy, first UTCHMMA and first LDTM of the warp-specialised
    # kernels (rough: an epilogue's wait for its accumulator precedes its LDTM and lands in the previous range)
    marks = []
    for tag, pat in (('producer', ('UBLKCP', 'UTMALDG')), ('issuer', ('UTCHMMA',)), ('epilogue', ('LDTM',))):
        idx = [i for i, r in enumerate(body) if any(q in r[col['Source']] for q in pat)]
        if idx:
            marks.append((idx[0], tag))
    marks.sort()
    for j, (start, tag) in enumerate(marks):
        lo = 0 if j == 0 else start
        hi = marks[j + 1][0] if j + 1 < len(marks) else len(body)
        n = sum(int(r[col['# Samples']]) for r in body[lo:hi])
        top_r = sorted(reasons, key=lambda k: -sum(int(r[col[k]]) for r in body[lo:hi]))[:3]
        print('code from first %-8s marker, instructions %5d..%5d: %.1f%% of samples (%s)' % (tag, lo, hi, 100.0 * n / tot, ', '.join(k[6:] for k in top_r)))
    ranked = sorted(range(len(body)), key=lambda i: -int(body[i][col['# Samples']]))[:top]
    print('| # | share | executed | main reason | instruction |')
    print('|---|---|---|---|---|')
    for i in sorted(ranked):
        r = body[i]
        n = int(r[col['# Samples']])
        main_r = max(reasons, key=lambda k: int(r[col[k]]))
        print('| %d | %.1f%% | %s | %s | `%s` |' % (i, 100.0 * n / tot, r[col['Instructions Executed']], main_r[6:],
                                                  ' '.join(r[col['Source']].split())))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 25)
