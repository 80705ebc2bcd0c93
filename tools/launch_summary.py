"""Per-kernel time and warp-instruction count of the last bench launch in an ncu launch list captured with
`--metrics gpu__time_duration.sum,smsp__inst_executed.sum`:  python tools/launch_summary.py list.csv [n_last]"""
import csv
import sys


def main(path, n_last=22):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    h = rows[hdr]
    ki, mi, vi = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value')
    byid, order = {}, []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        if r[0] not in byid:
            byid[r[0]] = {'k': r[ki].split('(')[0][:44]}
            order.append(r[0])
        byid[r[0]][r[mi]] = float(r[vi].replace(',', ''))
    tot_t = tot_n = 0.0
    for i in order[-n_last:]:
        d = byid[i]
        t, n = d.get('gpu__time_duration.sum', 0) / 1000, d.get('smsp__inst_executed.sum', 0)
        tot_t += t
        tot_n += n
        print('%-46s %8.1f us  %12.0f warp inst' % (d['k'], t, n))
    print('%-46s %8.1f us  %12.0f warp inst' % ('sum', tot_t, tot_n))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 22)
