"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the last step."""
import csv
import sys


def main(path, tail=40):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    h = rows[hdr]
    ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    seq = []
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            v = float(r[vi].replace(',', ''))
            if r[ui] in ('ns', 'nsecond'):
                v /= 1000.0
            seq.append((r[ki].split('(')[0][:44], v))
    for name, v in seq[-tail:]:
        print('%-46s %10.1f us' % (name, v))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
