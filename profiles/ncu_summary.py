#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel."""
import collections, csv, re, sys
def main(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg, tot = collections.OrderedDict(), 0.0
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', r['Kernel Name'])
        name = re.sub(r'^void (mvus::)?', '', name)
        v = float(r['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3}.get(r['Metric Unit'], 1.0)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    print('%-48s %5s %12s %10s %7s' % ('kernel', 'n', 'total_us', 'avg_us', 'share'))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-48s %5d %12.1f %10.1f %6.1f%%' % (k[:48], c, t, t / c, 100 * t / tot))
    print('%-48s %5s %12.1f' % ('TOTAL', '', tot))
if __name__ == '__main__':
    main(sys.argv[1])
