"""Sum an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/launch_summary.py launches.csv [first_launch_id last_launch_id]"""
import collections, csv, re, sys

def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    hdr = next(r for r in rows if 'Kernel Name' in r)
    rows = rows[rows.index(hdr) + 1:]
    kn, mv, mu, idc = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('ID')
    lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 1 << 60)
    agg = collections.OrderedDict()
    for r in rows:
        if not (lo <= int(r[idc]) <= hi):
            continue
        name = re.sub(r'\(.*', '', r[kn])
        v = float(r[mv].replace(',', ''))
        v *= {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0}[r[mu]]
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f'{sys.argv[1]}: {sum(a[0] for a in agg.values())} launches, {tot:.3f} ms')
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f'  {k[:70]:70s} n={a[0]:5d} {a[1]:9.3f} ms {100 * a[1] / tot:5.1f} %')

main()
