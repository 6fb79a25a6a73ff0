"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count / total / average (ms)."""
import collections, csv, re, sys

def summarise(path, top=30):
    lines = [l for l in open(path) if l.startswith('"')]
    tot = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        t = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        t *= {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3, 'second': 1e3}[unit]
        name = re.sub(r'\(.*', '', row['Kernel Name'])[:64]
        d = tot.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += t
    total = sum(v[1] for v in tot.values())
    out = []
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f"{k:66s} n={v[0]:5d} total={v[1]:10.3f} ms avg={v[1] / v[0]:9.4f} ms share={100 * v[1] / total:5.1f}%")
    out.append(f"{'TOTAL':66s} n={sum(v[0] for v in tot.values()):5d} total={total:10.3f} ms")
    return "\n".join(out)

if __name__ == "__main__":
    print(summarise(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30))
