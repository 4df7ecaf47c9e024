"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the phases of one
flow+warp step (feature encoder / context encoder + setup / per-iteration / tail).
    python tools/summarize_launches.py gpurun_out/launches_step.csv [title]"""
import collections
import csv
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = []
    for row in csv.DictReader(lines):
        try:
            v = float(row['Metric Value'])
        except (KeyError, ValueError):
            continue
        u = row.get('Metric Unit', 'ns')
        rows.append((row['Kernel Name'], v / 1000.0 if u in ('ns', 'nsecond') else v))
    return rows


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = load(path)
    total = sum(t for _, t in rows)
    print(f'# {title}')
    print(f'launches {len(rows)}   sum of kernel durations {total:.1f} us\n')
    agg = collections.OrderedDict()
    for n, t in rows:
        a = agg.setdefault(n, [0.0, 0])
        a[0] += t
        a[1] += 1
    for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if t / total < 0.0005:
            continue
        print(f'{t:9.1f} us {100 * t / total:5.1f}%  x{c:4d}  {n[:100]}')
    ours = sum(t for n, t in rows if 'sdof::' in n)
    print(f'\nhand-written (sdof::) kernels: {ours:.1f} us = {100 * ours / total:.1f}% of the step')
    names = [n for n, _ in rows]
    looks = [i for i, n in enumerate(names) if 'corr_lookup' in n]
    pyr = [i for i, n in enumerate(names) if 'corr_pyramid' in n or 'corr_volume_tc' in n]
    if looks and pyr:
        tot = lambda a, b: sum(t for _, t in rows[a:b])
        print(f'\nphases (launch order; side-stream kernels interleave): up to the pyramid kernel {tot(0, pyr[0] + 1):.1f} us in {pyr[0] + 1} launches; '
              f'pyramid..first lookup {tot(pyr[0] + 1, looks[0]):.1f} us; ')
        if len(looks) > 6:
            print(f'one GRU iteration (6th) {tot(looks[5], looks[6]):.1f} us in {looks[6] - looks[5]} launches:')
            for i in range(looks[5], looks[6]):
                print(f'    {rows[i][1]:7.1f}  {rows[i][0][:105]}')
        print(f'loop + tail {tot(looks[0], len(rows)):.1f} us')


if __name__ == '__main__':
    main()
