"""Prints the launches of the last training step found in an `ncu --metrics gpu__time_duration.sum --csv` log.
usage: python profiles/parse_launches.py gpurun_out/launches.csv"""
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') == 'gpu__time_duration.sum':
            v = float(row['Metric Value'].replace(',', ''))
            u = row['Metric Unit']
            v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
            rows.append((int(row['ID']), row['Kernel Name'].split('(')[0].replace('void ', '').replace('mpqe::<unnamed>::', '')
                         .replace('unnamed>::', ''), v,
                         row['Grid Size']))
    return rows


def main(path):
    rows = load(path)
    # a training step starts at the ids-only gather (or the forward gather) and ends before the next one / the eval
    starts = [i for i, r in enumerate(rows) if r[1].startswith('gather_ids_multi')] or \
             [i for i, r in enumerate(rows) if r[1].startswith('gather_fwd_multi') and r[3] != '(1536, 1, 1)']
    a = starts[-1]
    b = next((i for i in range(a + 1, len(rows)) if rows[i][1].startswith(('gather_ids_multi', 'cosine_scores'))),
             len(rows))
    while b > a and rows[b - 1][1].startswith(('pack_weights', 'gather_fwd_multi', 'layer_', 'transpose', 'matrix_sum')) and \
            rows[b - 1][0] > rows[a][0] + 40:
        b -= 1     # forward launches of the evaluation that follows the last step
    tot = 0.0
    agg = {}
    for r in rows[a:b]:
        print('%5d %-46s %8.1f us  grid %s' % (r[0], r[1][:46], r[2], r[3]))
        tot += r[2]
        agg[r[1][:46]] = agg.get(r[1][:46], 0.0) + r[2]
    print('launches %d, sum of durations %.1f us (cold caches, serialised)' % (b - a, tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        print('  %-46s %8.1f us  %5.1f%%' % (k, v, 100 * v / tot))


if __name__ == '__main__':
    main(sys.argv[1])
