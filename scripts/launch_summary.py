"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one training step, per-kernel shares."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
seq = []
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    if r[ui] == 'ns':
        v /= 1000
    seq.append((r[ki], v))
idx = [i for i, s in enumerate(seq) if 'start_fwd' in s[0]]
# the last complete step in the capture (the first ones carry one-time optimizer-state fills)
a, b = (idx[-2], idx[-1]) if len(idx) > 1 else (idx[0], len(seq))
step = seq[a:b]
tot = sum(s[1] for s in step)
print(f'one step: {len(step)} launches, {tot:.0f} us of (cold-cache, serialised) kernel time')
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in step:
    k = re.sub(r'void ', '', k)
    k = re.sub(r'\(.*', '', k)[:90]
    agg[k][0] += 1
    agg[k][1] += v
print('| share | total us | launches | kernel |\n|---|---|---|---|')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f'| {100 * t / tot:.1f}% | {t:.0f} | {n} | `{k}` |')
