"""Aggregate the ncu source page (SASS) of a kernel into hot blocks: python scripts/hot_blocks.py rep [min_share]"""
import csv, subprocess, sys
rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
print(rows[0][1][:100])
rows = rows[2:]
tot = sum(int(r[5]) for r in rows)
warps = int(rows[0][5])
print('total warp inst', tot, 'per warp', round(tot / warps, 1))
blocks, cur = [], None
for idx, r in enumerate(rows):
    n, samples = int(r[5]), int(r[4])
    op = ' '.join(r[1].split()[:2]) if r[1].split() and r[1].split()[0].startswith('@') else (r[1].split()[0] if r[1].split() else '')
    if cur and abs(n - cur['n']) <= 0.15 * max(n, cur['n'], 1):
        cur['sum'] += n; cur['cnt'] += 1; cur['end'] = idx; cur['samples'] += samples; cur['ops'].append(op)
    else:
        cur = {'n': n, 'sum': n, 'cnt': 1, 'start': idx, 'end': idx, 'samples': samples, 'ops': [op]}
        blocks.append(cur)
allsamp = max(1, sum(b['samples'] for b in blocks))
for b in blocks:
    if b['sum'] > min_share / 100 * tot:
        print(f"{b['start']:5d}-{b['end']:5d} n/warp={b['n']/warps:8.1f} instrs={b['cnt']:4d} share={100*b['sum']/tot:5.1f}% stalls={100*b['samples']/allsamp:5.1f}%  {' '.join(b['ops'][:12])}")
