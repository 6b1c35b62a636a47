#!/usr/bin/env python
"""Per-kernel table from an `ncu --page raw --csv` dump of one forward: launches, time, tensor-pipe %, DRAM bytes and
throughput, achieved occupancy.  usage: python scripts/ncu_table.py raw.csv > profiles/rNN_forward_ncu_table.txt"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
data = [r for r in rows[2:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}


def f(r, name):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return float('nan')
    return float(r[i].replace(',', ''))


units = dict(zip(hdr, rows[1]))
agg = collections.OrderedDict()
for r in data:
    name = re.sub(r'\(.*', '', r[col['Kernel Name']]).replace('void ', '').replace('dclip::', '')
    t = f(r, 'gpu__time_duration.sum')
    t = t * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(units.get('gpu__time_duration.sum', 'ns'), 1e-3)   # -> us
    a = agg.setdefault(name, dict(n=0, t=0.0, tensor=[], dr=0.0, dw=0.0, dram=[], occ=[], regs=0, xu=[], issue=[]))
    a['n'] += 1
    a['t'] += t
    a['tensor'].append(f(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'))
    mul = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    a['dr'] += f(r, 'dram__bytes_read.sum') * mul.get(units.get('dram__bytes_read.sum', 'byte'), 1)
    a['dw'] += f(r, 'dram__bytes_write.sum') * mul.get(units.get('dram__bytes_write.sum', 'byte'), 1)
    a['dram'].append(f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'))
    a['occ'].append(f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'))
    a['xu'].append(f(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'))
    a['issue'].append(f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'))
    a['regs'] = int(f(r, 'launch__registers_per_thread'))
tot = sum(a['t'] for a in agg.values())
mean = lambda v: sum(v) / max(len(v), 1)  # noqa: E731
print(f"{len(data)} launches, {tot / 1000:.3f} ms under ncu (serialised, replayed, cold caches: compare SHARES)")
print(f"{'ms':>8s} {'share':>6s} {'n':>4s} {'us/launch':>9s} {'tensor%':>7s} {'xu%':>5s} {'issue%':>6s} {'dramRd MB':>9s} {'dramWr MB':>9s} {'GB/s':>7s} {'dram%':>5s} {'occ%':>5s} {'regs':>4s}  kernel")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['t']):
    n = a['n']
    gbs = (a['dr'] + a['dw']) / (a['t'] * 1e-6) / 1e9 if a['t'] else 0
    print(f"{a['t'] / 1000:8.3f} {100 * a['t'] / tot:5.1f}% {n:4d} {a['t'] / n:9.1f} {mean(a['tensor']):7.1f} {mean(a['xu']):5.1f} {mean(a['issue']):6.1f} "
          f"{a['dr'] / n / 1e6:9.1f} {a['dw'] / n / 1e6:9.1f} {gbs:7.0f} {mean(a['dram']):5.1f} {mean(a['occ']):5.1f} {a['regs']:4d}  {k[:100]}")
