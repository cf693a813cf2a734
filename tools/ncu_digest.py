"""Digest of an .ncu-rep: key metrics per kernel + instruction/stall share by source line (needs -lineinfo)."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'sm__cycles_elapsed.max', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum']
seen = set()
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')][:40]
    if name in seen: continue
    seen.add(name)
    print('---', name)
    for w in want:
        if w in hdr: print(f'  {w:72s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur_file = None; h = None; kern = None
agg = collections.defaultdict(lambda: collections.OrderedDict())
def f(x):
    try: return float(x)
    except: return 0.0
for r in rows:
    if r and r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r and r[0] == 'Function Name': kern = r[1][:30]; continue
    if r and r[0] == 'Line No': h = r; continue
    if h and len(r) == len(h) and r[0]:
        try: ln = int(r[0])
        except: continue
        a = agg[kern].setdefault((cur_file, ln), [r[1][:95], 0, 0, 0])
        a[1] += f(r[4]); a[2] += f(r[7]); a[3] += f(r[8])
for kern, d in agg.items():
    ts = sum(a[1] for a in d.values()) or 1; ti = sum(a[2] for a in d.values()) or 1
    print(f"=== {kern}: samples {ts:.0f} inst {ti:.0f}")
    for (fl, ln), a in sorted(d.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{fl[:16]:16s} {ln:4d} samp {100*a[1]/ts:5.1f}% inst {100*a[2]/ti:5.1f}% thr/inst {a[3]/max(a[2],1):5.1f} | {a[0]}")
