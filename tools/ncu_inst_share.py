import csv, subprocess, sys, collections, io
rep = sys.argv[1]; top = int(sys.argv[2]); kfilter = sys.argv[3] if len(sys.argv) > 3 else ""
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
        a = agg[kern].setdefault((cur_file, ln), [r[1][:110], 0, 0, 0])
        a[1] += f(r[4]); a[2] += f(r[7]); a[3] += f(r[8])
for kern, d in agg.items():
    if kfilter and kfilter not in kern: continue
    ts = sum(a[1] for a in d.values()) or 1; ti = sum(a[2] for a in d.values()) or 1
    print(f"=== {kern}: samples {ts:.0f} inst {ti:.0f}")
    cum = 0
    for (fl, ln), a in sorted(d.items(), key=lambda kv: -kv[1][2])[:top]:
        cum += a[2]
        print(f"{fl[:16]:16s} {ln:4d} inst {100*a[2]/ti:5.1f}% cum {100*cum/ti:5.1f}% samp {100*a[1]/ts:5.1f}% thr/inst {a[3]/max(a[2],1):5.1f} | {a[0]}")
