"""Aggregates an ncu source-page SASS csv by CUDA source line, using nvdisasm line info of the matching cubin.
    python tools/ncu_lines.py report.ncu-rep cubin kernel_substring [min_percent]"""
import csv, re, subprocess, sys, io, collections
rep, cubin, kern = sys.argv[1:4]
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(r for r in rows if '# Samples' in r)
data = rows[rows.index(hdr) + 1:]
iS, iE, iSrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
# split per function
funcs = re.split(r'\n\s*\.section\s+\.text\.', dis)
body = next(f for f in funcs if kern in f.split('\n', 1)[0] and ('Id' in f.split('\n', 1)[0] if 'double' in kern else True)) if False else None
cands = [f for f in funcs if kern in f.split('\n', 1)[0]]
want = len(data)
best = None
for f in cands:
    lines, cur = [], None
    for ln in f.split('\n'):
        mm = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
        if mm:
            cur = (mm.group(1), int(mm.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
            lines.append(cur)
    if best is None or abs(len(lines) - want) < abs(len(best) - want):
        best = lines
print("sass rows in report %d, in cubin %d" % (want, len(best)))
agg = collections.Counter(); aggE = collections.Counter()
for r, loc in zip(data, best):
    agg[loc] += int(r[iS] or 0); aggE[loc] += int(r[iE] or 0)
tot = sum(agg.values())
src = {}
for loc, n in agg.most_common():
    if n < tot * minpct / 100: break
    fn, ln = loc if loc else ('?', 0)
    if fn not in src:
        try: src[fn] = open('/root/repo/nemo_b200/csrc/' + fn).read().split('\n')
        except Exception: src[fn] = []
    text = src[fn][ln - 1].strip()[:100] if 0 < ln <= len(src[fn]) else ''
    print("%5.1f%% samples  inst %11d  %s:%d  %s" % (100.0 * n / tot, aggE[loc], fn, ln, text))
