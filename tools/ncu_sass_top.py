"""Top SASS instructions by stall samples from an exported `ncu --page source --csv` file (one block per captured launch).
    python tools/ncu_sass_top.py source.csv kernel_substring [block_index] [top_n]"""
import csv, sys
path, kern = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
rows = list(csv.reader(open(path, errors='ignore')))
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
sel = [b for b in blocks if kern in b['name']]
b = sel[which]
hdr = b['rows'][0]
data = [r for r in b['rows'][1:] if len(r) == len(hdr)]
iS, iSrc, iE = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
tot = sum(int(r[iS] or 0) for r in data)
print(b['name'][:150])
print("instructions %d, samples %d, warp-instructions executed %d" % (len(data), tot, sum(int(r[iE] or 0) for r in data)))
# distribution along the program
n = len(data)
acc = 0
marks = []
for i, r in enumerate(data):
    acc += int(r[iS] or 0)
    if (i + 1) % max(1, n // 20) == 0:
        marks.append("%d:%.0f%%" % (i + 1, 100.0 * acc / tot))
print("cumulative samples by instruction index: " + " ".join(marks))
order = sorted(range(n), key=lambda i: -int(data[i][iS] or 0))[:topn]
for i in sorted(order):
    print("%5d %5.1f%%  %s" % (i, 100.0 * int(data[i][iS] or 0) / tot, data[i][iSrc].strip()[:90]))
