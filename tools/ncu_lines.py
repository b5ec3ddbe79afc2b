"""tools/ncu_lines.py report.ncu-rep source-file-substring [n-solves n-iterations] -- executed warp-instructions per
source function / line from an ncu capture taken with --import-source on (-lineinfo build)."""
import csv, collections, io, re, subprocess, sys
rep, fsub = sys.argv[1], sys.argv[2]
nsolve = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
niter = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
his = [i for i, r in enumerate(rows) if 'Source' in r and any(c == 'Instructions Executed' for c in r)]
grand = 0
sections = []
for n, s in enumerate(his):
    e = his[n + 1] if n + 1 < len(his) else len(rows)
    h = rows[s]; li, si, ei = h.index('Line No'), h.index('Source'), h.index('Instructions Executed')
    path = rows[s - 2][1] if s >= 2 else '?'
    lines = []
    for r in rows[s + 1:e]:
        try:
            lines.append((int(r[li]), int(r[ei]), r[si]))
        except Exception:
            pass
    tot = sum(x[1] for x in lines)
    grand += tot
    sections.append((path, tot, lines))
print("total warp-instructions %d  (%.0f per solve)" % (grand, grand / nsolve))
for path, tot, lines in sections:
    if tot > 0.003 * grand:
        print("  %-70s %6.1f%%  %8.0f/solve" % (path[-70:], 100 * tot / grand, tot / nsolve))
sec = [x for x in sections if fsub in x[0]][0]
src = open(sec[0]).read().splitlines()
owner, cur = {}, '?'
for i, l in enumerate(src, 1):
    if l.startswith('__device__') or (l.startswith('template') and '(' in l):
        m = re.search(r'(\w+)\(', l)
        if m and 'struct' not in l:
            cur = m.group(1)
    elif re.match(r'^[a-z].*\b(\w+)\(.*', l) and not l.startswith(('//', 'namespace', 'constexpr', 'struct', 'using')):
        m = re.search(r'(\w+)\(', l)
        if m:
            cur = m.group(1)
    owner[i] = cur
agg = collections.Counter()
for ln, n_, s in sec[2]:
    agg[owner.get(ln, '?')] += n_
print("by function in", fsub)
for k, v in agg.most_common():
    print("  %-24s %6.1f%%  %8.0f/solve  %7.1f/iter" % (k, 100 * v / grand, v / nsolve, v / niter))
if len(sys.argv) > 5:
    print("hottest lines")
    for ln, n_, s in sorted(sec[2], key=lambda x: -x[1])[:int(sys.argv[5])]:
        print("  %5d %9d %6.1f/iter  %s" % (ln, n_, n_ / niter, s.strip()[:100]))
