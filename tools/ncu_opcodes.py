"""tools/ncu_opcodes.py report.ncu-rep [n-iterations] -- executed SASS opcode histogram of a captured kernel."""
import csv, collections, io, subprocess, sys
rep = sys.argv[1]
niter = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if 'Source' in r and any(c == 'Instructions Executed' for c in r)][0]
h = rows[hi]; si, ei = h.index('Source'), h.index('Instructions Executed')
cnt, tot = collections.Counter(), 0
for r in rows[hi + 1:]:
    try:
        n = int(r[ei])
    except Exception:
        continue
    t = r[si].split()
    if not t:
        continue
    op = t[1] if t[0].startswith('@') else t[0]
    cnt[op.split('.')[0]] += n
    tot += n
print("total warp-inst %d, per iteration %.1f" % (tot, tot / niter))
for op, n in cnt.most_common(28):
    print("  %-10s %12d  %7.1f/iter  %5.1f%%" % (op, n, n / niter, 100 * n / tot))
