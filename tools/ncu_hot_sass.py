"""tools/ncu_hot_sass.py report.ncu-rep [min_exec_frac] -- per-region view of an ncu source page (SASS): groups consecutive
instructions by executed count, prints each region's instruction count, executed instructions, stall samples by reason."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
col = {n: i for i, n in enumerate(h)}
reasons = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    data.append(dict(src=r[col["Source"]].strip(), ex=int(r[col["Instructions Executed"]]), samp=int(r[col["# Samples"]]),
                     st={n: int(r[col[n]]) for n in reasons}))
tot_ex = sum(d["ex"] for d in data); tot_s = sum(d["samp"] for d in data)
print("instructions", len(data), "executed", tot_ex, "samples", tot_s)
# regions = maximal runs with the same executed count (+-2%)
regs, cur = [], []
for d in data:
    if cur and abs(d["ex"] - cur[-1]["ex"]) > 0.02 * max(cur[-1]["ex"], 1):
        regs.append(cur); cur = []
    cur.append(d)
if cur: regs.append(cur)
print("%6s %6s %9s %6s %6s  %s" % ("start", "n", "exec/ins", "ex%", "samp%", "top stall reasons"))
pos = 0
detail = int(sys.argv[2]) if len(sys.argv) > 2 else -1
for g in regs:
    ex = sum(d["ex"] for d in g); sm = sum(d["samp"] for d in g)
    if ex > 0.01 * tot_ex or sm > 0.01 * tot_s:
        st = collections.Counter()
        for d in g:
            for k, v in d["st"].items(): st[k] += v
        ops = collections.Counter((d["src"].split()[1] if d["src"].startswith("@") else d["src"].split()[0]).split(".")[0] for d in g)
        print("%6d %6d %9d %5.1f%% %5.1f%%  %s" % (pos, len(g), g[0]["ex"], 100 * ex / tot_ex, 100 * sm / tot_s,
              ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100 * v / max(sm, 1)) for k, v in st.most_common(5))))
        print("        ", dict(ops.most_common(12)))
        if detail == pos:
            for d in g:
                top = max(d["st"].items(), key=lambda kv: kv[1])
                print("            %6d %-70s %s" % (d["samp"], d["src"][:70], top[0] if top[1] else ""))
    pos += len(g)
