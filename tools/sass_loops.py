"""tools/sass_loops.py <object.o> <kernel-name-substring> -- instruction mix of the loops of one kernel (nvdisasm)."""
import collections, os, re, subprocess, sys, tempfile
obj, sub = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 100
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
L = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.split("\n")
starts = [i for i, l in enumerate(L) if l.startswith(".text.") and sub in l and l.rstrip().endswith(":")]
for start in starts:
    end = start + 1
    while end < len(L) and not L[end].startswith("//--------------------- .text."):
        end += 1
    cur, rows = None, []
    for l in L[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            rows.append((int(m.group(1), 16), m.group(2), cur))
        elif re.match(r"\.L_x_\d+:", l):
            rows.append((None, l.strip(), cur))
    print(L[start][:120], len([r for r in rows if r[0] is not None]), "instructions")
    labels = {t[:-1]: i for i, (a, t, c) in enumerate(rows) if a is None}
    back = []
    for i, (a, t, c) in enumerate(rows):
        m = re.search(r"BRA.*(\.L_x_\d+)", t)
        if m and m.group(1) in labels and labels[m.group(1)] < i:
            back.append((labels[m.group(1)], i))
    seen = []
    for s_, e_ in sorted(back, key=lambda p: -(p[1] - p[0])):
        if any(abs(s_ - a) < 40 and abs(e_ - b) < 200 for a, b in seen):
            continue
        seen.append((s_, e_))
        seg = [r for r in rows[s_:e_ + 1] if r[0] is not None]
        if len(seg) < minlen:
            continue
        def opc(t):
            p = t.split()
            return p[1] if p[0].startswith("@") else p[0]
        ops = collections.Counter(opc(r[1]).split(".")[0] for r in seg)
        print("  loop rows %d..%d: %d instr, src %s .. %s" % (s_, e_, len(seg), seg[0][2], seg[-1][2]))
        print("    ", dict(ops.most_common(24)))
        bysrc = collections.Counter(r[2] for r in seg if "LDL" in r[1] or "STL" in r[1])
        if bysrc:
            print("     spills by source line:", dict(bysrc.most_common(12)))
