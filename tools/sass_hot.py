"""tools/sass_hot.py <object.o> <kernel-substring> <first-src-line> <last-src-line> [file] -- opcode mix of the SASS
instructions whose -lineinfo source line falls in [first, last] of `file` (default admm_pair.cuh) plus everything
inlined between two such instructions inside the kernel's biggest loop (= the straight-line hot path)."""
import collections, os, re, subprocess, sys, tempfile
obj, sub = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
L = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(L) if l.startswith(".text.") and sub in l and l.rstrip().endswith(":")][0]
end = start + 1
while end < len(L) and not L[end].startswith("//--------------------- .text."):
    end += 1
rows, stack = [], []
for l in L[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        if m.group(3):
            stack.append((m.group(1).split("/")[-1], int(m.group(2)), m.group(3).split("/")[-1], int(m.group(4))))
        else:
            stack = [(m.group(1).split("/")[-1], int(m.group(2)), None, None)]
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        rows.append((m.group(2), list(stack)))
        stack = stack[:0] if False else stack
    elif re.match(r"\.L_x_\d+:", l):
        rows.append((l.strip(), None))
# print block structure: split at labels / branches
blocks, cur = [], []
for t, st in rows:
    if st is None:
        if cur: blocks.append(cur)
        cur = [(t, st)]
    else:
        cur.append((t, st))
        p = t.split()
        op = p[1] if p[0].startswith("@") else p[0]
        if op.startswith(("BRA", "EXIT", "RET", "BRX")):
            blocks.append(cur); cur = []
if cur: blocks.append(cur)
def opc(t):
    p = t.split()
    return (p[1] if p[0].startswith("@") else p[0]).split(".")[0]
minb = int(sys.argv[3]) if len(sys.argv) > 3 else 60
for bi, b in enumerate(blocks):
    ins = [x for x in b if x[1] is not None]
    if len(ins) < minb:
        continue
    ops = collections.Counter(opc(t) for t, _ in ins)
    lines = collections.Counter()
    for t, st in ins:
        if st:
            # outermost frame (the line in the caller) is the last "inlined at", else the line itself
            f = st[-1]
            lines[(f[2], f[3]) if f[2] else (f[0], f[1])] += 1
    lab = b[0][0] if b[0][1] is None else ""
    top = sorted(lines.items(), key=lambda kv: -kv[1])[:6]
    print("block %d %s: %d instr  %s" % (bi, lab, len(ins), dict(ops.most_common(16))))
    print("      lines:", top)
