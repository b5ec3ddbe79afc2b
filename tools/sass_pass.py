"""tools/sass_pass.py <object.o> <kernel-substring> -- instruction mix between the PMTRIG markers of a kernel compiled with
-DMPC_QUAD_MARK (admm_quad.cuh: marker 1 = pass begins, marker 2 = pass ends): what ONE ADMM pass costs, spills included."""
import collections, os, re, subprocess, sys, tempfile
obj, sub = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
L = subprocess.run(["nvdisasm", "-c", cub], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(L) if l.startswith(".text.") and sub in l and l.rstrip().endswith(":")][0]
ins = []
for l in L[start + 1:]:
    if l.startswith("//--------------------- .text."):
        break
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        ins.append(m.group(2))
marks = [i for i, t in enumerate(ins) if "PMTRIG" in t]
print("markers at", marks, "of", len(ins), "instructions")
def opc(t):
    p = t.split()
    return (p[1] if p[0].startswith("@") else p[0])
for a, b in zip(marks, marks[1:]):
    seg = ins[a + 1:b]
    ops = collections.Counter(opc(t).split(".")[0] + (".128" if ".128" in opc(t) else (".64" if ".64" in opc(t) else "")) for t in seg)
    print("segment %d..%d: %d instructions" % (a, b, len(seg)))
    print("  ", dict(ops.most_common(30)))
