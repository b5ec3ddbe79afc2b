"""tools/ncu_traffic.py report.ncu-rep kernel-regex precision batch -- dram__bytes_read.sum + dram__bytes_write.sum of one launch
from an `ncu --set full` capture, merged into profiles/solve_kernel_traffic.json (what bench.py reports as roofline.traffic,
labelled static: ncu cannot run inside a timed bench)."""
import csv, io, json, os, re, subprocess, sys
rep, rx, prec, batch = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if not re.search(rx, d.get("Kernel Name", "")):
        continue
    tot = sum(to_bytes(d[k], units[hdr.index(k)]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    p = os.path.join(REPO, "profiles", "solve_kernel_traffic.json")
    j = json.load(open(p)) if os.path.exists(p) else {}
    j.setdefault(prec, {})[str(batch)] = {"dram_bytes": tot, "kernel": d["Kernel Name"][:60],
                                          "source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full --clock-control none (%s)" % os.path.basename(rep)}
    json.dump(j, open(p, "w"), indent=1)
    print(prec, batch, tot)
    break
