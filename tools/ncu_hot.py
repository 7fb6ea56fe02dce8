"""Top stall-sampled SASS instructions of an .ncu-rep source page."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for si, s in enumerate(start[:1]):
    hdr = rows[s]
    end = start[si + 1] - 1 if si + 1 < len(start) else len(rows)
    body = [r for r in rows[s + 1:end] if len(r) == len(hdr)]
    ci = hdr.index("# Samples"); src = hdr.index("Source"); ex = hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(float(r[ci]) for r in body)
    print("total samples", tot, " instructions", len(body))
    ranked = sorted(enumerate(body), key=lambda t: -float(t[1][ci]))[:top]
    for idx, r in ranked:
        reasons = sorted(((float(r[i]), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        rs = " ".join(f"{n}={v:.0f}" for v, n in reasons if v > 0)
        print(f"{idx:5d} {float(r[ci]):7.0f} {100 * float(r[ci]) / tot:5.1f}%  exec={r[ex]:>8s}  {r[src].strip()[:70]:70s} {rs}")
