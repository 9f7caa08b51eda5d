"""Top stalled SASS instructions of an ncu report (source page)."""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; body = rows[2:]
si = hdr.index('# Samples'); src = hdr.index('Source')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[si]) for r in body if r[si].isdigit())
print("total samples", tot)
idx = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:n]
for i in sorted(idx):
    r = body[i]
    st = sorted(((int(r[c] or 0), hdr[c]) for c in stall_cols), reverse=True)[:2]
    print("%5d %5.1f%%  %-70s %s" % (i, 100.0 * int(r[si]) / tot, r[src].strip()[:70], st))
