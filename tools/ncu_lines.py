"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per CUDA source line:
samples, dominant stall reasons.  usage: ncu_lines.py file.csv [kernel_index] [top]"""
import csv, sys, collections
path = sys.argv[1]; kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "File Path"]
starts.append(len(rows))
# group consecutive file sections into kernels: a kernel starts at "Function Name" following File Path
secs = []
for a, b in zip(starts[:-1], starts[1:]):
    secs.append((rows[a][1], rows[a + 1][1] if rows[a + 1][0] == "Function Name" else "", rows[a + 2], rows[a + 3:b]))
kernels = []
for s in secs:
    if not kernels or kernels[-1][0] != s[1]:
        kernels.append((s[1], []))
    kernels[-1][1].append(s)
# kernels may repeat per launch; pick by index over (function name) occurrences
name, ss = kernels[kidx]
print("kernel:", name[:80], "sections:", [s[0].split('/')[-1] for s in ss])
per_line = collections.defaultdict(lambda: collections.Counter())
src_text = {}
tot = 0
for fpath, fn, hdr, body in ss:
    f = fpath.split('/')[-1]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    si = hdr.index("# Samples")
    cur = None
    for r in body:
        if len(r) < 3: continue
        if r[0] != "":
            cur = (f, int(r[0])); src_text[cur] = r[1]
        if len(r) > si and r[2] != "":
            try: n = int(r[si])
            except ValueError: continue
            per_line[cur]["samples"] += n; tot += n
            for h, i in stall_cols:
                try: per_line[cur][h] += int(r[i])
                except (ValueError, IndexError): pass
print("total samples", tot)
nb = sum(c["samples"] - c["stall_barrier"] for c in per_line.values())
print("non-barrier samples", nb)
order = sorted(per_line.items(), key=lambda kv: -(kv[1]["samples"] - kv[1]["stall_barrier"]))
for (f, ln), c in order[:top]:
    s = c["samples"]; x = s - c["stall_barrier"]
    reasons = ", ".join(f"{h[6:]}={v}" for h, v in c.most_common(6) if h.startswith("stall_") and v > 0 and h != "stall_barrier")
    print(f"{f}:{ln:4d} nonbar={x:6d} ({100*x/max(nb,1):4.1f}%) bar={c['stall_barrier']:6d} | {reasons} | {src_text[(f,ln)].strip()[:90]}")
