"""List the hottest source lines of each kernel in an ncu report (development aid).
usage: python tools/ncu_hot.py report.ncu-rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kern = None; hdr = None; lines = []; seen = set()
def flush():
    if not lines: return
    tot = sum(x[1] for x in lines) or 1
    print("==== %s   (samples %d)" % (kern[:100], tot))
    for ln, smp, inst, src, stalls in sorted(lines, key=lambda x: -x[1])[:top]:
        print("%5s %6.2f%% inst %11d  %-70s %s" % (ln, 100.0 * smp / tot, inst, src[:70], stalls))
for r in rows:
    if not r: continue
    if r[0] == "Function Name":
        flush(); kern = r[1]; lines = []; hdr = None; seen = set(); continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or r[0] == "" or not r[0].isdigit(): continue
    if r[0] in seen: continue
    seen.add(r[0])
    d = dict(zip(hdr[4:], r[4:]))
    try: smp = int(d.get("# Samples", "0")); inst = int(d.get("Instructions Executed", "0"))
    except ValueError: continue
    st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
    stalls = " ".join("%s:%d" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    lines.append((r[0], smp, inst, r[1].strip(), stalls))
flush()
