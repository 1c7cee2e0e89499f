"""Summarise an ncu report's source page by CUDA source line (top lines by warp-stall samples).
usage: ncu_lines.py report.ncu-rep [top]      (needs -lineinfo at compile time and --import-source on at capture)"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = None; hdr = None; acc = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": cur_fn = r[1][:60]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "-"):
        i_s = hdr.index("# Samples"); i_e = hdr.index("Instructions Executed")
        try: s = int(r[i_s]); e = int(r[i_e])
        except ValueError: continue
        key = (cur_fn, fname, r[0])
        a = acc.setdefault(key, [0, 0, r[1].strip()[:100]])
        a[0] += s; a[1] += e
for fn in sorted(set(k[0] for k in acc)):
    sub = {k: v for k, v in acc.items() if k[0] == fn}
    tot = sum(v[0] for v in sub.values()) or 1
    print("==", fn, "total samples", tot, "warp-instructions", sum(v[1] for v in sub.values()))
    for k, v in sorted(sub.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*v[0]/tot:5.1f}% {k[1]}:{k[2]:>5} inst={v[1]:>9}  {v[2]}")
