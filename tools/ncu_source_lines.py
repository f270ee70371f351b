"""Development aid: per-source-line share of the executed warp instructions (and stall samples) of one kernel from an
`ncu --set full --import-source on` report.  usage: python tools/ncu_source_lines.py REPORT [kernel-substring] [top-n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else "(int)4, (int)1, (int)1"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(raw.splitlines()))
lines, sass = {}, {}
fpath, fn, hdr = None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        fn = r[1]; hdr = None; continue
    if r[0] == "Line No":
        hdr = r; ci = r.index("Instructions Executed"); cs = r.index("# Samples"); continue
    if hdr is None or want not in fn:
        continue
    if r[0] != "":
        try:
            key = (fpath, int(r[0]), r[1].strip())
            lines[key] = lines.get(key, (0, 0))
            lines[key] = (lines[key][0] + int(r[ci]), lines[key][1] + int(r[cs]))
            cur = key
        except ValueError:
            pass
    else:
        op = r[3].split()
        if op:
            name = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
            name = name.split(".")[0]
            try:
                sass[name] = sass.get(name, 0) + int(r[ci])
            except ValueError:
                pass
tot = sum(v[0] for v in lines.values()); tots = sum(v[1] for v in lines.values())
print(f"kernel ~ {want}: {tot/1e9:.3f} G warp instructions, {tots} samples")
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0][:22]:22s}:{k[1]:5d} {100*v[0]/tot:6.2f}% inst {100*v[1]/max(tots,1):6.2f}% samp  {k[2][:100]}")
print("-- SASS opcodes")
st = sum(sass.values())
for k, v in sorted(sass.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{k:12s} {100*v/st:6.2f}%")
