"""Development aid: SASS-level view of one kernel of an `ncu --set full` report: opcode histogram weighted by executed
warp instructions, and the instruction stream with per-instruction counts written to a text file for reading.
usage: python tools/ncu_sass_hot.py REPORT kernel-substring OUT.txt"""
import collections
import csv
import subprocess
import sys

rep, want, outp = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
fn, hdr, out = None, None, []
for r in csv.reader(raw.splitlines()):
    if not r:
        continue
    if r[0] == "Kernel Name":
        if fn is not None and want in fn and out:
            break                       # the page lists every kernel twice: keep the first listing
        fn = r[1]; hdr = None; continue
    if r[0] == "Address":
        hdr = r; ci = r.index("Instructions Executed"); cs = r.index("# Samples"); continue
    if hdr is None or fn is None or want not in fn:
        continue
    try:
        out.append((r[0], r[1].strip(), int(r[ci]), int(r[cs])))
    except ValueError:
        pass
tot = sum(o[2] for o in out); tots = sum(o[3] for o in out)
print(f"{len(out)} SASS instructions, {tot/1e9:.3f} G executed, {tots} samples")
h = collections.Counter()
for a, s, n, sm in out:
    op = s.split(); name = op[1] if op[0].startswith("@") else op[0]
    h[name.split(".")[0]] += n
for k, v in h.most_common(36):
    print("%-10s %6.2f%%" % (k, 100 * v / tot))
with open(outp, "w") as f:
    for k, (a, s, n, sm) in enumerate(out):
        f.write("%5d %s %9.3fM %6.2f%% samp %5.2f%%  %s\n" % (k, a[-5:], n / 1e6, 100 * n / tot, 100 * sm / max(tots, 1), s))
