"""Attribute ncu per-SASS-instruction counts to CUDA source lines.

usage: ncu_lines.py <ncu source-page csv (sass)> <cubin> <kernel substring> [top N]
The SASS csv comes from `ncu -i X.ncu-rep --page source --csv --kernel-id ::regex:NAME:1`;
the cubin from `cuobjdump -xelf all lib.so`; line info from `nvdisasm --print-line-info`.
"""
import csv, re, subprocess, sys, collections

sass_csv, cubin, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ins = [(r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]]),
        int(r[ix["# Samples"]] or 0)) for r in rows[2:] if len(r) > 5]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
# find the function section
lines, cur_line, infn, stack = [], None, False, []
for l in dis:
    if l.startswith(".text."):
        infn = kname in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = m.group(3)
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)), inl.strip())
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((cur_line, m.group(2).strip()))
print(f"sass rows in ncu: {len(ins)}, in cubin: {len(lines)}")
n = min(len(ins), len(lines))
tot = sum(i[1] for i in ins)
by_line = collections.Counter()
by_line_s = collections.Counter()
for k in range(n):
    key = lines[k][0][:2] if lines[k][0] else ("?", 0)
    by_line[key] += ins[k][1]
    by_line_s[key] += ins[k][3]
src_cache = {}
def src(f, ln):
    if f not in src_cache:
        try:
            src_cache[f] = open("/root/repo/sdf-viewer_b200/csrc/" + f.replace("sdfgpu_fill_jit.cu","fill_device.cuh")).read().splitlines()
        except Exception:
            src_cache[f] = []
    s = src_cache[f]
    return s[ln - 1].strip()[:110] if 0 < ln <= len(s) else ""
print(f"total warp instructions: {tot}")
for (f, ln), c in by_line.most_common(topn):
    print(f"{100*c/tot:6.2f}%  {c:>12}  samp {by_line_s[(f, ln)]:>6}  {f}:{ln}  {src(f, ln)}")
