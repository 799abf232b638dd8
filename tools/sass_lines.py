#!/usr/bin/env python
"""Join an .ncu-rep's per-instruction counters with the line table of the cubin (read on the CPU box).

    python tools/sass_lines.py file.ncu-rep [--lib path/to/lib.so] [--listing out.sass] [--top N]

For the (first) kernel of the report: instructions executed and stall samples per CUDA source line, and --
with --listing -- the kernel's SASS annotated with file:line, executed count, threads per instruction and samples
(the listing kept under profiles/).  Needs cuobjdump / nvdisasm / ncu on PATH; no GPU.
"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
rep = args[0]
lib = os.path.join(ROOT, "cpp-11-ray-trace-march-framework_b200", "libcuda_trace.so")
listing, top = None, 40
if "--lib" in args: lib = args[args.index("--lib") + 1]
if "--listing" in args: listing = args[args.index("--listing") + 1]
if "--top" in args: top = int(args[args.index("--top") + 1])

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kernel = rows[0][1]
hdr, data = rows[1], rows[2:]
iA, iS, iE, iT, iSm = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "Avg. Threads Executed", "# Samples"))
base = int(data[0][iA], 16)
prof = {int(r[iA], 16) - base: (int(r[iE]), r[iT], int(r[iSm]), r[iS].strip()) for r in data}

# mangled-name fragment from the demangled template arguments: <(int)2, (bool)0, ...> -> ILi2ELb0E...
m = re.search(r"(\w+)<(.*)>\(", kernel)
name, targs = m.group(1), m.group(2)
frag = name + "I" + "".join(("Li%sE" if t.strip().startswith("(int)") else "Lb%sE") % t.split(")")[1].strip()
                            for t in targs.split(","))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = None
for f in sorted(os.listdir(tmp)):
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if frag in out:
        lines = out.split("\n")
        break
if lines is None:
    raise SystemExit("kernel %s not found in %s" % (frag, lib))
start = next(i for i, l in enumerate(lines) if l.startswith("_Z") and frag in l and l.rstrip().endswith(":"))
cur = ("?", 0)
table = []  # (offset, file, line, text)
for l in lines[start + 1:]:
    if l.startswith(".text.") or (l.startswith("_Z") and l.rstrip().endswith(":")):
        break
    mm = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if mm:
        cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
        continue
    mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if mm:
        table.append((int(mm.group(1), 16), cur[0], cur[1], mm.group(2).strip()))
    elif re.match(r"\.L_x_\d+:", l.strip()):
        table.append((None, "", 0, l.strip()))

per_line = defaultdict(lambda: [0, 0, 0])  # instr, samples, sass count
tot_e = tot_s = 0
for off, f, ln, text in table:
    if off is None or off not in prof:
        continue
    e, _, s, _ = prof[off]
    per_line[(f, ln)][0] += e
    per_line[(f, ln)][1] += s
    per_line[(f, ln)][2] += 1
    tot_e += e
    tot_s += s
print("kernel:", kernel)
print("total warp instructions %.3f G, samples %d, SASS instructions %d" % (tot_e / 1e9, tot_s, len(prof)))
print("== instructions executed per source line (top %d)" % top)
for (f, ln), (e, s, n) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6.2f%% instr %6.2f%% samples  %3d SASS  %s:%d" % (100.0 * e / tot_e, 100.0 * s / max(tot_s, 1), n, f, ln))
if listing:
    with open(listing, "w") as o:
        o.write("# %s\n# offset | executed (warp instr) | avg threads | stall samples | SASS | source line\n" % kernel)
        for off, f, ln, text in table:
            if off is None:
                o.write("%s\n" % text)
                continue
            e, t, s, _ = prof.get(off, (0, "-", 0, ""))
            o.write("%05x %12d %5s %7d  %-70s %s:%d\n" % (off, e, t, s, text, f, ln))
    print("listing written:", listing)
