"""Join an ncu SASS-page CSV (ncu -i X.ncu-rep --page source --csv) with nvdisasm -g line info of the same cubin and
aggregate executed instructions / stall samples per CUDA source line and per function (line ranges).
usage: ncu_lines.py <sass_csv> <nvdisasm.sass> <kernel-substring> <source.cu> [top]"""
import csv, re, sys, collections

csv_path, sass_path, kern, src_path = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][0], 16)
# nvdisasm: find section of the kernel
lines = open(sass_path).read().split("\n")
start = None
for i, l in enumerate(lines):
    if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"):
        start = i
        break
assert start is not None, "kernel not found"
off2line = {}
cur = None
for l in lines[start + 1:]:
    if l.startswith("//---") and ".text." in l:
        break
    m = re.search(r'//## File "[^"]*", line (\d+)(.*)', l)
    if m:
        # for inlined code nvdisasm prints "inlined at" chains on following lines; keep the innermost line
        cur = int(m.group(1))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s", l)
    if m and cur is not None:
        off2line[int(m.group(1), 16)] = cur
src = open(src_path).read().split("\n")
# function ranges: top-level "name(" definitions
func_at = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__|__host__).*?\b(\w+)\s*\(", l)
    if m:
        func_at.append((i, m.group(1)))
    else:
        m2 = re.match(r"^\s{4}(\w+_kernel)\(", l)
        if m2:
            func_at.append((i, m2.group(1)))
def func_of(line):
    name = "?"
    for i, n in func_at:
        if i <= line:
            name = n
    return name
agg_line = collections.defaultdict(lambda: collections.Counter())
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    off = int(r[0], 16) - base
    ln = off2line.get(off, -1)
    c = agg_line[ln]
    c["inst"] += int(r[col["Instructions Executed"]] or 0)
    c["thr"] += int(r[col["Thread Instructions Executed"]] or 0)
    c["samples"] += int(r[col["# Samples"]] or 0)
    for s in stall_cols:
        c[s] += int(r[col[s]] or 0)
    if "DFMA" in r[1] or "DADD" in r[1] or "DMUL" in r[1] or "DSETP" in r[1] or "MUFU" in r[1]:
        c["fp64"] += int(r[col["Instructions Executed"]] or 0)
for ln, c in agg_line.items():
    tot.update(c)
print(f"total inst {tot['inst']:,} samples {tot['samples']:,} fp64-inst {tot['fp64']:,} avg thr/inst {tot['thr']/max(1,tot['inst']):.1f}")
byf = collections.defaultdict(lambda: collections.Counter())
for ln, c in agg_line.items():
    byf[func_of(ln)].update(c)
print("\nper function: inst% samples% fp64% thr/inst top-stalls")
for f, c in sorted(byf.items(), key=lambda kv: -kv[1]["samples"]):
    st = sorted(((s, c[s]) for s in stall_cols), key=lambda kv: -kv[1])[:4]
    print(f"  {f:28s} inst {100*c['inst']/tot['inst']:5.1f}%  samples {100*c['samples']/tot['samples']:5.1f}%  fp64 {100*c['fp64']/max(1,c['inst']):4.0f}%  thr/inst {c['thr']/max(1,c['inst']):4.1f}  " + " ".join(f"{s[6:]}={100*v/max(1,c['samples']):.0f}%" for s, v in st))
print(f"\ntop {top} lines by samples:")
for ln, c in sorted(agg_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((s, c[s]) for s in stall_cols), key=lambda kv: -kv[1])[:3]
    text = src[ln - 1].strip()[:90] if 0 < ln <= len(src) else "?"
    print(f"  L{ln:4d} {100*c['samples']/tot['samples']:5.2f}% inst {100*c['inst']/tot['inst']:5.2f}% thr {c['thr']/max(1,c['inst']):4.1f} " + " ".join(f"{s[6:]}={100*v/max(1,c['samples']):.0f}%" for s, v in st) + f" | {text}")
