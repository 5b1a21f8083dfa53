"""SASS footprint (bytes of code) of a kernel attributed to source statements, following the full inline chain of
`nvdisasm -gi`.  usage: code_map.py [lib.so] [kernel-substring] [--fn NAME]...   Every instruction is charged to the
outermost-but-one statement inside each requested function (default: eval_cost_grad, sweep_segment, solve_kernel)."""
import re, subprocess, os, sys, tempfile, collections
args = sys.argv[1:]
fns = []
while "--fn" in args:
    i = args.index("--fn"); fns.append(args[i + 1]); del args[i:i + 2]
so = args[0] if len(args) > 0 else "dftpav_b200/csrc/libdftpav_b200.so"
kern = args[1] if len(args) > 1 else "solve_kernelILb0"
fns = fns or ["solve_kernel", "eval_cost_grad", "sweep_segment"]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.startswith("dftpav_kernels") and "api" not in f][0]
sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
src = open("dftpav_b200/csrc/dftpav_kernels.cu").read().split("\n")
# function extents: a definition starts at a column-0 line containing NAME( and ends at the next column-0 '}'
ext = {}
for name in fns:
    s = [i for i, l in enumerate(src, 1) if re.search(r"\b%s\s*\(const" % name, l) and not l.lstrip().startswith("//")][0]
    e = next(i for i in range(s, len(src) + 1) if src[i - 1].startswith("}"))
    ext[name] = (s, e)
start = [i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l][0]
chain = []; fresh = True
per = {n: collections.Counter() for n in fns}
total = 0
for l in sass[start + 1:]:
    if l.startswith("//---") and ".text." in l: break
    m = re.search(r'//## File "([^"]*)", line (\d+)', l)
    if m:
        if fresh: chain = []; fresh = False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    if re.match(r"\s+/\*([0-9a-f]{4,})\*/", l):
        fresh = True; total += 16
        for name, (s, e) in ext.items():
            hit = [ln for f, ln in chain if f == "dftpav_kernels.cu" and s <= ln <= e]
            if hit: per[name][hit[-1] if name != "leaf" else hit[0]] += 16
print(f"{kern}: {total/1024:.1f} KB")
for name in fns:
    tot = sum(per[name].values())
    print(f"== {name} lines {ext[name]}: {tot/1024:.1f} KB")
    for ln, v in sorted(per[name].items()):
        if v >= 256: print(f"  {ln:5d} {v:6d} B  {src[ln-1].strip()[:110]}")
