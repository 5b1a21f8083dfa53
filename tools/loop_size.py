"""Footprint of the sweep pass loop in the built library (KB of SASS)."""
import re, subprocess, os, sys, tempfile
so = sys.argv[1] if len(sys.argv) > 1 else "dftpav_b200/csrc/libdftpav_b200.so"
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.startswith("dftpav_kernels") and "api" not in f][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
src = open("dftpav_b200/csrc/dftpav_kernels.cu").read().split("\n")
l0 = [i for i, l in enumerate(src, 1) if "for (int pass = pass0; pass < npass; pass += kTeam) {" in l][0]
l1 = [i for i, l in enumerate(src, 1) if "// segment totals" in l][0]
for kern in ("solve_kernelILb0",):
    start = [i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l][0]
    cur = None; addrs = []; alla = []
    for l in sass[start + 1:]:
        if l.startswith("//---") and ".text." in l: break
        m = re.search(r'//## File "[^"]*", line (\d+)', l)
        if m: cur = int(m.group(1)); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
        if m:
            a = int(m.group(1), 16); alla.append(a)
            if cur and l0 <= cur <= l1: addrs.append(a)
    print(f"{so}: kernel {(max(alla)+16)/1024:.1f} KB, sweep pass loop {(max(addrs)-min(addrs))/1024:.1f} KB")
