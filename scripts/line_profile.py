#!/usr/bin/env python
"""Join an ncu SASS-level source page with nvdisasm line info: instruction and stall shares per source line.
usage: line_profile.py <report.ncu-rep> <mangled-function-substring> [top]"""
import collections, csv, re, subprocess, sys, os
rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof_fast.ncu-rep"
func = sys.argv[2] if len(sys.argv) > 2 else "slicer_fast_kernelILi256ELi4ELi3E"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs("/tmp/cub", exist_ok=True)
subprocess.run("cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all %s/usrp_nfc_b200/libusrp_nfc_b200.so > /dev/null 2>&1" % root, shell=True)
dis = subprocess.run("nvdisasm --print-line-info /tmp/cub/slicer.sm_100a.cubin", shell=True, capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and func in l][0]
end = len(dis)
for i in range(start + 1, len(dis)):
    if dis[i].startswith("//--------------------- .text."):
        end = i
        break
cur, seq = None, []
for l in dis[start:end]:
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((m.group(2).strip(), cur))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[1]; data = rows[2:]
iS, iI, iSt = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
n = min(len(seq), len(data))
by, st, ops = collections.Counter(), collections.Counter(), collections.Counter()
for k in range(n):
    c = int(data[k][iI] or 0)
    by[seq[k][1]] += c; st[seq[k][1]] += int(data[k][iSt] or 0)
    ops[data[k][iS].split()[0] if not data[k][iS].strip().startswith("@") else data[k][iS].split()[1]] += c
tot, tots = sum(by.values()), max(1, sum(st.values()))
srcs = {}
def line(key):
    if not key: return "?"
    f, ln = key
    if f not in srcs:
        pth = os.path.join(root, "usrp_nfc_b200/csrc", f)
        srcs[f] = open(pth).read().splitlines() if os.path.exists(pth) else []
    return srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else f
print("sass", len(seq), "ncu rows", len(data), "warp instr", tot)
print("opcodes:", ", ".join("%s %.1f%%" % (o, 100 * c / tot) for o, c in ops.most_common(28)))
for key, c in by.most_common(top):
    print("%-22s inst %5.1f%% stall %5.1f%%  %s" % ("%s:%d" % key if key else "?", 100 * c / tot, 100 * st[key] / tots, line(key)))
if len(sys.argv) > 4:
    # regions: file:a-b,...
    print("regions:")
    for spec in sys.argv[4].split(","):
        f, rng = spec.split(":"); a, b = [int(v) for v in rng.split("-")]
        ci = sum(c for k, c in by.items() if k and k[0] == f and a <= k[1] <= b)
        cs = sum(c for k, c in st.items() if k and k[0] == f and a <= k[1] <= b)
        print("  %-28s inst %5.1f%%  stall %5.1f%%" % (spec, 100 * ci / tot, 100 * cs / tots))
    other = sum(c for k, c in by.items() if not k or k[0] not in ("slicer_fast.cuh", "slicer.cu"))
    print("  other files inst %5.1f%%" % (100 * other / tot))
