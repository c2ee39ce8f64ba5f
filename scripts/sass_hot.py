#!/usr/bin/env python
"""Hot SASS of a kernel in program order with executed counts and source lines.
usage: sass_hot.py <report> <mangled-substring> <min-fraction-of-max-count>"""
import csv, re, subprocess, sys, os
rep, func, frac = sys.argv[1], sys.argv[2], float(sys.argv[3])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs("/tmp/cub", exist_ok=True)
subprocess.run("cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all %s/usrp_nfc_b200/libusrp_nfc_b200.so > /dev/null 2>&1" % root, shell=True)
dis = subprocess.run("nvdisasm --print-line-info /tmp/cub/slicer.sm_100a.cubin", shell=True, capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and func in l][0]
end = len(dis)
for i in range(start + 1, len(dis)):
    if dis[i].startswith("//--------------------- .text."):
        end = i; break
cur, seq = None, []
for l in dis[start:end]:
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: seq.append((m.group(2).strip(), cur))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[1]; data = rows[2:]
iS, iI, iSt = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
cnt = [int(d[iI] or 0) for d in data]
mx = max(cnt)
tot = sum(cnt)
acc = 0
for k in range(min(len(seq), len(data))):
    if cnt[k] >= frac * mx:
        acc += cnt[k]
        print("%5d %6.3f%% st%5s %-18s %s" % (k, 100.0 * cnt[k] / tot, data[k][iSt], "%s:%d" % seq[k][1] if seq[k][1] else "?", data[k][iS][:110]))
print("shown share %.1f%%" % (100.0 * acc / tot))
