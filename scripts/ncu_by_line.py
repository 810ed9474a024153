"""Join an ncu SASS profile with nvdisasm line info: warp-instructions, active threads and stall
samples per CUDA source line.  usage: python scripts/ncu_by_line.py rep.ncu-rep lib.so kernel_substr src.cu [top]"""
import csv, io, re, subprocess, sys, os, tempfile, collections

rep, lib, kname, srcfile = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 45
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if os.path.basename(srcfile).split(".")[0] in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# walk the disassembly of the wanted kernel: remember the current //## File ..., line N marker per instruction
offs_line = {}
cur = None; infunc = False; idx = 0; order = []
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        infunc = kname in m.group(1); idx = 0; continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        offs_line[int(m.group(1), 16)] = cur; order.append(int(m.group(1), 16))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = list(csv.reader(io.StringIO(src)))
st = [i for i, l in enumerate(lines) if l and l[0] == "Address"][0]
h = lines[st]; body = [l for l in lines[st + 1:] if len(l) == len(h) and l[0].startswith("0x")]
base = int(body[0][0], 16)
ie, te, sm = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for l in body:
    key = offs_line.get(int(l[0], 16) - base, ("?", 0))
    a = agg[key]; a[0] += int(l[ie]); a[1] += int(l[te]); a[2] += int(l[sm]); a[3] += 1
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
text = open(srcfile).read().splitlines()
print("total warp-instructions %d, stall samples %d" % (tot, tots))
print("| line | SASS | warp-inst %% | avg thr | samples %% | source |\n|---|---|---|---|---|---|")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    s = text[key[1] - 1].strip()[:90] if key[0] == os.path.basename(srcfile) and 0 < key[1] <= len(text) else str(key)
    print("| %s:%d | %d | %.2f | %.1f | %.2f | `%s` |" % (key[0], key[1], a[3], 100.0 * a[0] / tot, a[1] / max(1, a[0]), 100.0 * a[2] / max(1, tots), s))
