"""Attribute executed instructions / stall samples of one kernel in an ncu report to CUDA source lines
(correlates the SASS page of the report with nvdisasm -g line info of the in-tree libkml.so)."""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kernel, cubin_name = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', cubin_name, os.path.join(ROOT, 'kimera-multi_b200', 'libkml.so')], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
sass = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split('\n')
start = [i for i, l in enumerate(sass) if l.strip().startswith('.section') and kernel in l and '.text.' in l][0]
seq = {}; cur = None
for l in sass[start + 1:]:
    if l.strip().startswith('.section'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: seq[int(m.group(1), 16)] = cur
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '-k', 'regex:' + kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split('\n')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; data = []
for r in rows[hi + 1:]:
    if not r or r[0] in ('Kernel Name', 'Address'): break
    data.append(r)
ie = hdr.index('Instructions Executed'); ss = hdr.index('# Samples')
base = int(data[0][0], 16)
byline = collections.Counter(); samp = collections.Counter()
for r in data:
    key = seq.get(int(r[0], 16) - base) or ('?', 0)
    byline[key] += int(r[ie]); samp[key] += int(r[ss])
tot = sum(byline.values()); tots = sum(samp.values())
src = {}
for f in os.listdir(os.path.join(ROOT, 'kimera-multi_b200', 'csrc')):
    src[f] = open(os.path.join(ROOT, 'kimera-multi_b200', 'csrc', f)).read().split('\n')
print("static SASS %d, executed %d, samples %d" % (len(data), tot, tots))
# per-phase aggregation for fivept_warp.cuh
for (f, ln), v in byline.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 40):
    s = src[f][ln - 1].strip()[:64] if f in src and ln - 1 < len(src[f]) else ''
    print("%5.1f%% instr %5.1f%% stall  %-16s:%4d  %s" % (100 * v / tot, 100 * samp[(f, ln)] / max(tots, 1), f, ln, s))
