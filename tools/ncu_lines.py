"""Executed warp instructions per CUDA source line: joins `ncu --page source --csv` (SASS rows in order) with `nvdisasm -g` of the
same kernel.  usage: ncu_lines.py <source.csv> <nvdisasm.txt> [top] [source dir of the profiled build]"""
import csv, re, collections, sys, os
dis = open(sys.argv[2]).read().splitlines()
cur = None; seq = []
for l in dis:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m: seq.append(cur)
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ie = hdr.index("Instructions Executed"); ist = hdr.index("# Samples")
data = rows[2:]
assert len(seq) == len(data), (len(seq), len(data))
by = collections.Counter(); bys = collections.Counter()
for i in range(len(seq)):
    by[seq[i]] += int(data[i][ie]); bys[seq[i]] += int(data[i][ist])
tot = sum(by.values()); print("total warp instructions", tot, " samples", sum(bys.values()))
root = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fast-3d-pointcloud-segmentation_b200", "csrc")
src = {}
for k, v in by.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 30):
    line = ""
    if k:
        if k[0] not in src:
            try: src[k[0]] = open(os.path.join(root, k[0])).read().splitlines()
            except OSError: src[k[0]] = []
        if 0 < k[1] <= len(src[k[0]]): line = src[k[0]][k[1] - 1].strip()[:100]
    print("%6.2f%% %10d  samples %5d  %s:%s  %s" % (100 * v / tot, v, bys[k], k[0] if k else None, k[1] if k else 0, line))
