"""Correlates an ncu report's per-SASS-instruction counters with CUDA source lines (nvdisasm -g line info).

    python tools/ncu_lines.py <report.ncu-rep> <kernel-substring> [object.cubin-substring] [top N]

Needs ncu, cuobjdump and nvdisasm on PATH (no GPU).  Output: share of executed warp instructions and of
warp-stall samples per source line."""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    kname = rows[0][1]
    hdr, data = rows[1], [r for r in rows[2:] if len(r) == len(rows[1])]
    ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    so = os.path.join(ROOT, "maddpg_b200", "_lib", "libmaddpg_b200.so")
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
    mang = None
    insts = []
    for cub in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.splitlines()
        secs = [i for i, l in enumerate(txt) if l.startswith(".text.")]
        for si, s in enumerate(secs):
            dem = subprocess.run(["c++filt", txt[s][6:].rstrip(":")], capture_output=True, text=True).stdout.strip()
            if _norm(dem) != _norm(kname):
                continue
            end = secs[si + 1] if si + 1 < len(secs) else len(txt)
            cur = None
            for l in txt[s:end]:
                m = re.search(r'//## File "([^"]+)", line (\d+)', l)
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
                if m:
                    insts.append((int(m.group(1), 16), cur))
            mang = dem
            break
        if mang:
            break
    assert insts and len(data) >= len(insts), (len(insts), len(data), kname)
    data = data[:len(insts)]  # the report may hold several launches of the kernel: use the first
    base = int(data[0][0], 16)
    a2l = dict(insts)
    by, samp = collections.Counter(), collections.Counter()
    for r in data:
        ln = a2l.get(int(r[0], 16) - base)
        by[ln] += float(r[ia])
        samp[ln] += float(r[isamp])
    tot, ts = sum(by.values()), max(1.0, sum(samp.values()))
    print("kernel: %s\nwarp instructions executed: %d" % (kname, tot))
    src = {}
    for ln, v in sorted(by.items(), key=lambda kv: -kv[1])[:top]:
        f, n = ln if ln else ("?", 0)
        if f not in src:
            p = os.path.join(ROOT, "maddpg_b200", "csrc", f)
            src[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[f][n - 1].strip() if 0 < n <= len(src[f]) else ""
        print("%5.1f%% inst %5.1f%% stall-samples  %-18s:%4d  %s" % (100 * v / tot, 100 * samp[ln] / ts, f, n, text[:100]))


def _norm(name):
    """kernel name up to the argument list, with template arguments normalised"""
    n = name.replace("void ", "").replace("(int)", "").replace("(bool)", "").replace(" ", "")
    n = n.replace("true", "1").replace("false", "0")
    return n.split("(")[0]


if __name__ == "__main__":
    main()
