"""Summarise FLD_TC_TRACE dumps of deconv_gemm_kernel (CTA 0): per-role event deltas."""
import glob
import sys
import numpy as np
for fn in sorted(glob.glob(sys.argv[1] + "/trace_deconv_*_s8_*.txt"))[-2:]:
    ev = {0: [], 1: [], 2: []}
    for ln in open(fn):
        r, tag, clk = ln.split()
        ev[int(r)].append((int(tag), int(clk)))
    print("==", fn)
    t0 = min(e[0][1] for e in ev.values() if e)
    for r, name in ((0, "producer"), (1, "mma"), (2, "epilogue")):
        e = ev[r]
        print(name, "events", len(e), "span", (e[-1][1] - e[0][1]) if e else 0)
        print("   first 60:", " ".join("%d:%d" % (t, c - t0) for t, c in e[:60]))
    m = ev[1]
    tiles = [c for t, c in m if t == 3]
    if len(tiles) > 4:
        d = np.diff(tiles)
        print("mma tile period: median %d min %d max %d (n=%d)" % (np.median(d), d.min(), d.max(), len(d)))
    fw = [(m[i][1] - m[i - 1][1]) for i in range(1, len(m)) if m[i][0] == 1]
    if fw:
        print("mma wait-for-full (from previous event): median %d mean %d max %d" % (np.median(fw), np.mean(fw), max(fw)))
    e = ev[2]
    w = [(e[i][1] - e[i - 1][1]) for i in range(1, len(e)) if e[i][0] == 1 and e[i - 1][0] == 0]
    l = [(e[i][1] - e[i - 1][1]) for i in range(1, len(e)) if e[i][0] == 2 and e[i - 1][0] == 1]
    b = [(e[i][1] - e[i - 1][1]) for i in range(1, len(e)) if e[i][0] == 0 and e[i - 1][0] == 2]
    if w:
        print("epilogue: wait tfull median %d | tmem ld median %d | rest-of-tile median %d" % (np.median(w), np.median(l), np.median(b)))
