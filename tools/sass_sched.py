#!/usr/bin/env python3
"""Decode the static scheduling fields (stall count, yield, barriers) from `cuobjdump -sass` output for an
address range and print the single-warp issue timeline: sum of stall counts = lower bound on the cycles one
warp needs for the range when nothing else hides the latencies.  Usage: sass_sched.py file.sass 0x499a0 0x4a2c0 [-v]"""
import re, sys
def parse(path):
    ins = []
    cur = None
    for line in open(path):
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", line)
        if m:
            cur = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]
            continue
        m = re.match(r"\s*/\* (0x[0-9a-f]{16}) \*/", line)
        if m and cur:
            cur[3] = int(m.group(1), 16)
            ins.append(tuple(cur)); cur = None
    return ins
def fields(hi):
    c = hi >> 41
    return {"stall": c & 0xF, "yield": (c >> 4) & 1, "wr": (c >> 5) & 7, "rd": (c >> 8) & 7, "wait": (c >> 11) & 0x3F, "reuse": (c >> 17) & 0xF}
if __name__ == "__main__":
    ins = parse(sys.argv[1]); lo = int(sys.argv[2], 16); hi = int(sys.argv[3], 16); v = "-v" in sys.argv
    tot = 0; n = 0; ops = {}
    for a, txt, w0, w1 in ins:
        if a < lo or a >= hi: continue
        f = fields(w1); tot += max(f["stall"], 1); n += 1
        op = txt.split()[0] if not txt.startswith("@") else txt.split()[1]
        ops[op] = ops.get(op, 0) + 1
        if v: print("%05x  s%-2d y%d wr%d rd%d w%02x  %s" % (a, f["stall"], f["yield"], f["wr"], f["rd"], f["wait"], txt))
    print("instructions %d, sum of stall counts %d cycles (%.2f per instr)" % (n, tot, tot / max(n, 1)))
    print(sorted(ops.items(), key=lambda x: -x[1]))
