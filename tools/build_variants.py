"""Builds variants/libkgv_<name>.so for tools/variant_bench.py: the same sources with different -D switches.
usage: python tools/build_variants.py name=-DFLAG=1,-DOTHER=2 ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
csrc = g.CSRC
units = sorted(os.path.join(d, f) for d, _, fs in os.walk(csrc) for f in fs if f.endswith((".cu", ".cpp")))
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    out = os.path.join(ROOT, "variants", f"libkgv_{name}.so")
    cmd = ["nvcc"] + g.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-o", out] + units
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    out, _ = p.communicate()
    print(name, "ok" if p.returncode == 0 else "FAILED\n" + out[-2000:])
