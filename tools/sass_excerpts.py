#!/usr/bin/env python
"""profiles/rNN_sass_excerpts.txt: opcode histograms and short SASS excerpts of the two big kernels of libvp_b200.so
(the blend loop of k_reproject_hoist4, one row of the FAST block of k_grad_circ<6>, its TMA issue sequences).

    python tools/sass_excerpts.py > profiles/r02_sass_excerpts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTR = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")


def listing(name):
    tmp = f"/tmp/_{name}.sass"
    hist = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_fn.py"), name, "--loop", "--out", tmp], capture_output=True, text=True, check=True).stdout
    return hist, [(int(m.group(1), 16), m.group(2).rstrip()) for line in open(tmp) if (m := INSTR.match(line))]


def show(ins, lo, hi):
    return "".join(f"        /*{a:04x}*/  {t}\n" for a, t in ins[max(lo, 0):hi])


h4_hist, h4 = listing("k_reproject_hoist4ILi0")
gc_hist, gc = listing("k_grad_circILi6ELb0")
out = ["SASS excerpts of the two big kernels (cuobjdump -sass of vision-processor_b200/lib/libvp_b200.so via tools/sass_fn.py; sm_100a, nvcc 12.9)\n",
       "=" * 100 + "\n1. k_reproject_hoist4<RGGB>: opcode histogram of the frame loop (one iteration = 2 quads of frames x 4 pixels per thread = 32 pixel-frames)\n",
       h4_hist,
       """
Reading: 704 PRMT = 512 byte -> fp32-denormal unpacks (one PRMT per tap, pixel and frame: 16 taps x 32 pixel-frames)
+ 64 for the byte transpose of the staging + 128 for the paired integer tail; 256 FMUL2 = the 16 products of a pixel for two frames
at once (scalar weight broadcast: the ".F32" operand); 256 FADD2.FTZ = the 12 sums + 4 rounding adds; 128 FMUL = the 16 weights of a
pixel, formed once per quad of frames from the eight axis values; 139 LDS = 128 taps (one word = one texel of FOUR frames) + staging.
Round 1 had 2421 instructions here (FFMA2 for every product and sum, one more FMUL per tap, the integer tail frame by frame).
"""]
ftz = [i for i, (a, t) in enumerate(h4) if t.startswith("FADD2.FTZ")]
out.append("\nExcerpt (taps of one pixel; LDS = one texel of four frames, PRMT x4 = four frames' bytes as denormals, FMUL2 x2 = (A,B) and (C,D) products,\n"
           "FADD2.FTZ = the rounded sums, which ptxas cannot contract with the FMUL2 feeding them):\n")
out.append(show(h4, ftz[40] - 30, ftz[40] + 30))
k = [i for i, (a, t) in enumerate(h4) if re.match(r"IMAD R\d+, R\d+, 0xc0, R\d+", t)]
out.append("\nExcerpt (integer tail of TWO frames in 16-bit lanes: PRMT 0x5410 packs, IMAD x 0xc0 = (3x - s + 510) << 6 in both lanes,\n"
           "PRMT 0x7351 / 0x4341 / 0x5410 / 0x7632 unpack RGBA; interleaved with the next pixel's taps):\n")
out.append(show(h4, k[0] - 16, k[0] + 18))
out.append("\n" + "=" * 100 + "\n2. k_grad_circ<6, even offset>: whole function histogram\n")
out.append(gc_hist)
idp = [i for i, (a, t) in enumerate(gc) if t.startswith("IDP")]
a0, e = idp[0], idp[63]
while not gc[e][1].startswith("VOTE"):
    e += 1
hist = collections.Counter((t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0] for a, t in gc[a0 - 4:e])
out.append(f"\nFAST group block (8 rows x 64 columns per warp, straight-line): {e - a0 + 4} instructions = {(e - a0 + 4) / 8:.1f} per row\n   "
           + "  ".join(f"{k_}:{v}" for k_, v in hist.most_common()) + "\n")
out.append("\nExcerpt (about one row of the FAST block: LDS.64 taps, 8 DP4A, magic int->float, gradDot store, vertical window V, strip sum REDUX,\n"
           "horizontal window by SHFL, the far term by SHFL.UP, FMNMX3, exact division by R^2 as FMUL2 + 2 FFMA2, blobCenter store):\n")
out.append(show(gc, a0 + 38, a0 + 108))
u = [i for i, (a, t) in enumerate(gc) if "UTMALDG" in t][0]
out.append("\nExcerpt (staging one group of D = 8 flat rows: mbarrier expect_tx + ONE tensor copy, issued under elect.sync -- operands in uniform registers, no per-lane loop):\n")
out.append(show(gc, u - 16, u + 3))
ub = [i for i, (a, t) in enumerate(gc) if "UBLKCP" in t][0]
out.append("\nExcerpt (top / bottom groups: row-by-row cp.async.bulk with the row index clamped, CLAMP_TO_EDGE of gradientDot.cl:20):\n")
out.append(show(gc, ub - 8, ub + 2))
sys.stdout.write("".join(out))
