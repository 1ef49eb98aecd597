"""SASS of one kernel of libvp_b200.so (or any cubin-bearing file) and an opcode histogram of it.

    python tools/sass_fn.py k_reproject_hoist4ILi0 [--loop] [--out FILE] [--lib PATH]

--loop restricts the histogram to the instructions between the first backward branch target and that branch with the
largest body (the hot loop of a streaming kernel); the listing written by --out is always the whole function."""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def functions(lib):
    text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, name, body = {}, None, []
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                out[name] = body
            name, body = m.group(1), []
        elif name:
            body.append(line)
    if name:
        out[name] = body
    return out


INSTR = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)\s*(.*?);")


def parse(body):
    ins = []
    for line in body:
        m = INSTR.match(line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2), m.group(3), m.group(4)))
    return ins


def hot_loop(ins):
    best = None
    for addr, op, mod, args in ins:
        if op == "BRA":
            m = re.search(r"0x([0-9a-f]+)", args)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                    best = (tgt, addr)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("pattern")
    ap.add_argument("--lib", default=os.path.join(ROOT, "vision-processor_b200", "lib", "libvp_b200.so"))
    ap.add_argument("--loop", action="store_true")
    ap.add_argument("--full-op", action="store_true", help="histogram with modifiers (LDS.128 vs LDS)")
    ap.add_argument("--out")
    a = ap.parse_args()
    fns = functions(a.lib)
    hits = [n for n in fns if a.pattern in n]
    if not hits:
        raise SystemExit(f"no function matches {a.pattern!r}")
    for name in hits:
        ins = parse(fns[name])
        rng = hot_loop(ins) if a.loop else None
        sel = [i for i in ins if rng is None or rng[0] <= i[0] <= rng[1]]
        hist = collections.Counter((op + mod) if a.full_op else op for _, op, mod, _ in sel)
        print(f"== {name}: {len(ins)} instructions" + (f", loop {rng[0]:#x}..{rng[1]:#x} = {len(sel)}" if rng else ""))
        print("   " + "  ".join(f"{op}:{n}" for op, n in hist.most_common()))
        if a.out:
            with open(a.out if len(hits) == 1 else f"{a.out}.{hits.index(name)}", "w") as f:
                f.write(f"Function : {name}\n" + "\n".join(fns[name]) + "\n")


if __name__ == "__main__":
    main()
