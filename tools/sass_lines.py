#!/usr/bin/env python3
"""tools/sass_lines.py OBJ KERNEL_SUBSTR [LO HI] - static instruction budget of a kernel, per source line.

Disassembles OBJ (.o or .cubin built with -lineinfo) with `nvdisasm -g`, picks the kernel whose mangled name
contains KERNEL_SUBSTR, and prints how many SASS instructions inside the address range [LO, HI) (hex; default:
whole kernel) each source line owns, plus a histogram of opcodes.  The per-burst loop of the demodulation kernels is
straight-line code apart from a few cold branches and the early/late round loop, so "instructions in the loop's
address range, by source line" is a usable estimate of warp-instructions per burst without a GPU; ncu's
smsp__inst_executed on the box is the measurement (profiles/).
"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def disasm(obj):
    if obj.endswith(".cubin"):
        cub = obj
    else:
        d = tempfile.mkdtemp()
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
        cub = os.path.join(d, [f for f in os.listdir(d) if f.endswith(".cubin")][0])
    return subprocess.run(["nvdisasm", "-g", "-c", cub], check=True, capture_output=True, text=True).stdout


def main():
    obj, sub = sys.argv[1], sys.argv[2]
    lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
    txt = disasm(obj).split("\n")
    cur, inside = None, False
    per_line = collections.Counter()
    ops = collections.Counter()
    total = 0
    for ln in txt:
        if ln.startswith(".text."):
            inside = sub in ln
            cur = None
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            addr = int(m.group(1), 16)
            if lo <= addr < hi:
                ins = m.group(2).split()
                op = ins[1] if ins[0].startswith("@") else ins[0]
                per_line[cur] += 1
                ops[op.split(".")[0]] += 1
                total += 1
    print("total", total)
    for (f, l), c in sorted(per_line.items(), key=lambda kv: (kv[0][0], kv[0][1]) if kv[0] else ("", 0)):
        print(f"{f}:{l}\t{c}")
    print("--- opcodes")
    for op, c in ops.most_common(40):
        print(f"{op}\t{c}")


if __name__ == "__main__":
    main()
