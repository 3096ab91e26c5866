#!/usr/bin/env python3
"""Instruction mix and global/shared/local memory instructions of the hot kernels of libfmb200.so (cuobjdump -sass), written as
profiles/rNN_sass_hot_kernels.txt:   python tools/sass_summary.py > profiles/r02_sass_hot_kernels.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fmindex-collection_b200", "libfmb200.so")
HOT = ("gather_probe_kernel", "locate_pair_kernel", "locate_shortcut_kernel", "exact_search2_kernel", "exact_search_kernel",
       "scheme_search_kernel", "scheme_text_kernel", "unpack_queries_kernel", "text_class_keys_kernel", "gather_items_kernel")
MEM = re.compile(r"^(LD|ST|ATOM|RED|LDG|STG|LDS|STS|LDL|STL|LDC|LDCU|UBLKCP|UTMA)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    print("# cuobjdump -sass libfmb200.so (sm_100a): instruction mix and every memory instruction of the hot kernels")
    print("# exact_search2_kernel: one LDG.E.256 per lane for a pair-table quarter (4 lanes = one 128-byte line), LDG.E.128 for the merged LF16/LF32")
    print("# entry, LDG.E.64 for k-mer / LF4 entries; scheme_text_kernel: the window words live in shared memory (LDS/STS), node stacks in local")
    print("# memory (LDL/STL); no tensor-core or TMA instructions anywhere: the path is integer gather work (DESIGN.md section 3)")
    name, ops = None, None
    out = []

    def flush():
        if name and any(h in name for h in HOT):
            out.append((name, ops))

    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            flush()
            name, ops = m.group(1), []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            ops.append(m.group(1))
    flush()
    for name, ops in out:
        mix = collections.Counter(o.split(".")[0] for o in ops)
        mem = collections.Counter(o for o in ops if MEM.match(o))
        print(f"\n== {name}")
        print(f"   {len(ops)} instructions; " + ", ".join(f"{k} {v}" for k, v in mix.most_common(14)))
        print("   memory instructions: " + ", ".join(f"{k} x{v}" for k, v in sorted(mem.items())))


if __name__ == "__main__":
    sys.exit(main())
