"""Dump the SASS of the hot loops of the library's kernels into profiles/ (cuobjdump -sass of the
built .so, here, no GPU needed): for each kernel the innermost backward-branch loop that contains
its signature instruction, with an instruction-mix header.  Usage: python tools/sass_hot_loops.py [round tag]"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
sass = subprocess.run(["cuobjdump", "-sass", str(ROOT / "taxi2_b200" / "lib" / "libtaxi2_b200.so")], capture_output=True, text=True, check=True).stdout
functions = {}
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        functions[name] = []
    elif name:
        functions[name].append(line)

TARGETS = [  # (mangled-name fragment, signature instruction, minimum count in the loop, output name, note)
    ("gotoh_pair16_kernelILi21ELi1ELb0E", "VIMNMX3.U16x2", 21, "gotoh_pair16_21_1", "one column step of 21 row slots x 2 pairs per lane (42 cells per lane)"),
    ("gotoh_pair16_kernelILi21ELi1ELb1E", "VIMNMX3.U16x2", 21, "gotoh_pair16_21_1_sym", "both-orientations variant (Ix / Iy tie bit per cell), one column step"),
    ("gotoh_pair16_kernelILi21ELi2ELb0E", "VIMNMX3.U16x2", 21, "gotoh_pair16_21_2", "multi-stripe variant, one column step"),
    ("gotoh_warp_kernelILi21E", "VIMNMX3", 21, "gotoh_warp_21", "general int32 kernel, one column step of 21 rows per lane"),
    ("count_rect_kernel", "POPC", 40, "count_rect", "popcount kernel: 4 x rows x 5 words per iteration (20 word pairs)"),
    ("count_tc_kernel", "UTCIMMA", 4, "count_tc_mma", "tensor-core kernel: the MMA issue loop (4 UTCIMMA per 128-byte k-block)"),
]
for frag, sig, need, out, note in TARGETS:
    fname = next((n for n in functions if frag in n), None)
    if not fname:
        print("missing", frag)
        continue
    ins = []
    for line in functions[fname]:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for addr, text in ins:
        m = re.search(r"BRA\S* (?:\S+, )?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr:
            loops.append((addr - int(m.group(1), 16), int(m.group(1), 16), addr))
    body = None
    for _, lo, hi in sorted(loops):
        cand = [(a, t) for a, t in ins if lo <= a <= hi]
        if sum(sig in t.split()[1 if t.startswith("@") else 0] for _, t in cand) >= need:
            body = cand
            break
    if body is None:
        print("no loop found for", frag)
        continue
    mix = collections.Counter((t.split()[1] if t.startswith("@") else t.split()[0]) for _, t in body)
    path = ROOT / "profiles" / f"sass_{out}_{tag}.txt"
    with open(path, "w") as f:
        f.write(f"# {fname}\n# {note}\n# {len(body)} instructions; mix: " + ", ".join(f"{k} {v}" for k, v in mix.most_common()) + "\n")
        for a, t in body:
            f.write(f"/*{a:05x}*/ {t}\n")
    print(path.name, len(body), "instructions")
