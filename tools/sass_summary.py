#!/usr/bin/env python
"""Build evidence kept under profiles/ (VERDICT r01 item 9):
  profiles/r02_sass_summary.txt  per kernel of libdfx.so: architecture, instruction count, and the counts of the
                                 SASS mnemonics that prove the Blackwell data path (UBLKCP = cp.async.bulk / TMA bulk
                                 copy, UTMACMDFLUSH, DMMA = FP64 tensor-core mma.sync, DFMA, LDGSTS, ...)
  profiles/r02_ptxas.txt         registers / spills / stack / shared memory per kernel from `nvcc -Xptxas -v`
Run after `python -c "import __graft_entry__ as g; g.build_lib(force=True, verbose=True)"`."""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dune_functions_b200", "lib", "libdfx.so")
OBJ = os.path.join(ROOT, "dune_functions_b200", "lib", "obj")
WATCH = ["UBLKCP", "UTMACMDFLUSH", "UTMASTG", "UTMALDG", "DMMA", "HMMA", "UTCHMMA", "LDTM", "DFMA", "DADD", "DMUL", "LDGSTS", "STS", "LDS", "STG", "LDG", "ST", "BAR", "MEMBAR", "FENCE"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            kernels[cur]["_total"] += 1
            kernels[cur][m.group(1)] += 1
    names = demangle(list(kernels))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "r02_sass_summary.txt"), "w") as f:
        f.write(f"cuobjdump -sass dune_functions_b200/lib/libdfx.so   arch = {', '.join(arch)}   kernels = {len(kernels)}\n")
        tot = collections.Counter()
        for c in kernels.values():
            tot.update(c)
        f.write("whole library: " + "  ".join(f"{k}={tot[k]}" for k in WATCH if tot[k]) + "\n")
        f.write("no tcgen05 (UTC*MMA / LDTM) by design: there is no f64 tcgen05 kind; the FP64 tensor path is mma.sync (DMMA)\n\n")
        for k, c in kernels.items():
            short = re.sub(r"\(.*", "", names[k].replace("dfx::", "").replace("void ", ""))
            f.write(f"{short}\n    instructions={c['_total']}  " + "  ".join(f"{w}={c[w]}" for w in WATCH if c[w]) + "\n")
    rows = []
    for log in sorted(glob.glob(os.path.join(OBJ, "*.ptxas.txt"))):
        txt = open(log).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for '(sm_\w+)'\n(?:ptxas info\s+: Function properties for \S+\n)?\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", txt):
            rows.append((os.path.basename(log).replace(".ptxas.txt", ".cu"), m.group(1), m.group(2), int(m.group(6)), int(m.group(3)), int(m.group(4)), int(m.group(5)), int(m.group(9) or 0)))
    names = demangle([r[1] for r in rows])
    with open(os.path.join(ROOT, "profiles", "r02_ptxas.txt"), "w") as f:
        f.write("nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -Xptxas -v   (per kernel: registers, stack frame, spill stores / loads, static smem)\n\n")
        for src, mangled, arch_, regs, stack, sst, sld, smem in rows:
            short = re.sub(r"\(.*", "", names[mangled].replace("dfx::", "").replace("void ", ""))
            f.write(f"{src:22s} {arch_}  regs={regs:3d}  stack={stack:4d}  spill_st={sst:4d}  spill_ld={sld:4d}  smem={smem:6d}  {short}\n")
    print(f"{len(kernels)} kernels, {len(rows)} ptxas rows")


if __name__ == "__main__":
    main()
