"""Measurement aid: the SASS excerpt committed as profiles/r02_sass_excerpt.txt — the instructions that prove what the kernels
use (TMA bulk copies, mbarrier, bulk L2 prefetch, shared-memory and global 64-bit atomics, the lean projection's FP64 pipe) from
`cuobjdump -sass slicer_b200/_build/libslicer_b200.so`, plus instruction counts over the whole library.
usage: python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt"""
import collections, os, re, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "slicer_b200", "_build", "libslicer_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout.splitlines()
funcs, cur = collections.OrderedDict(), None
for l in sass:
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        funcs[cur].append(l.rstrip())


def excerpt(name_re, pat, limit, title):
    name = next(f for f in funcs if re.search(name_re, f))
    print(f"# {title}")
    print(f"\t\tFunction : {name}")
    n = 0
    for l in funcs[name]:
        if re.search(pat, l):
            print(l)
            n += 1
            if n >= limit:
                break
    print()


print("# SASS evidence (cuobjdump -sass slicer_b200/_build/libslicer_b200.so, sm_100a), round 2; made by tools/sass_excerpt.py")
excerpt(r"deposit_pipelined_kernelILi0ELi0ELb1ELi4E", r"UBLKCP|SYNCS|ATOMS|MUFU|F2F\.F64|DFMA|STG\.E", 44,
        "K1 = pipe::deposit_pipelined_kernel<TSC, AOS, SINGLE, PATH_EMIT_LEAN>: TMA bulk copies (UBLKCP), mbarrier (SYNCS.*), the lean\n"
        "# projection (MUFU.RSQ/RCP seeds, F2F conversions, DFMA), shared-memory atomics (ATOMS), record stores (STG)")
excerpt(r"bin_scatter_kernelILi4096ELb0ELb0", r"UBLKPF|LDG\.E\.128|ATOMS|STG\.E\.64|BAR\.SYNC", 24,
        "K2d = binned::bin_scatter_kernel<4096, false, false, ...>: paired loads (LDG.E.128), bulk L2 prefetch of the next batch (UBLKPF.L2),\n"
        "# rank atomics (ATOMS.ADD), sorted write-out (STG.E.64)")
excerpt(r"tile_deposit_kernelILi0E", r"ATOMS|REDG|IADD3\.X|F2I\.S64", 40,
        "K3 = binned::tile_deposit_kernel<TSC>: contribution -> int64 (F2I.S64), shared-memory limb atomics (ATOMS.ADD: low limb with its\n"
        "# old value, carry by IADD3 / IADD3.X, high limb fire-and-forget) and the 64-bit flush (REDG.E.ADD.64)")
print("# instruction counts over the whole library:")
cnt = collections.Counter()
for body in funcs.values():
    for l in body:
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m and re.match(r"ATOMS\.ADD|REDG|SYNCS|UBLKCP|UBLKPF", m.group(1)):
            cnt[m.group(1)] += 1
for k in sorted(cnt):
    print(f"{cnt[k]:7d} {k}")
print("# (no UTC*MMA / HMMA: the path is streaming + scatter, not a contraction)")
