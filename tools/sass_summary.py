"""SASS evidence for the Blackwell-native kernels: per kernel of csrc/libvfr.so, the counts of the instructions that prove
tcgen05 / TMEM / TMA use (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = bulk copy,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, UTCATOMSWS = TMEM alloc) plus a short excerpt around the first MMA.
Runs on the build machine (cuobjdump, no GPU):  python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "video-fragments-retrieval_b200", "csrc", "libvfr.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMAPF", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "REDUX", "FMNMX3", "HMMA", "FFMA", "MUFU"]

def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur is not None and "/*" in line:
        funcs[cur].append(line)
names = demangle(list(funcs))
print(f"# {os.path.relpath(LIB, ROOT)}: {len(funcs)} kernels, sm_100a SASS (cuobjdump -sass)")
print("# kernel | instructions | " + " ".join(KEYS))
tensor_kernels = []
for f, lines in funcs.items():
    ops = collections.Counter()
    for l in lines:
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            ops[m.group(1).split(".")[0]] += 1
    counts = [sum(v for k, v in ops.items() if k.startswith(key)) for key in KEYS]
    short = re.sub(r"\(.*", "", names[f])
    print(f"{short} | {sum(ops.values())} | " + " ".join(str(c) for c in counts))
    if counts[0] or counts[1]:
        tensor_kernels.append((short, lines))
print()
for short, lines in tensor_kernels:
    idx = next(i for i, l in enumerate(lines) if "UTCHMMA" in l or "UTCQMMA" in l)
    print(f"## {short}: around the first tensor-core instruction")
    for l in lines[max(0, idx - 6):idx + 8]:
        print("   " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l.rstrip()))
    tm = [l for l in lines if "UTMALDG" in l][:2] + [l for l in lines if "LDTM" in l][:2]
    for l in tm:
        print("   " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l.rstrip()))
    print()
