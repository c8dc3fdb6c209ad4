"""Static evidence, no GPU needed: for every kernel in csrc/libcmx_b200.so list registers / shared memory / spills
(`cuobjdump -res-usage`) and how many Blackwell-native SASS instructions it holds (`cuobjdump -sass`):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, HMMA = legacy mma.sync (must be 0),
LDGSTS = cp.async, ATOMS / REDG+ATOMG = shared / global atomics, FFMA2 = packed fp32 FMA.
Usage: python scripts/sass_summary.py > profiles/r1_sass_resource_summary.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "rgbx_semantic_segmentation_b200", "csrc", "libcmx_b200.so")
PATS = [("UTCMMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
        ("UBLKCP", r"\bUBLKCP"), ("HMMA", r"\bHMMA"), ("LDGSTS", r"\bLDGSTS"), ("ATOMS", r"\bATOMS"), ("REDG", r"\b(REDG|RED|ATOMG)\b"),
        ("FFMA2", r"\bFFMA2"), ("MUFU", r"\bMUFU")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(sig):
    sig = re.sub(r"^void ", "", sig)
    m = re.match(r"([A-Za-z0-9_:]+(?:<[^(]*>)?)\(", sig)
    return m.group(1) if m else sig[:60]


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL|CONSTANT\[0\])\s*:\s*(\d+)", line)}
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    ninstr = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or "/*" not in line:
            continue
        if re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            ninstr[cur] += 1
            for name, pat in PATS:
                if re.search(pat, line):
                    counts[cur][name] += 1
    names = sorted(set(usage) | set(ninstr))
    dm = demangle(names)
    print("# %s" % __doc__.strip().replace("\n", "\n# "))
    print("# %d kernels; static shared memory only (dynamic shared memory is set at launch)" % len(names))
    hdr = ["kernel", "regs", "stack", "smem_static", "instrs"] + [n for n, _ in PATS]
    print(",".join(hdr))
    rows = []
    for n in names:
        u = usage.get(n, {})
        rows.append(['"%s"' % short(dm.get(n, n)), u.get("REG", ""), u.get("STACK", ""), u.get("SHARED", ""), ninstr.get(n, 0)] +
                    [counts[n].get(p, 0) for p, _ in PATS])
    rows.sort(key=lambda r: (-int(r[5] or 0), -int(r[7] or 0), r[0]))
    for r in rows:
        print(",".join(str(x) for x in r))
    tot = collections.Counter()
    for n in names:
        tot.update(counts[n])
    print("# totals: " + ", ".join("%s=%d" % (p, tot[p]) for p, _ in PATS))
    print("# kernels with stack frames (possible spills): %s" %
          ([short(dm[n]) for n in names if usage.get(n, {}).get("STACK", 0) > 0] or "none"))
    hm = [short(dm[n]) for n in names if counts[n]["HMMA"]]
    print("# kernels holding legacy HMMA (mma.sync): %s" % (hm or "none"))
    if set(hm) - {"gemm_wmma_kernel"}:   # the generic strided fallback for tcgen05-ineligible shapes is the only allowed one
        sys.exit("legacy HMMA outside the generic fallback GEMM")


if __name__ == "__main__":
    main()
