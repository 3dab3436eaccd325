"""profiles/sass_digest.txt: per-kernel SASS instruction counts of the shipped libvacv_cuda.so (cuobjdump -sass), the mnemonics that
prove which hardware paths a kernel uses: UBLKCP (cp.async.bulk, 1-D TMA), UTMALDG (tensor-map TMA loads), SYNCS (mbarrier),
LDGSTS (cp.async), IDP (dp2a / dp4a), FFMA2 (packed fp32), VIADDMNMX / VIMNMX3 (DPX), plus total instructions and registers.
    python profiles/_sass_digest.py > profiles/sass_digest.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "arm-neon-opencv_b200", "libvacv_cuda.so")
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "IDP", "FFMA2", "VIADDMNMX", "VIMNMX3", "PRMT", "LDS", "STS", "LDG", "STG", "BSSY"]


def main():
    sass = subprocess.check_output(["cuobjdump", "-sass", LIB], text=True)
    res = subprocess.check_output(["cuobjdump", "-res-usage", LIB], text=True)
    regs = {}
    name = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and name:
            regs[name] = int(m.group(1))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur["total"] += 1
            op = m.group(1)
            if op in WATCH:
                cur[op] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    arch = sorted(set(re.findall(r"sm_\d+a?", subprocess.check_output(["cuobjdump", "--list-elf", LIB], text=True))))
    print(f"# SASS digest of arm-neon-opencv_b200/libvacv_cuda.so ({len(kernels)} kernels, architectures {arch}); counts are static instruction counts")
    print("# " + "\t".join(["total", "regs"] + WATCH + ["kernel"]))
    tot = collections.Counter()
    for (mangled, c), nice in zip(kernels.items(), demangle):
        nice = re.sub(r"\(.*$", "", nice.replace("vacv::", ""))[:120]
        print("\t".join([str(c["total"]), str(regs.get(mangled, "?"))] + [str(c[w]) for w in WATCH] + [nice]))
        tot.update(c)
    print("# TOTAL\t" + "\t".join(f"{w}={tot[w]}" for w in WATCH))


if __name__ == "__main__":
    main()
