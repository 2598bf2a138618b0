"""Per-kernel count of the SASS mnemonics that prove which Blackwell units a kernel drives (B200_PROFILING.md):
UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld from tensor memory), UTMALDG (TMA tensor loads), UBLKCP (bulk copies),
SYNCS.* (mbarriers), UTCBAR (tcgen05.commit), DMMA (fp64 tensor tiles), LDGSTS (cp.async).  Runs without a GPU:

    python tools/sass_summary.py admm_b200/libb200admm.so > profiles/<round>_sass_mnemonics.txt
"""
import collections
import re
import subprocess
import sys

PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCOMMA|UTCIMMA|UTCBAR|UTCATOMSWS|UTCCP|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|SYNCS|DMMA|HMMA|LDGSTS|REDUX|USETMAXREG|ELECT)\b")


def main(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    demangle = lambda names: subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    kernels, cur = collections.OrderedDict(), None
    arch = set(re.findall(r"arch = (sm_\w+)", out))
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = PAT.search(line)
        if m:
            kernels[cur][m.group(1)] += 1
        if re.search(r"/\*[0-9a-f]{4}\*/", line):
            kernels[cur]["_instructions"] += 1
    names = demangle(list(kernels))
    print("# %s: %d kernels, arch %s" % (path, len(kernels), ", ".join(sorted(arch))))
    print("# kernel | SASS instructions | Blackwell-unit mnemonics")
    for (k, c), nm in zip(kernels.items(), names):
        nm = re.sub(r"\(.*", "", nm)
        tags = ", ".join("%s x%d" % (t, n) for t, n in sorted(c.items()) if not t.startswith("_"))
        print("%-70s %7d  %s" % (nm[:70], c["_instructions"], tags or "-"))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "admm_b200/libb200admm.so")
