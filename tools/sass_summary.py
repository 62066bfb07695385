"""Counts the Blackwell-specific SASS mnemonics of the built library (cuobjdump -sass) per kernel family -> profiles/.

    python tools/sass_summary.py [profiles/sass_r2_summary.json]

UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor loads / stores, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit -> mbarrier,
SYNCS = mbarrier ops, UTCATOMSWS = TMEM allocation; HMMA / IMMA would be legacy mma.sync (there must be none)."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "noisediff_b200", "csrc", "libnoisediff_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "IMMA", "LDGSTS",
         "MUFU.TANH", "MUFU.EX2", "ACQBULK", "UBLKCP"]


def main(dst):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    arch = set(re.findall(r"arch = (sm_\w+)", txt))
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"^void ", "", name).split("(")[0]
            cur = kernels.setdefault(name, collections.Counter())
            cur["_instructions"] += 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        cur["_instructions"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w.count(".") and op.startswith(w)):
                cur[w] += 1
    total = collections.Counter()
    fam = collections.OrderedDict()
    for k, c in kernels.items():
        base = k.split("<")[0]
        f = fam.setdefault(base, collections.Counter())
        f["_kernels"] += 1
        for kk, v in c.items():
            f[kk] += v
            total[kk] += v
    out = {"library": os.path.relpath(LIB, ROOT), "arch": sorted(arch), "n_kernels": len(kernels),
           "totals": {k: v for k, v in sorted(total.items()) if v},
           "per_kernel_family": {k: {kk: vv for kk, vv in sorted(v.items()) if vv} for k, v in fam.items()},
           "legacy_mma_sync_instructions": total["HMMA"] + total["IMMA"]}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["totals"]), "kernels:", len(kernels), "arch:", sorted(arch))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_r2_summary.json"))
