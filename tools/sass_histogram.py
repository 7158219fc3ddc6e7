"""SASS opcode histogram of every kernel object of libnrc_b200.so (profiles/sass_opcodes.txt): the mnemonics that prove
which hardware path a kernel takes - UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UBLKCP (bulk async
copies), SYNCS (mbarrier), HMMA (mma.sync), RED / ATOM (atomics), LDG / STG widths.
    python tools/sass_histogram.py > profiles/sass_opcodes.txt"""
import collections, glob, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "HMMA", "DFMA", "RED", "ATOMG", "ATOMS", "LDG.E.128", "LDG.E.64",
        "STG.E.128", "LDS", "STS", "LDSM", "BAR", "MUFU", "SHFL"]
for obj in sorted(glob.glob(os.path.join(ROOT, "neural_radiance_caching_b200", "build", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)[:90]
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            per[cur]["total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    per[cur][k] += 1
    print(f"== {os.path.basename(obj)}")
    for fn, c in per.items():
        if c["total"] == 0:
            continue
        print(f"  {fn}: {c['total']} instr | " + " ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
