"""SASS mnemonic counts per kernel: cuobjdump -sass msm_we_b200/libmsm_we_b200.so | python tools/sass_mnemonics.py [name filters...]"""
import sys, re, collections, subprocess
want = ["UTCHMMA", "UTCBAR", "STTM", "LDTM", "UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "DMMA", "DADD", "DMUL", "DFMA", "F2F", "LDS", "STS",
        "ATOMS", "ATOMG", "RED", "MATCH", "SHFL"]
keep = sys.argv[1:]
counts = collections.OrderedDict()
cur = None
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and m.group(1) in want:
        counts[cur][m.group(1)] += 1
mangled = list(counts)
dem = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.splitlines()
for mname, d in zip(mangled, dem):
    d = re.sub(r"\(.*", "", d)
    if (not keep or any(k in d for k in keep)) and counts[mname]:
        print(d)
        print("    " + "  ".join(f"{k}={counts[mname][k]}" for k in want if counts[mname][k]))
