"""profiles/r2_sass_tc.txt: per kernel of libbk_b200.so, how many tcgen05 / TMEM / TMA instructions its SASS holds
(B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor).  Runs on the CPU box:
    python scripts/r2/sass_listing.py > profiles/r2_sass_tc.txt"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = os.path.join(root, "bayes-kit_b200", "libbk_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names, counts, cur = [], collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); names.append(cur); continue
    if cur:
        for op in ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "DFMA", "DADD", "DMUL", "HMMA", "MUFU"):
            if re.search(r"\b" + op + r"\b|\b" + op + r"\.", line):
                counts[cur][op] += 1
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass bayes-kit_b200/libbk_b200.so -- {len(names)} kernels (sm_100a); instruction counts per kernel")
print("# UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit, SYNCS = mbarrier\n")
hdr = ["UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "SYNCS", "DFMA", "MUFU"]
print(f"{'kernel':100s} " + " ".join(f"{h:>8s}" for h in hdr))
for n, d in sorted(zip(names, dem), key=lambda t: -counts[t[0]]["UTCHMMA"]):
    c = counts[n]
    if c["UTCHMMA"] or c["UTMALDG"] or "smc" in d or "acf" in d or "sep_sampler<float, 8, 4, 0, 0" in d:
        short = re.sub(r"\(.*", "", d)[:100]
        print(f"{short:100s} " + " ".join(f"{c[h]:8d}" for h in hdr))
