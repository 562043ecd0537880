"""Mnemonic counts per kernel from `cuobjdump -sass libsatk.so` (profiles/r0N_sass_evidence.md): which kernels really use the
tensor-core / TMEM / TMA / bulk-copy / mbarrier-transaction / DSMEM machinery.  `python tools/sass_evidence.py > profiles/...md`."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "self-attention-tacotron_b200", "libsatk.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
MN = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "UBLKCP", "SYNCS", "UTCBAR", "MUFU.EX2", "MUFU.RCP", "REDUX", "STAS",
      "ACQBULK", "PREEXIT"]
rows, cur, k = collections.OrderedDict(), None, -1
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        k += 1
        full = names[k] if k < len(names) else m.group(1)
        full = re.sub(r"\((?:int|bool)\)", "", full)                 # "<(int)5, (int)3>" -> "<5, 3>"
        cur = full[:full.rindex(">(") + 1] if ">(" in full else full.split("(")[0]
        rows[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for mn in MN:
        if re.search(r"\b" + re.escape(mn) + r"\b", line):
            rows[cur][mn] += 1
print("# SASS evidence of the Blackwell-native paths in libsatk.so\n")
print("`cuobjdump -sass self-attention-tacotron_b200/libsatk.so` (`tools/sass_evidence.py`), mnemonic counts per kernel (only kernels that use the "
      "tensor-core / TMEM / TMA / bulk-copy / mbarrier-transaction / DSMEM machinery are listed).  UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG / "
      "UTMAREDG = TMA tensor load / store / reduce-add, LDTM / STTM = tcgen05.ld / tcgen05.st (tensor memory), UBLKCP = cp.async.bulk, SYNCS = "
      "mbarrier transaction ops, STAS = st.async (DSMEM store completing a remote mbarrier), ACQBULK / PREEXIT = griddepcontrol.wait / "
      ".launch_dependents (programmatic dependent launch).\n")
print("| kernel | " + " | ".join(MN) + " |\n|---|" + "---:|" * len(MN))
tot = collections.Counter()
for n, c in rows.items():
    tot.update(c)
    if any(c[m] for m in MN if m not in ("MUFU.EX2", "MUFU.RCP", "REDUX")):
        print(f"| `{n[:80]}` | " + " | ".join(str(c[m]) for m in MN) + " |")
print("| **whole library** | " + " | ".join(str(tot[m]) for m in MN) + " |")
