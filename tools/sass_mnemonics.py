"""Per-kernel counts of the SASS mnemonics that prove what the built library uses (TMA, bulk copies, mbarriers,
packed f32x2 arithmetic, XU-pipe transcendentals, shared-memory atomics; tensor-core mnemonics would show up too).
  python tools/sass_mnemonics.py imagepipe_b200/libipb200.so > profiles/rNN_sass_mnemonics.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "imagepipe_b200/libipb200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
WANT = ("UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "MUFU", "ATOMS", "LDS", "STS", "LDG", "STG", "HMMA", "UTCMMA", "UTCHMMA", "TCGEN")
counts, total, order, cur, i = collections.defaultdict(collections.Counter), collections.Counter(), [], None, 0
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\(anonymous namespace\)::", "", names[i]).split("(")[0].replace("void ", "").replace("ipb::", "")
        cur += "" if cur not in order else ""
        i += 1
        if cur not in order:
            order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for w in WANT:
            if op.startswith(w):
                counts[cur][w] += 1
print(f"# {lib}: sm_100a SASS, instruction sites per kernel (static counts)")
print(f"{'kernel':58s} {'total':>6s} " + " ".join(f"{w:>7s}" for w in WANT))
for k in order:
    print(f"{k[:58]:58s} {total[k]:6d} " + " ".join(f"{counts[k][w]:7d}" for w in WANT))
