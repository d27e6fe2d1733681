"""Per-kernel opcode histogram of librst_sm100.so (cuobjdump -sass): the Blackwell-native evidence of B200_PROFILING.md."""
import collections, re, subprocess, sys
lib = sys.argv[1]
out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = pat.match(line)
    if m and kern:
        hist[kern][m.group(1)] += 1
cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMAPF", "LDTM", "UTCBAR", "HMMA", "FFMA", "total"]
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---|" * len(cols))
tot = collections.Counter()
for k, h in hist.items():
    def cnt(prefix, exact=False):
        return sum(v for op, v in h.items() if (op == prefix if exact else op.startswith(prefix)))
    row = [cnt("UTCHMMA") + cnt("UTCQMMA") - cnt("UTCHMMA.2CTA"), cnt("UTCHMMA.2CTA"), cnt("UTMALDG"), cnt("UTMAPF"), cnt("LDTM"),
           cnt("UTCBAR"), cnt("HMMA"), cnt("FFMA"), sum(h.values())]
    if row[0] + row[1] + row[2] + row[4] == 0 and "--all" not in sys.argv:
        continue
    for c, v in zip(cols, row):
        tot[c] += v
    print(f"| `{k}` | " + " | ".join(str(v) for v in row) + " |")
print("| **all tensor-core kernels** | " + " | ".join(str(tot[c]) for c in cols) + " |")
