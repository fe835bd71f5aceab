"""Per-kernel SASS opcode census of libpf_b200.so (no GPU needed):
  python profiles/sass_census.py > profiles/sass_census.txt
Counts the mnemonics that prove Blackwell-native code (UTC*MMA = tcgen05.mma, UTMALDG = TMA loads, LDTM = tcgen05.ld,
FFMA2 = packed fp32) next to the atomics (RED / ATOM / ATOMS) and the legacy tensor path (HMMA, none expected)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "panoptic_forecasting_b200", "libpf_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "FFMA2", "HMMA", "REDG", "RED", "ATOMG", "ATOM", "ATOMS",
         "SYNCS", "UBLKCP", "MUFU", "SHFL", "VOTE", "LDGSTS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur["_total"] += 1
            cur[m.group(1)] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    print("SASS census of panoptic_forecasting_b200/libpf_b200.so (sm_100a), instructions per kernel")
    print("%-72s %7s  %s" % ("kernel", "instrs", "watched opcodes"))
    tot = collections.Counter()
    for (name, c), dn in zip(per.items(), demangled):
        dn = re.sub(r"\(.*", "", dn).replace("pf::", "")
        w = ", ".join("%s %d" % (k, c[k]) for k in WATCH if c[k])
        print("%-72s %7d  %s" % (dn[:72], c["_total"], w))
        tot.update({k: c[k] for k in WATCH})
    print()
    print("whole library: " + ", ".join("%s %d" % (k, tot[k]) for k in WATCH if tot[k]))
    print("HMMA (legacy mma.sync path): %d" % tot["HMMA"])


if __name__ == "__main__":
    main()
