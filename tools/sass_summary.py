"""Mnemonic counts per kernel from `cuobjdump -sass` of the in-tree library -> profiles/r2_sass_summary.txt (no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pnpinversion_b200", "libpnpinv.so")
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "HMMA.16816", "MUFU.EX2", "MUFU.RCP", "SYNCS", "ELECT",
        "STG.E.ENL2.256", "LDG.E.ENL2.256", "LDGSTS", "UTCATOMSWS"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_summary.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    funcs, cur = [], None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = {"n": 0, "c": collections.Counter()}
            funcs.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["n"] += 1
        for k in KEYS:
            if op.startswith(k):
                cur["c"][k] += 1
                break
    rows = sorted(zip(names, funcs), key=lambda t: -t[1]["n"])
    with open(out, "w") as f:
        f.write("# profiles/r2_sass_summary.txt -- `cuobjdump -sass pnpinversion_b200/libpnpinv.so` (sm_100a), mnemonic counts per kernel\n"
                "# (tools/sass_summary.py).  UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (tensor memory), UTMALDG = cp.async.bulk.tensor (TMA),\n"
                "# UTCBAR = tcgen05.commit, HMMA.16816 = mma.sync.m16n8k16 (legacy tensor-core path of the <= 1024-token / cross-attention\n"
                "# kernels), MUFU.EX2 = exp2, LDGSTS = cp.async\n\n")
        tot = collections.Counter()
        for name, fn in rows:
            tot.update(fn["c"])
            f.write(f"{name[:110]:110s} instr {fn['n']:6d}  " + "  ".join(f"{k} {v}" for k, v in fn["c"].items()) + "\n")
        f.write("\nTOTAL  " + "  ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]) + "\n")
    print("wrote", out, "kernels:", len(rows))


if __name__ == "__main__":
    main()
