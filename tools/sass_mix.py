"""Static instruction mix of one kernel of the built library (cuobjdump -sass): opcode histogram grouped into
arithmetic / data movement / control, plus the longest straight-line region.  For the fully unrolled
factorisation kernels the static mix is close to the dynamic one per frame.
Usage: python tools/sass_mix.py <substring of the mangled name> [more substrings ...]"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "keypoint_moseq_b200", "libkpms_b200.so")
GROUPS = [
    ("fp32 fma (packed)", r"^FFMA2"), ("fp32 fma", r"^FFMA"), ("fp32 mul/add (packed)", r"^(FMUL2|FADD2)"),
    ("fp32 mul/add", r"^(FMUL|FADD)"), ("fp64", r"^(DFMA|DMUL|DADD|DMMA)"), ("special func", r"^MUFU"),
    ("shuffle", r"^SHFL"), ("shared load", r"^LDS"), ("shared store", r"^STS"), ("global load", r"^(LDG|LD\b)"),
    ("global store", r"^(STG|ST\b)"), ("async copy", r"^(LDGSTS|LDGDEPBAR|DEPBAR)"), ("local (spill)", r"^(LDL|STL)"),
    ("move / select", r"^(MOV|SEL|FSEL|PRMT|IMAD\.MOV|UMOV)"), ("integer / address", r"^(IMAD|IADD|LEA|SHF|LOP|IABS|I2F|F2I|ULEA|UIADD|UIMAD|ULOP|USHF|VIADD)"),
    ("predicate / compare", r"^(ISETP|FSETP|PLOP|P2R|R2P|FSET|DSETP|UISETP)"), ("barrier / sync", r"^(BAR|WARPSYNC|BSYNC|BSSY|NANOSLEEP|MEMBAR|ERRBAR|CCTL)"),
    ("branch", r"^(BRA|EXIT|RET|CALL|BRX|JMP)"), ("other", r".")]


def main():
    names = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    mangled = re.findall(r"Function (\S+):", names)
    for want in sys.argv[1:]:
        hits = [m for m in mangled if want in m]
        if not hits:
            print(f"no kernel matches {want}")
            continue
        fn = hits[0]
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, LIB], capture_output=True, text=True).stdout
        ops = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass, flags=re.M)
        demangled = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
        total = len(ops)
        short = demangled.split("(")[0]
        print(f"# {short}: {total} SASS instructions")
        counts = collections.OrderedDict((g, 0) for g, _ in GROUPS)
        detail = collections.Counter()
        for op in ops:
            for g, pat in GROUPS:
                if re.match(pat, op):
                    counts[g] += 1
                    break
            detail[op.split(".")[0]] += 1
        for g, c in counts.items():
            if c:
                print(f"{g:24s} {c:7d}  {100.0 * c / total:5.1f} %")
        print("top opcodes:", ", ".join(f"{k} {v}" for k, v in detail.most_common(14)))
        print()


if __name__ == "__main__":
    main()
