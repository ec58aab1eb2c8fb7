"""SASS evidence for the Blackwell-native instructions of the shipped library (no GPU needed):

    python scripts/sass_summary.py > profiles/r2_sass_summary.txt

Extracts the sm_100a cubins from proxmin_b200/libproxmin_b200.so with cuobjdump and counts, per kernel, the mnemonics
that prove tcgen05 / TMA / TMEM use (B200_PROFILING.md): UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor load), LDTM / STTM
(tcgen05.ld / st), UTCBAR (tcgen05.commit), SYNCS (mbarrier), USETMAXREG (setmaxnreg), plus the memory instructions."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "proxmin_b200", "libproxmin_b200.so")
KEYS = ["UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "USETMAXREG", "ELECT", "F2FP", "LDG",
        "STG", "RED", "LDS", "STS", "FFMA2", "FFMA", "DFMA", "MUFU", "BAR"]


def main():
    cubins = sys.argv[1:] or ["grad_umma", "pgm_tail", "admm"]
    with tempfile.TemporaryDirectory() as tmp:
        for cb in cubins:
            name = cb + ".sm_100a.cubin"
            subprocess.run(["cuobjdump", "-xelf", name, LIB], cwd=tmp, check=True, capture_output=True)
            sass = subprocess.run(["cuobjdump", "-sass", os.path.join(tmp, name)], capture_output=True, text=True).stdout
            print("== %s" % name)
            for f in re.split(r"\n\s*Function : ", sass)[1:]:
                fname = f.split("\n", 1)[0].strip()
                demangled = subprocess.run(["c++filt", fname], capture_output=True, text=True).stdout.strip()
                ins = re.findall(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", f)
                c = collections.Counter(ins)
                short = re.sub(r"\(anonymous namespace\)::", "", demangled)
                short = re.sub(r"\(.*", "", short)
                print("%-46s %6d instructions  %s" % (short[:46], len(ins), "  ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])))
            print()


if __name__ == "__main__":
    main()
