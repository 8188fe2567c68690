"""SASS evidence of the Blackwell-native paths: per object of the in-tree build, how many tcgen05 / TMEM / bulk-copy
instructions its sm_100a SASS contains.  Runs without a GPU (cuobjdump on dsvt-ai-trt_b200/lib/obj/*.o).

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG",
             "SYNCS.ARRIVE", "SYNCS.PHASECHK", "FFMA2", "HMMA", "REDG", "ATOMG"]


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "dsvt-ai-trt_b200", "lib", "obj", "*.o")))
    print("# cuobjdump -sass of the in-tree sm_100a objects: instruction counts per mnemonic (whole object, all kernels)")
    print("# UTCHMMA = tcgen05.mma kind::f16/tf32, LDTM / STTM = tcgen05.ld / st (TMEM), UBLKCP = cp.async.bulk (TMA engine, no")
    print("# tensor map), UBLKPF = cp.async.bulk.prefetch.L2, SYNCS.* = mbarrier, FFMA2 = packed FP32 FMA; UTMALDG / UTMASTG")
    print("# (tensor-map TMA) are not used: every bulk copy here is a contiguous 1-D block")
    print(f"{'object':28s} " + " ".join(f"{m:>14s}" for m in MNEMONICS))
    for o in objs:
        if os.path.basename(o).startswith("plugin_"):
            continue
        sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
        counts = [len(re.findall(r"\b" + re.escape(m) + r"\b", sass)) for m in MNEMONICS]
        if any(counts):
            print(f"{os.path.basename(o):28s} " + " ".join(f"{c:14d}" for c in counts))
    kern = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "dsvt-ai-trt_b200", "lib", "obj", "attention_split.o")],
                          capture_output=True, text=True).stdout
    print("\n# kernels of attention_split.o with tcgen05.mma:")
    cur = None
    per = {}
    for line in kern.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
        elif cur and "UTCHMMA" in line:
            per[cur] = per.get(cur, 0) + 1
    for k, v in per.items():
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        print(f"{v:5d} UTCHMMA  {name[:150]}")


if __name__ == "__main__":
    main()
