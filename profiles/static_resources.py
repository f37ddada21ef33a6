"""Static (no-GPU) evidence from the build: registers / spills / shared memory per kernel from the ptxas logs, and
counts of the Blackwell-specific SASS mnemonics in libair_b200.so (cuobjdump -sass).
    python profiles/static_resources.py > profiles/r1_static_resources.md"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "tf-attend-infer-repeat_b200", "csrc")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return [re.sub(r"\(.*$", "", o) for o in out]


def main():
    print("# Static build evidence (sm_100a; `nvcc -Xptxas -v` logs and `cuobjdump -sass libair_b200.so`)\n")
    print("## ptxas: registers / spills / static shared memory per kernel\n")
    print("| file | kernel | regs | spill st/ld (B) | smem (B) | barriers |")
    print("|---|---|---|---|---|---|")
    for log in sorted(glob.glob(os.path.join(CSRC, "build", "*.ptxas.log"))):
        txt = open(log).read()
        blocks = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'\n(?:.*\n)*?.*?(\d+) bytes stack frame, (\d+) bytes "
                            r"spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used (\d+) barriers)?"
                            r"(?:, (\d+) bytes smem)?", txt)
        names = demangle([b[0] for b in blocks])
        for b, n in zip(blocks, names):
            print(f"| {os.path.basename(log).replace('.ptxas.log', '.cu')} | `{n[:90]}` | {b[4]} | {b[2]}/{b[3]} | "
                  f"{b[6] or 0} | {b[5] or 0} |")
    print("\n## SASS mnemonics (whole library)\n")
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, "..", "libair_b200.so")], capture_output=True,
                          text=True).stdout
    fn, per = None, collections.defaultdict(collections.Counter)
    want = ("UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "UTCATOMSWS", "LDTM", "SYNCS", "CREDUX", "UTMACCTL", "ACQBULK",
            "UCGABAR", "HMMA", "FFMA", "MUFU")
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            for w in want:
                if op.startswith(w):
                    per[fn][w + ("" if "MULTICAST" not in op else ".MULTICAST")] += 1
    names = demangle(list(per))
    cols = want[:9]
    pdl = sum(1 for c in per.values() if c["ACQBULK"])
    print(f"UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = "
          f"cp.async.bulk, SYNCS = mbarrier ops, CREDUX = redux.sync min/max.  {pdl} kernels contain ACQBULK "
          f"(griddepcontrol.wait: programmatic dependent launch); kernels with none of the columns below are omitted.\n")
    print("| kernel | " + " | ".join(cols) + " | UTMALDG.MULTICAST |")
    print("|---|" + "---|" * (len(cols) + 1))
    for (fn, c), n in sorted(zip(per.items(), names), key=lambda t: t[1]):
        if any(c[w] for w in cols) or c["UTMALDG.MULTICAST"]:
            print(f"| `{n[:80]}` | " + " | ".join(str(c[w]) for w in cols) + f" | {c['UTMALDG.MULTICAST']} |")


if __name__ == "__main__":
    main()
