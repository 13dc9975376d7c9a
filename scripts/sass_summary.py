#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of every object file in csrc/, the instruction count and the counts of the
mnemonics that prove the Blackwell paths (tcgen05 MMA / TMEM load / TMA / bulk copy / mbarrier ...), plus a short
excerpt around the first tcgen05 MMA of the main convolution kernel.
usage: python scripts/sass_summary.py > profiles/r02_sass_summary.md   (needs cuobjdump; no GPU)"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "object-detection-tensorflow_b200", "csrc")
KEY = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "UTMASTG", "UBLKPF", "ELECT",
       "MUFU", "SHFL", "REDUX", "ATOMG", "ATOMS", "HFMA2", "FFMA", "LDG", "STG", "LDS", "STS", "BAR"]


def main():
    print("# r02: SASS summary of the product kernels (cuobjdump -sass of csrc/*.o, sm_100a)\n")
    print("Counts are STATIC instructions.  `UTCHMMA` = tcgen05.mma, `UTCHMMA.2CTA` = cta_group::2, `UTMALDG` = TMA "
          "tensor load (`.IM2COL` = im2col mode), `UBLKCP` = cp.async.bulk, `LDTM` = tcgen05.ld, `UTCBAR` = "
          "tcgen05.commit, `SYNCS` = mbarrier.\n")
    print("| object | kernel | instr | " + " | ".join(KEY) + " |")
    print("|---|---|---|" + "---|" * len(KEY))
    excerpt = None
    for obj in sorted(glob.glob(os.path.join(CSRC, "*.o"))):
        txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        for part in re.split(r"\s+Function : ", txt)[1:]:
            name = part.split("\n")[0].strip()
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            short = re.sub(r"\(.*", "", dem).replace("odt::", "")
            lines = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(.*?);", part)
            ops = [re.sub(r"^@!?U?P\d+\s+", "", ln).split()[0] for ln in lines if ln.strip()]
            cnt = collections.Counter()
            for o in ops:
                base = o.split(".")[0]
                cnt[base] += 1
                if o.startswith("UTCHMMA.2CTA"):
                    cnt["UTCHMMA.2CTA"] += 1
            print("| %s | `%s` | %d | " % (os.path.basename(obj), short[:60], len(ops)) +
                  " | ".join(str(cnt.get(k, 0)) if cnt.get(k, 0) else "" for k in KEY) + " |")
            if excerpt is None and "conv_tc_kernel<1>" in dem:
                idx = next((i for i, ln in enumerate(lines) if "UTCHMMA" in ln), None)
                if idx is not None:
                    excerpt = (dem, lines[max(0, idx - 6): idx + 14])
            variants = sorted({o for o in ops if o.split(".")[0] in ("UTMALDG", "UTCHMMA", "LDTM", "UTCBAR", "UBLKCP")})
            if variants:
                sys.stderr.write("%s: %s\n" % (short[:50], " ".join(variants)))
    if excerpt:
        print("\n## Excerpt: first tcgen05.mma of `%s`\n\n```" % re.sub(r"\(.*", "", excerpt[0]))
        for ln in excerpt[1]:
            print("  " + ln)
        print("```")


if __name__ == "__main__":
    main()
