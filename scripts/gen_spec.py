"""Generate the mode-specialised build of the BSIM4 evaluator: a copy of the kernel sources in which the
model card's integer mode switches are replaced by the constants of one mode set, so that the compiler
drops every branch that this mode set never takes (smaller kernel image, denser instruction stream).
The launcher only uses this object for groups whose model cards all carry exactly that mode set.
usage: gen_spec.py <src_dir> <out_dir>"""
import os, re, shutil, sys

src, out = sys.argv[1], sys.argv[2]
# the mode set of a plain digital CMOS card (BSIM4 defaults with capMod = 2, dioMod = 1); dtype stays a variable
SPEC = dict(capMod=2, cvchargeMod=0, dioMod=1, gidlMod=0, igbMod=0, igcMod=0, lambdaGiven=0, mobMod=0, mtrlCompatMod=0,
            mtrlMod=0, pigcdGiven=0, rdsMod=0, rbodyMod=0, tempMod=0, tnoiMod=0, vtlGiven=0)
os.makedirs(out, exist_ok=True)
for f in os.listdir(src):
    if f.endswith((".h", ".cuh", ".def")) or f == "b4_kernels.cu":
        text = open(os.path.join(src, f)).read()
        if f.startswith("bsim4_") and f.endswith(".h") and f not in ("bsim4_types.h",):
            for k, v in SPEC.items():
                text = re.sub(r"\bM\.%s\b" % k, "(%d)" % v, text)
            text = re.sub(r"\bM\.versionDouble\b", "(4.82)", text)      # the specialised build is the 4.8.2 evaluator
        open(os.path.join(out, f), "w").write(text)
print(",".join("%s=%d" % kv for kv in sorted(SPEC.items())))
