"""Generate the mode-specialised builds of the BSIM4 evaluator: for every mode tuple of
xyce_b200/csrc/bsim4_spec_tuples.def a copy of the kernel sources in which the model card's integer mode switches are
replaced by that tuple's constants, so that the compiler drops every branch the tuple never takes (smaller kernel image,
denser instruction stream).  The launcher uses an object only for groups whose model cards all carry exactly its tuple.
usage: gen_spec.py <src_dir> <out_dir_prefix> [id]      (out dir = prefix for id 0, prefix + id otherwise; no id = all)"""
import os, re, sys

src, prefix = sys.argv[1], sys.argv[2]
only = int(sys.argv[3]) if len(sys.argv) > 3 else None
NAMES = ["capMod", "cvchargeMod", "dioMod", "dtype", "gidlMod", "igbMod", "igcMod", "lambdaGiven", "mobMod", "mtrlCompatMod",
         "mtrlMod", "pigcdGiven", "rdsMod", "rbodyMod", "tempMod", "tnoiMod", "vtlGiven"]
tuples = []
for m in re.finditer(r"^\s*X\(([-0-9, ]+)\)", open(os.path.join(src, "bsim4_spec_tuples.def")).read(), re.M):
    v = [int(t) for t in m.group(1).split(",")]
    tuples.append((v[0], dict(zip(NAMES, v[1:]))))
for sid, spec in tuples:
    if only is not None and sid != only:
        continue
    out = prefix if sid == 0 else "%s%d" % (prefix, sid)
    os.makedirs(out, exist_ok=True)
    for f in os.listdir(src):
        if f.endswith((".h", ".cuh", ".def")) or f == "b4_kernels.cu":
            text = open(os.path.join(src, f)).read()
            if f.startswith("bsim4_") and f.endswith(".h") and f not in ("bsim4_types.h",):
                for k, v in spec.items():
                    if v != -2:
                        text = re.sub(r"\bM\.%s\b" % k, "(%d)" % v, text)
                text = re.sub(r"\bM\.versionDouble\b", "(4.82)", text)      # the specialised builds are the 4.8.2 evaluator
            open(os.path.join(out, f), "w").write(text)
    print(sid, ",".join("%s=%d" % kv for kv in sorted(spec.items()) if kv[1] != -2))
