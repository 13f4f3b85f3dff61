"""Generate the lock-step build of the BSIM4 evaluator: a copy of the kernel sources in which a block-wide
barrier (XB_SYNC_POINT) is inserted at statement boundaries that every thread of a block reaches and that are
at least `gap` source lines after the previous barrier.  "Every thread reaches" = brace depth 1 of the stage
functions, or nested only inside if/else chains whose conditions read nothing but the model card (M.), the
bin record (P.), the solver flags (S.) or `charge_needed` -- values that are uniform per block in the
uniform-record kernel, the only kernel the lock-step object contains.
usage: gen_lockstep.py <src_dir> <out_dir> [gap]"""
import os, re, shutil, sys

src, out = sys.argv[1], sys.argv[2]
gap = int(sys.argv[3]) if len(sys.argv) > 3 else 8
os.makedirs(out, exist_ok=True)
for f in os.listdir(src):
    if f.endswith((".h", ".cuh", ".def")) or f == "b4_kernels.cu":
        shutil.copy(os.path.join(src, f), os.path.join(out, f))
FUNCS = ("stage_dc", "stage_cv", "stage_caps", "stage_fvars", "emit_vectors", "emit_matrices")


def uniform_cond(c):
    c = re.sub(r"\b[MPS]\.\w+", "", c)
    c = re.sub(r"\bcharge_needed\b", "", c)
    c = re.sub(r"\bk[A-Z]\w*", "", c)                      # named constants
    c = re.sub(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", "", c)  # literals
    return re.fullmatch(r"[\s()!&|<>=]*", c) is not None


total = 0
for f in ("bsim4_eval_dc.h", "bsim4_eval_cv.h", "bsim4_load.h"):
    lines = open(os.path.join(src, f)).read().split("\n")
    res, infunc, since, paren = [], False, 0, 0
    stack = []          # one entry per open brace: True = block reached uniformly
    chain = {}          # depth -> is the current if/else chain at this depth uniform so far
    for l in lines:
        code = l.split("//")[0]
        starts = any(re.match(r"XB_HD void %s\(" % fn, l) for fn in FUNCS)
        if starts:
            infunc, stack, since, paren, chain = True, [], 0, 0, {}
        if infunc and stack and all(stack) and paren == 0 and since >= gap:
            prev = res[-1].split("//")[0].rstrip() if res else ""
            nxt = code.strip()
            ok = (prev.endswith(";") or prev.endswith("}")) and nxt != "" and not nxt.startswith(
                ("else", "}", ")", "&&", "||", "+", "-", "*", "/", "?", ":", "#", "case", "default", "break")) and "XB_SYNC_POINT" not in prev
            if ok:
                res.append("  " * len(stack) + "XB_SYNC_POINT(1);")
                since = 0
                total += 1
        res.append(l)
        if not infunc:
            continue
        since += 1
        paren += code.count("(") - code.count(")")
        # brace bookkeeping, character by character so that "} else {" pops then pushes
        stripped = code.strip()
        m_if = re.match(r"^(\} else )?if \((.*)\) \{$", stripped)
        m_else = re.match(r"^\} else \{$", stripped)
        opens_multi = stripped.endswith("{") and code.count("{") - code.count("}") >= (0 if stripped.startswith("}") else 1)
        for ch in code:
            if ch == "}":
                if stack:
                    stack.pop()
                if not stack and not starts:
                    infunc = False
            elif ch == "{":
                d = len(stack)
                if not stack:
                    stack.append(True)            # function body
                elif m_if and opens_multi and ch == code.rstrip()[-1:] and code.rstrip().endswith("{") and code.rfind("{") == code.rstrip().rfind("{") and code.index("{", 0) >= 0 and code.count("{") == 1:
                    u = uniform_cond(m_if.group(2)) and (chain.get(d, True) if m_if.group(1) else True)
                    chain[d] = u
                    stack.append(u)
                elif m_else and code.count("{") == 1:
                    stack.append(chain.get(d, False))
                else:
                    chain[d] = False
                    stack.append(False)
    open(os.path.join(out, f), "w").write("\n".join(res))
    print(f, sum(1 for x in res if "XB_SYNC_POINT" in x))
print("inserted", total, "sync points")
