"""ADMS -> CUDA: translate the C++ that admsXml emits with Xyce's `_nosac` templates
(utils/ADMS/xyceImplementationFile_nosac.xml: probe derivatives propagated by hand, no Sacado) into a
single-source evaluator of the shape the small-device kernels of xyce_b200 consume (simple_kernels.cu).

What the generated C++ of every such model looks like (reference example
src/DeviceModelPKG/ADMS/N_DEV_ADMSmvs_2_0_0_etsoi.C):
  * header: `static const int admsNodeID_x / admsBRA_ID_x_y / admsProbeID_V_x_y = k;`, the model-card and
    instance members between `Begin verilog ... Variables` markers;
  * constructor: `jacobianElements.push_back(IntPair(row, col))` = the Jacobian stamp;
  * Instance::updateIntermediateVars: local declarations, probe extraction from the solution vector, then the
    analog block as straight C statements that fill staticContributions[node], dynamicContributions[node] and
    d_*Contributions[node][probe];
  * loadDAEFVector / loadDAEQVector / loadDAEdFdx / loadDAEdQdx: pure copies of those arrays onto the unknowns
    and the stamp entries.
The translator keeps the arithmetic statements verbatim (same operations in the same order as the reference
object, so results agree to the last bits) and rewrites only the plumbing around them:
  model_.X            -> rec.f[k] (flat per-instance record; k = position of "M:X" in the field list)
  instance member X   -> rec.f[k] ("I:X"; admsTemperature / adms_vt_nom included)
  (*solVectorPtr)[li] -> V[node]  (node voltages gathered by the kernel)
  d_probeVars[p][p]   -> 1.0
  loads               -> o.F / o.Q rows and o.JF / o.JQ slots in constructor order
$limit (limited probes, Jdxp correction terms, origFlag), analog functions, node collapsing, given() / $port_connected
tests and output variables are handled (see emit()); noise contributions are dropped.

usage: python -m xyce_b200.adms.translate <N_DEV_ADMSname.C> <N_DEV_ADMSname.h> <out.h> [name]
       python -m xyce_b200.adms.translate --all <dir with N_DEV_ADMS*.C/.h> <out_dir> name...
"""
import os
import re
import sys


class Unsupported(Exception):
    pass


def _section(text, begin, end):
    i = text.index(begin)
    j = text.index(end, i)
    return text[i:j]


def _function_body(text, signature):
    """Text of the function whose definition starts with `signature` (brace matched)."""
    i = text.index(signature)
    j = text.index("{", i)
    depth, k = 0, j
    while True:
        c = text[k]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return text[j + 1:k]
        k += 1


def _namespace_block(text, name):
    """Body of `namespace <name> { ... }` (first occurrence, brace matched), or None."""
    m = re.search(r"namespace\s+%s\s*\{" % re.escape(name), text)
    if not m:
        return None
    depth, k = 0, m.end() - 1
    while True:
        c = text[k]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return text[m.end():k]
        k += 1


def _analog_functions(C, H):
    """Analog functions of the Verilog-A source: admsXml writes each as a value function, a derivative function and an
    Evaluator class (declared in the header, defined in the .C, both inside namespace AnalogFunctions).  They are carried
    over whole: `double` -> real, every function and member marked host / device inline."""
    decl, defn = _namespace_block(H, "AnalogFunctions"), _namespace_block(C, "AnalogFunctions")
    if decl is None or defn is None:
        return ""
    def common(t):
        t = re.sub(r"/\*.*?\*/", "", t, flags=re.S)
        t = re.sub(r"//[^\n]*", "", t)
        t = re.sub(r"\bstd::(exp|log|sqrt|pow|fabs|tanh|sinh|cosh|atan|sin|cos|tan|log10|abs)\b", r"\1", t)
        t = t.replace("std::max", "adms_max").replace("std::min", "adms_min")
        t = re.sub(r"static_cast<double>\(([^()]*)\)", r"real(\1)", t)
        t = re.sub(r"\bpow\(", "rpow(", t)
        return re.sub(r"\bdouble\b", "real", t)
    decl, defn = common(decl), common(defn)
    # declarations: free functions and members (lines ending in ");" that are not statements inside a body)
    out_decl, depth = [], 0
    for line in decl.split("\n"):
        st = line.strip()
        is_fn = re.match(r"^[\w:<>&\s\*]*\b[\w~]+\s*\([^;{}]*\)\s*(const)?\s*;$", st) is not None and not st.startswith(("return", "typedef"))
        out_decl.append(("XB_HD " + st) if is_fn else line)
    # definitions: a signature is a non-blank line at depth 0 that is followed by "{"
    lines = defn.split("\n")
    out_defn, depth = [], 0
    for i, line in enumerate(lines):
        st = line.strip()
        if depth == 0 and st and not st.startswith(("{", "}", "#")) and st.endswith(")"):
            j = i + 1
            while j < len(lines) and not lines[j].strip():
                j += 1
            if j < len(lines) and lines[j].strip().startswith("{"):
                line = "XB_HD " + st
        out_defn.append(line)
        depth += line.count("{") - line.count("}")
    return "namespace AnalogFunctions {\n" + "\n".join(out_decl) + "\n" + "\n".join(out_defn) + "\n}  // namespace AnalogFunctions\n"


def _members(section):
    """`double X;` / `int X;` / `bool X;` declarations of a header section (Given flags excluded)."""
    out = []
    for m in re.finditer(r"^\s*(double|int|bool)\s+(\w+)\s*;", section, re.M):
        if not m.group(2).endswith("Given"):
            out.append((m.group(2), m.group(1)))
    return out


def parse(cfile, hfile):
    C = open(cfile).read()
    H = open(hfile).read()
    info = {}
    m = re.search(r"namespace (ADMS\w+)\s*\{", C)
    info["namespace"] = m.group(1)
    ids = {}
    for m in re.finditer(r"static const int (adms(?:NodeID|BRA_ID|ProbeID)_\w+)\s*=\s*([-+0-9 ]+);", H):
        ids[m.group(1)] = eval(m.group(2))
    info["ids"] = ids
    unknowns = sorted((v, k) for k, v in ids.items() if (k.startswith("admsNodeID_") or k.startswith("admsBRA_ID_")) and v >= 0)
    if [v for v, _ in unknowns] != list(range(len(unknowns))):
        raise Unsupported("node / branch ids are not 0..n-1: %r" % unknowns)
    info["unknowns"] = [k for _, k in unknowns]
    probes = sorted((v, k) for k, v in ids.items() if k.startswith("admsProbeID_"))
    info["probes"] = [k for _, k in probes]
    info["n_ext"] = int(re.search(r"numExtVars\s*=\s*(\d+);", C).group(1))
    # record fields
    inst = _members(_section(H, "// Begin verilog Instance Variables", "// end verilog Instance Variables"))
    model = _members(_section(H, "// Begin verilog Model Variables", "// end verilog model variables"))
    info["inst_members"] = inst + [("admsTemperature", "double"), ("adms_vt_nom", "double")]
    info["model_members"] = model
    # Jacobian stamp in constructor order
    stamp = re.findall(r"jacobianElements\.push_back\(IntPair\((\w+),(\w+)\)\);", C)
    info["stamp"] = [(ids[r], ids[c]) for r, c in stamp]
    # Node collapsing (collapseNode_x set in Instance::collapseNodes from the parameters): the generated registerLIDs
    # aliases li_x to the node it collapses onto (or -1 = ground), which is exactly what the gather map needs -- the
    # evaluator keeps all its unknowns and the aliased rows / entries sum up.  The flags themselves are instance members
    # that the analog block and the loads may test.
    info["inst_members"] += [(m, "bool") for m in sorted(set(re.findall(r"\bbool\s+(collapseNode_\w+)\s*;", H)))]
    # analog block
    fn = _function_body(C, "bool Instance::updateIntermediateVars()")
    a = fn.index("// Local variables")
    dbg = [m.start() for m in re.finditer(r"if \(DEBUG_DEVICE", fn)]
    b = [p for p in dbg if p > a][0]
    body = fn[a:b]
    s0 = body.index("// set the sizes of the Fad arrays")
    s1 = body.index("// extract solution variables")
    info["locals_text"] = body[:s0]
    info["analog_text"] = re.sub(r"//[^\n]*", "", body[s1:])      # the Verilog-A source echoed in comments is not code
    # $limit: limited probes.  The generated code keeps the limited voltages (X_limited / X_old / X_orig), the correction
    # terms (probeDiffs, Jdxp_static / Jdxp_dynamic) and origFlag as ordinary statements of the analog block; they are
    # carried over, with the previous-iterate values read from the store vectors the kernel passes in.
    info["has_limit"] = "Jdxp_static" in info["analog_text"]
    info["analog_functions"] = _analog_functions(C, H) if re.search(r"AnalogFunctions::", info["analog_text"]) else ""
    if re.search(r"AnalogFunctions::", info["analog_text"]) and not info["analog_functions"]:
        raise Unsupported("analog functions are called but namespace AnalogFunctions was not found")
    # loads: vector rows and matrix slots (offset form after #else: row and column are spelled out)
    loads = {}
    for key, fname, vec in (("F", "bool Instance::loadDAEFVector()", "daeFVectorPtr"), ("Q", "bool Instance::loadDAEQVector()", "daeQVectorPtr")):
        t = _function_body(C, fname)
        rows = re.findall(r"\(\*extData\.%s\)\[li_(\w+)\]\s*\+=\s*([^;]+);" % vec, t)
        loads[key] = rows
    for key, fname, mat in (("JF", "bool Instance::loadDAEdFdx()", "dFdx"), ("JQ", "bool Instance::loadDAEdQdx()", "dQdx")):
        t = _function_body(C, fname)
        t = t[t.index("#else"):t.index("#endif")]
        loads[key] = re.findall(r"%s\[li_(\w+)\]\[A_(\w+?)_?Offset\]\s*\+=\s*([^;]+);" % mat, t)
    for key, fname, vec, arr in (("FL", "bool Instance::loadDAEFVector()", "dFdxdVp", "Jdxp_static"), ("QL", "bool Instance::loadDAEQVector()", "dQdxdVp", "Jdxp_dynamic")):
        t = _function_body(C, fname)
        loads[key] = re.findall(r"%s\[li_(\w+)\]\s*\+=\s*([^;]+);" % vec, t)
    info["loads"] = loads
    # output variables (operating-point quantities for .PRINT): Instance::updatePrimaryState copies them to the store
    # vector, slots in registerStoreLIDs order
    order, writes = [], []
    if "void Instance::registerStoreLIDs" in C:
        order = re.findall(r"li_store_(\w+)\s*=\s*stoLIDVec\[i\+\+\];", _function_body(C, "void Instance::registerStoreLIDs"))
    if "bool Instance::updatePrimaryState()" in C:
        writes = re.findall(r"stoVec\[li_store_(\w+)\]\s*=\s*([^;]+);", _function_body(C, "bool Instance::updatePrimaryState()"))
    info["store_order"] = order
    info["store_writes"] = [(order.index(n), e.strip()) for n, e in writes if n in order]
    return info


def _unknown_index(info, li_name):
    ids = info["ids"]
    for k in ("admsNodeID_" + li_name, "admsBRA_ID_" + li_name[4:] if li_name.startswith("BRA_") else None):
        if k and k in ids:
            return ids[k]
    raise Unsupported("unknown LID name li_%s" % li_name)


def emit(info, name):
    ids = info["ids"]
    nunk, nprobe = len(info["unknowns"]), len(info["probes"])
    locals_text, analog = info["locals_text"], info["analog_text"]
    nstore = len(info.get("store_order", []))
    for slot, expr in info.get("store_writes", []):
        analog += "\nXBSTORE_%d_ = %s;" % (slot, expr)
    # given("X") / model_.given("X") (DeviceEntity::given: was the parameter set in the netlist?) -> flag fields
    given_fields = []
    def given_sub(m):
        key = ("MG:" if m.group(1) else "IG:") + m.group(2)
        if key not in given_fields:
            given_fields.append(key)
        return "XB_GIVEN_%d_" % given_fields.index(key)
    analog = re.sub(r"(model_\.)?\bgiven\(\"(\w+)\"\)", given_sub, analog)
    # $port_connected: DeviceInstance::portsConnected_[k]
    def port_sub(m):
        key = "IP:" + m.group(1)
        if key not in given_fields:
            given_fields.append(key)
        return "XB_GIVEN_%d_" % given_fields.index(key)
    analog = re.sub(r"\bportsConnected_\[(\w+)\]", port_sub, analog)
    local_names = set(re.findall(r"\b(?:double|int|bool)\s+(\w+)\s*=", locals_text + analog))
    # ---- record fields: model members and instance members that the analog block reads ----
    fields = []
    used_model = []
    for m in re.finditer(r"model_\.(\w+)", analog):
        if m.group(1) not in used_model:
            used_model.append(m.group(1))
    model_names = [n for n, _ in info["model_members"]]
    for n in used_model:
        if n not in model_names:
            raise Unsupported("model_.%s is not a declared model variable" % n)
        fields.append("M:" + n)
    inst_used = []
    for n, _ in info["inst_members"]:
        if n in local_names:
            continue
        if re.search(r"(?<![\w.])%s\b(?!\s*\()" % re.escape(n), analog):
            inst_used.append(n)
            fields.append("I:" + n)
    fields += given_fields
    fidx = {f: k for k, f in enumerate(fields)}

    # ---- rewrite the analog block ----
    t = analog
    t = re.sub(r"/\*.*?\*/", "", t, flags=re.S)
    # probe extraction: (*solVectorPtr)[li_x] -> V[k]
    def sol(m):
        return "V[%d]" % _unknown_index(info, m.group(1))
    t = re.sub(r"\(\*solVectorPtr\)\[li_(\w+)\]", sol, t)
    has_limit = info.get("has_limit", False)
    if has_limit:
        # previous-iterate values of the limited probes: store slots in registerStoreLIDs order
        order = info.get("store_order", [])
        def sto(m):
            return "%s[%d]" % ("xbcs_" if m.group(1) == "curr" else "xbns_", order.index(m.group(2)))
        t = re.sub(r"\(\(\*extData\.(curr|next)StoVectorPtr\)\)\[li_store_(\w+)\]", sto, t)
        t = re.sub(r"\bdevSupport\.(\w+)\(", r"adms_\1(", t)
        t = re.sub(r"getSolverState\(\)\.newtonIter\b", "S.newtonIter", t)
        t = re.sub(r"getSolverState\(\)\.initJctFlag_?\b", "(S.initJctFlag != 0)", t)
        t = re.sub(r"getSolverState\(\)\.locaEnabledFlag\b", "(S.locaEnabledFlag != 0)", t)
        t = re.sub(r"getSolverState\(\)\.inputOPFlag\b", "false", t)
        t = t.replace("getDeviceOptions().voltageLimiterFlag", "(S.voltageLimiterFlag != 0)")
    else:
        t = re.sub(r"^\s*d_probeVars\[\w+\]\[\w+\]\s*=\s*1\.0;\s*$", "", t, flags=re.M)      # the independent variables' own seeds
    if not has_limit:
      t = re.sub(r"d_probeVars\[(\w+)\]\[(\w+)\]", lambda m: "1.0" if m.group(1) == m.group(2) else "0.0", t)
    # diagnostics inside the analog block ($strobe / $warning / $error of the Verilog-A source) have no place in a kernel
    t = re.sub(r"\b(?:UserWarning0?|UserError0?|UserFatal0?|UserInfo0?|Report::\w+)\s*\([^;]*;", ";", t)
    mtype = dict(info["model_members"])
    itype = dict(info["inst_members"])
    # members that the analog block ASSIGNS (variables of global_instance / global_model scope, operating-point outputs):
    # they become locals initialised from the record
    assigned = []
    for n in used_model:
        if re.search(r"model_\.%s\)?\s*(?:[-+*/]?=)(?!=)" % re.escape(n), t):
            assigned.append(("M:" + n, "m_%s_" % n, mtype[n]))
            t = re.sub(r"\(?model_\.%s\b\)?" % re.escape(n), lambda m, n=n: "m_%s_" % n if not (m.group(0).startswith("(") ^ m.group(0).endswith(")")) else m.group(0).replace("model_.%s" % n, "m_%s_" % n), t)
    for n in list(inst_used):
        if re.search(r"(?<![\w.])%s\)?\s*(?:[-+*/]?=)(?!=)" % re.escape(n), t):
            assigned.append(("I:" + n, "i_%s_" % n, itype[n]))
            t = re.sub(r"(?<![\w.])%s\b(?!\s*\()" % re.escape(n), "i_%s_" % n, t)
    def field(key, ctype):          # integer / boolean members keep their C type (conditions, integer arithmetic)
        return "xbrec_.f[%d]" % fidx[key] if ctype == "double" else "%s(to_double(xbrec_.f[%d]))" % (ctype, fidx[key])
    t = re.sub(r"model_\.(\w+)", lambda m: field("M:" + m.group(1), mtype[m.group(1)]), t)
    for n in inst_used:
        t = re.sub(r"(?<![\w.])%s\b(?!\s*\()" % re.escape(n), field("I:" + n, itype[n]), t)
    t = t.replace("getDeviceOptions().gmin", "S.gmin")
    t = t.replace("getSolverState().noiseFlag", "false")
    t = re.sub(r"getSolverState\(\)\.(\w+)", lambda m: {"dcopFlag": "(S.dcopFlag != 0)", "tranopFlag": "(S.tranopFlag != 0)"}.get(m.group(1), "false"), t)
    t = re.sub(r"static_cast<double>\(([^()]*)\)", r"real(\1)", t)
    t = t.replace("std::max", "adms_max").replace("std::min", "adms_min")
    t = re.sub(r"\bstd::(exp|log|sqrt|pow|fabs|tanh|sinh|cosh|atan|sin|cos|tan|log10|abs)\b", r"\1", t)
    t = re.sub(r"\bpow\(", "rpow(", t)
    t = re.sub(r"\bdouble\b", "real", t)
    t = re.sub(r"XBSTORE_(\d+)_", lambda m: "o.store[%s]" % m.group(1), t)
    t = re.sub(r"XB_GIVEN_(\d+)_", lambda m: "(xbrec_.f[%d] != 0.0)" % fidx[given_fields[int(m.group(1))]], t)
    leftovers = re.findall(r"\b(?:std::\w+|Xyce::\w+|UserError|Report::\w+|extData\.\w+|getName\(\))", t)
    if leftovers:
        raise Unsupported("untranslated constructs in the analog block: %s" % sorted(set(leftovers))[:6])
    lt = re.sub(r"\bdouble\b", "real", locals_text)
    for key, local, ctype in assigned:
        lt += "\n  %s %s = %s;" % ("real" if ctype == "double" else ctype, local,
                                   "xbrec_.f[%d]" % fidx[key] if ctype == "double" else "%s(to_double(xbrec_.f[%d]))" % (ctype, fidx[key]))

    # ---- loads ----
    slot_of = {rc: k for k, rc in enumerate(info["stamp"])}
    lines = []
    for key, arr in (("F", "staticContributions"), ("Q", "dynamicContributions")):
        for li, expr in info["loads"][key]:
            lines.append("  o.%s[%d] += %s;" % (key, _unknown_index(info, li), expr.strip()))
    lim_lines = []
    for key in ("FL", "QL"):
        for li, expr in info["loads"].get(key, []):
            lim_lines.append("    o.%s[%d] += %s;" % (key, _unknown_index(info, li), expr.strip()))
    for key in ("JF", "JQ"):
        for li, ptr, expr in info["loads"][key]:
            row = _unknown_index(info, li)
            assert ptr.startswith(li + "_Equ_"), (li, ptr)
            col_name = ptr[len(li) + 5:]
            col_name = re.sub(r"_(Node|Var)_?$", "", col_name)
            col = _unknown_index(info, col_name)
            if (row, col) not in slot_of:
                raise Unsupported("load onto (%d, %d), which is not in the Jacobian stamp" % (row, col))
            lines.append("  o.%s[%d] += %s;" % (key, slot_of[(row, col)], expr.strip()))

    enum = ",\n  ".join("%s = %d" % (k, v) for k, v in sorted(ids.items(), key=lambda kv: (kv[0].split("_")[0], kv[1])))
    rows = ", ".join(str(r) for r, _ in info["stamp"])
    cols = ", ".join(str(c) for _, c in info["stamp"])
    out = []
    out.append("// GENERATED by xyce_b200/adms/translate.py from the admsXml output %s -- do not edit, do not commit.\n" % info["namespace"])
    out.append("#pragma once\n#include \"../xb_common.h\"\n")
    out.append("namespace xb {\nnamespace adms {\nnamespace gen_%s {\n" % name)
    out.append("enum {\n  %s\n};\n" % enum)
    out.append("constexpr int kNodes = %d, kExt = %d, kSlots = %d, kProbes = %d, kNumFields = %d, kNumStore = %d;\n" % (nunk, info["n_ext"], len(info["stamp"]), nprobe, len(fields), nstore))
    out.append("#define XB_ADMS_GEN_%s_FIELDS \"%s\"\n" % (name, " ".join(fields)))
    out.append("struct Rec { real f[kNumFields > 0 ? kNumFields : 1]; };\n")
    out.append("struct Out { real F[kNodes], Q[kNodes], FL[kNodes], QL[kNodes], JF[kSlots], JQ[kSlots], store[kNumStore > 0 ? kNumStore : 1]; int origFlag; };\n")
    out.append("XB_HD real adms_vt(real T) { return kKoverQ * T; }\n")
    out.append("XB_HD real adms_max(real a, real b) { return a < b ? b : a; }\nXB_HD real adms_min(real a, real b) { return b < a ? b : a; }\n")
    out.append("// the templates' limited exponential (N_DEV_ADMS*.h: exp below 80, its tangent above)\n"
               "XB_HD real limexp(real x) { return (x < 80.0) ? exp(x) : exp(real(80.0)) * (x - 79.0); }\n")
    out.append("// SPICE3 junction limiters as the templates call them (Core/N_DEV_DeviceSupport.C:161-300)\n"
               "XB_HD real adms_pnjlim(real vnew, real vold, real vt, real vcrit, int *icheck) { int ic = 0; const real v = pnjlim(vnew, vold, vt, vcrit, ic); *icheck = ic; return v; }\n"
               "XB_HD real adms_pnjlim_new(real vnew, real vold, real vt, real vcrit, int *icheck) {\n"
               "  if ((vnew > vcrit) && (fabs(vnew - vold) > (vt + vt))) {\n"
               "    if (vold > 0) { const real arg = (vnew - vold) / vt; vnew = (arg > 0) ? vold + vt * (2 + log(arg - 2)) : vold - vt * (2 + log(2 - arg)); }\n"
               "    else vnew = vt * log(vnew / vt);\n"
               "    *icheck = 1;\n"
               "  } else if (vnew < 0) {\n"
               "    const real arg = (vold > 0) ? -vold - 1 : 2 * vold - 1;\n"
               "    if (vnew < arg) { vnew = arg; *icheck = 1; } else *icheck = 0;\n"
               "  } else *icheck = 0;\n"
               "  return vnew;\n}\n"
               "XB_HD real adms_fetlim(real vnew, real vold, real vto) { return fetlim(vnew, vold, vto); }\n"
               "XB_HD real adms_limvds(real vnew, real vold) { return limvds(vnew, vold); }\n")
    if info.get("analog_functions"):
        out.append(info["analog_functions"])
    out.append("// RecT: anything with f[k] -> field k (Rec on the host; on the device a view that loads a field where it is used,\n"
               "// so that a 76-field record does not sit in registers for the whole evaluation)\n")
    out.append("template <class RecT>\nXB_HD void evaluate(const SolverFlags &S, const RecT &xbrec_, const real *V, Out &o, const real *xbcs_ = nullptr, const real *xbns_ = nullptr) {\n")
    out.append("  real probeVars[kProbes];\n  real staticContributions[kNodes], dynamicContributions[kNodes];\n")
    out.append("  real d_staticContributions[kNodes][kProbes], d_dynamicContributions[kNodes][kProbes];\n")
    out.append("  real noiseContribsPower[16], noiseContribsExponent[16];\n  (void)noiseContribsPower; (void)noiseContribsExponent; (void)S; (void)xbcs_; (void)xbns_;\n")
    out.append("  bool origFlag = true; (void)origFlag;\n")
    if info.get("has_limit"):
        out.append("  real d_probeVars[kProbes][kProbes], probeDiffs[kProbes], Jdxp_static[kNodes], Jdxp_dynamic[kNodes];\n"
                   "#pragma unroll\n  for (int i = 0; i < kProbes; ++i) { probeDiffs[i] = 0.0;\n#pragma unroll\n    for (int j = 0; j < kProbes; ++j) d_probeVars[i][j] = 0.0; }\n"
                   "#pragma unroll\n  for (int i = 0; i < kNodes; ++i) { Jdxp_static[i] = 0.0; Jdxp_dynamic[i] = 0.0; }\n")
    out.append("#pragma unroll\n  for (int i = 0; i < kNodes; ++i) {\n    staticContributions[i] = 0.0; dynamicContributions[i] = 0.0;\n"
               "    o.F[i] = 0.0; o.Q[i] = 0.0; o.FL[i] = 0.0; o.QL[i] = 0.0;\n#pragma unroll\n"
               "    for (int j = 0; j < kProbes; ++j) { d_staticContributions[i][j] = 0.0; d_dynamicContributions[i][j] = 0.0; }\n  }\n")
    out.append("#pragma unroll\n  for (int s = 0; s < kSlots; ++s) { o.JF[s] = 0.0; o.JQ[s] = 0.0; }\n")
    out.append("#pragma unroll\n  for (int s = 0; s < (kNumStore > 0 ? kNumStore : 1); ++s) o.store[s] = 0.0;\n")
    out.append("  " + lt.strip() + "\n")
    out.append(t)
    out.append("\n  // ---- loads (loadDAEFVector / loadDAEQVector / loadDAEdFdx / loadDAEdQdx) ----\n")
    out.append("\n".join(lines) + "\n")
    if lim_lines:
        out.append("  if ((S.voltageLimiterFlag != 0) && !origFlag) {      // loadDAEFVector / loadDAEQVector: dFdxdVp, dQdxdVp\n" + "\n".join(lim_lines) + "\n  }\n")
    out.append("  o.origFlag = origFlag ? 1 : 0;\n}\n")
    out.append("static const int kSlotRow[kSlots] = {%s};\nstatic const int kSlotCol[kSlots] = {%s};\n" % (rows, cols))
    out.append("// what the generic kernel (simple_kernels.cu: adms_gen_kernel<Traits>) and the registry need\n")
    out.append("struct Traits {\n  typedef gen_%s::Rec Rec;\n  typedef gen_%s::Out Out;\n" % (name, name))
    out.append("  static constexpr int kNodes = gen_%s::kNodes, kExt = gen_%s::kExt, kSlots = gen_%s::kSlots, kNumFields = gen_%s::kNumFields, kNumStore = gen_%s::kNumStore;\n" % (name, name, name, name, name))
    out.append("  template <class RecT> static XB_HD void eval(const SolverFlags &S, const RecT &R, const real *V, Out &o, const real *cs = nullptr, const real *ns = nullptr) { evaluate(S, R, V, o, cs, ns); }\n")
    out.append("  static constexpr bool kHasLimit = %s;\n" % ("true" if info.get("has_limit") else "false"))
    out.append("  static const char *name() { return \"%s\"; }\n  static const char *fields() { return XB_ADMS_GEN_%s_FIELDS; }\n" % (name, name))
    out.append("  static const int *slot_row() { return kSlotRow; }\n  static const int *slot_col() { return kSlotCol; }\n};\n")
    out.append("}  // namespace gen_%s\n}  // namespace adms\n}  // namespace xb\n" % name)
    return "".join(out), fields


def emit_fill(info, name, fields):
    """Reference-side glue: copy the members the evaluator reads out of the admsXml-generated Instance / Model objects
    into the flat record, and the unknowns' LIDs in evaluator order (what a GpuMaster adaptor for this model does)."""
    out = ["// GENERATED by xyce_b200/adms/translate.py -- reference-side record filler for %s\n#pragma once\n" % info["namespace"]]
    out.append("template <class Instance>\ninline int adms_fill_%s(const Instance &in, double *rec, int *lids) {\n  int k = 0;\n" % name)
    for f in fields:
        kind, n = f.split(":")
        if kind == "IP":
            out.append("  rec[k++] = in.portsConnected_[%s%s] ? 1.0 : 0.0;\n" % ("" if n.isdigit() else "Instance::", n))
        elif kind in ("MG", "IG"):
            out.append("  rec[k++] = %sgiven(\"%s\") ? 1.0 : 0.0;\n" % ("in.model_." if kind == "MG" else "in.", n))
        else:
            out.append("  rec[k++] = (double)%s%s;\n" % ("in.model_." if kind == "M" else "in.", n))
    out.append("  int j = 0;\n")
    for u in info["unknowns"]:
        li = "li_" + (u[len("admsNodeID_"):] if u.startswith("admsNodeID_") else "BRA_" + u[len("admsBRA_ID_"):])
        out.append("  lids[j++] = in.%s;\n" % li)
    out.append("  return k;\n}\n")
    out.append("static const int adms_nlids_%s = %d;      // unknowns of the evaluator (collapsed nodes alias their targets, -1 = ground)\n" % (name, len(info["unknowns"])))
    return "".join(out)


def translate(cfile, hfile, out_path, name=None):
    info = parse(cfile, hfile)
    name = name or re.sub(r"^N_DEV_ADMS", "", os.path.splitext(os.path.basename(cfile))[0])
    text, fields = emit(info, name)
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    open(out_path, "w").write(text)
    open(os.path.splitext(out_path)[0] + "_fill.h", "w").write(emit_fill(info, name, fields))
    return dict(name=name, namespace=info["namespace"], fields=fields, nodes=len(info["unknowns"]), n_ext=info["n_ext"],
                slots=len(info["stamp"]), unknown_names=info["unknowns"], stamp=info["stamp"], nstore=len(info.get("store_order", [])))


def translate_all(adms_dir, out_dir, names):
    """Translate the listed models of one directory of admsXml output and write the registry the library build and the
    oracle build include (registry.h: evaluators + X-macro list; oracle_registry.h: reference headers + fillers)."""
    done = []
    for n in names:
        c, h = os.path.join(adms_dir, "N_DEV_ADMS%s.C" % n), os.path.join(adms_dir, "N_DEV_ADMS%s.h" % n)
        r = translate(c, h, os.path.join(out_dir, "adms_%s.h" % n), n)
        done.append(r)
    # registry.h: every evaluator (host mirror); registry_info.h: the tables only (library dispatch, compiles in no
    # time); kernel_<model>.cu: one translation unit per model so that the library build compiles them in parallel
    with open(os.path.join(out_dir, "registry.h"), "w") as f:
        f.write("// GENERATED by xyce_b200/adms/translate.py\n#pragma once\n")
        for r in done:
            f.write("#include \"adms_%s.h\"\n" % r["name"])
        f.write("#define XB_ADMS_GEN_COUNT %d\n#define XB_ADMS_GEN_LIST(X) %s\n" % (len(done), " ".join("X(%d, %s)" % (i, r["name"]) for i, r in enumerate(done))))
    with open(os.path.join(out_dir, "registry_info.h"), "w") as f:
        f.write("// GENERATED by xyce_b200/adms/translate.py -- tables of the translated models, no evaluator code\n#pragma once\n")
        for r in done:
            f.write("static const int kAdmsRow_%s[] = {%s};\nstatic const int kAdmsCol_%s[] = {%s};\n"
                    % (r["name"], ", ".join(str(a) for a, _ in r["stamp"]), r["name"], ", ".join(str(b) for _, b in r["stamp"])))
            f.write("static const char kAdmsFields_%s[] = \"%s\";\n" % (r["name"], " ".join(r["fields"])))
        f.write("#define XB_ADMS_GEN_COUNT %d\n// X(index, name, unknowns, external nodes, stamp entries, record fields, store slots)\n#define XB_ADMS_GEN_LIST(X) %s\n"
                % (len(done), " ".join("X(%d, %s, %d, %d, %d, %d, %d)" % (i, r["name"], r["nodes"], r["n_ext"], r["slots"], len(r["fields"]), r["nstore"]) for i, r in enumerate(done))))
    for r in done:
        with open(os.path.join(out_dir, "kernel_%s.cu" % r["name"]), "w") as f:
            f.write("// GENERATED by xyce_b200/adms/translate.py -- the generic kernel instantiated for one translated model\n"
                    "#include \"../adms_gen_kernel.cuh\"\n#include \"adms_%s.h\"\n"
                    "namespace xb {\nnamespace simple {\n"
                    "void launch_adms_gen_%s(const GroupDev &g, const b4::LoadArgs &a, cudaStream_t s) { launch_adms_gen<adms::gen_%s::Traits>(g, a, s); }\n"
                    "}  // namespace simple\n}  // namespace xb\n" % (r["name"], r["name"], r["name"]))
    with open(os.path.join(out_dir, "oracle_registry.h"), "w") as f:
        f.write("// GENERATED by xyce_b200/adms/translate.py -- oracle (test infrastructure) side\n#pragma once\n")
        for r in done:
            f.write("#include <N_DEV_ADMS%s.h>\n#include \"adms_%s_fill.h\"\n" % (r["name"], r["name"]))
        f.write("#define XB_ADMS_ORACLE_LIST(X) %s\n" % " ".join("X(%s, %s)" % (r["name"], r["namespace"]) for r in done))
    return done


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[1] == "--all":
        for r in translate_all(sys.argv[2], sys.argv[3], sys.argv[4:]):
            print("%s: %d unknowns (%d external), %d stamp entries, %d record fields" % (r["name"], r["nodes"], r["n_ext"], r["slots"], len(r["fields"])))
        sys.exit(0)
    if len(sys.argv) < 4:
        sys.exit(__doc__)
    r = translate(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    print("%s: %d unknowns (%d external), %d stamp entries, %d record fields" % (r["name"], r["nodes"], r["n_ext"], r["slots"], len(r["fields"])))
