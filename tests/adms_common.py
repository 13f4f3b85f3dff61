"""Circuits of isolated instances of the models that went through the ADMS translator (xyce_b200/adms/translate.py),
built on the reference's own admsXml-generated classes (oracle/_ref)."""
import numpy as np

# model cards: parameter names as in the generated loadModelParameters (upper-cased by the harness like the parser does)
ADMS_CARDS = {
    "mvs_2_0_0_etsoi": {
        "nmos": ("NMOS", dict(TYPE=1), {}),
        "pmos": ("PMOS", dict(TYPE=-1, W=2e-6, LGDR=60e-9, RS0=200e-6, N0=1.5, DELTA=0.15, ND=0.05, MU_EFF=0.6, KSEE=0.15), {}),
        "short": ("NMOS", dict(TYPE=1, LGDR=32e-9, DLG=6e-9, BETA=1.8, THETA=2.2, TJUN=350.0, NU=0.6, ENERGY_DIFF_VOLT=0.1), {}),
    },
    "mvs_2_0_0_hemt": {
        "nmos": ("NMOS", dict(TYPE=1), {}),
        "wide": ("NMOS", dict(TYPE=1, W=5e-6, LGDR=120e-9, RC0=200e-6, N0=1.4, DELTA=0.1, ND=0.02, TJUN=330.0), {}),
    },
    "ekv_va": {
        "nmos": ("NMOS", dict(TYPE=1), dict(L=1e-6, W=10e-6)),
        "pmos": ("PMOS", dict(TYPE=-1, VTO=0.55, KP=35e-6, GAMMA=0.6, PHI=0.8, THETA=0.03, LAMBDA=0.6, UCRIT=3e6), dict(L=0.5e-6, W=20e-6)),
        "short_hot": ("NMOS", dict(TYPE=1, VTO=0.45, GAMMA=0.65, PHI=0.85, KP=180e-6, E0=9e7, UCRIT=4e6, LAMBDA=0.3, WETA=0.1,
                                   LETA=0.3, IBA=2e8, IBB=2e8, IBN=0.6, TCV=1e-3, BEX=-1.4, UCEX=1.6, TNOM=25.0, TRISE=40.0),
                      dict(L=0.25e-6, W=5e-6, AS=5e-12, AD=5e-12, PS=12e-6, PD=12e-6)),
    },
    # JUNCAP 200.3 junction diode (2 terminals, 129 record fields)
    "JUNCAP200": {
        "default": ("D", {}, {}),
        "sized": ("D", dict(CJORBOT=1.5e-3, CJORSTI=2e-9, IDSATRBOT=5e-12, CSRHBOT=2e2, CTATBOT=1e2, VBIRBOT=0.9), dict(AB=2e-12, LS=3e-6, LG=1e-6)),
    },
    # HICUM level 0 and level 2 bipolar transistors: default card = every internal node collapsed, "res" = series
    # resistances present (internal collector / base / emitter nodes are real unknowns)
    "hic0_full": {
        "default": ("NPN", {}, {}),
        "res": ("NPN", dict(RCX=12.0, RBX=8.0, RBI0=25.0, RE=1.5, IS=2e-17, CJE0=8e-15, CJCI0=3e-15, T0=3e-12), {}),
    },
    "hicumL2va": {
        "default": ("NPN", {}, {}),
        "res": ("NPN", dict(RCX=12.0, RBX=8.0, RBI0=25.0, RE=1.5, C10=3e-30, QP0=4e-14, CJEI0=8e-15, CJCI0=3e-15, T0=3e-12), {}),
    },
    # PSP 103 MOSFET (13 unknowns before collapsing, 636 record fields)
    "PSP103VA": {
        "nmos": ("NMOS", dict(TYPE=1), dict(L=1e-7, W=1e-6)),
        "pmos_rg": ("PMOS", dict(TYPE=-1, RGO=30.0, RBULKO=50.0, SWJUNCAP=3), dict(L=2e-7, W=2e-6)),
    },
    # BSIM6 bulk MOSFET (analog functions, given() tests, node collapsing incl. the thermal node to ground)
    "bsim6": {
        "nmos": ("NMOS", dict(TYPE=1), dict(L=1e-7, W=1e-6)),
        "pmos_rg": ("PMOS", dict(TYPE=-1, RGATEMOD=1, RBODYMOD=1, RDSMOD=1), dict(L=2e-7, W=2e-6, NF=2)),
    },
    # BSIM-CMG 110 multi-gate (FinFET) model: 1 379 record fields
    "bsimcmg_110": {
        "nfin": ("NMOS", dict(DEVTYPE=1), dict(L=3e-8, NFIN=4)),
        "pfin_rg": ("PMOS", dict(DEVTYPE=0, RGATEMOD=1, RDSMOD=1), dict(L=5e-8, NFIN=2)),
    },
    # CMC diode (internal nodes for the series resistance and the depletion / charge states)
    "DIODE_CMC": {
        "default": ("D", {}, {}),
        "rs": ("D", dict(RS=5.0, CJORBOT=1.5e-3, IDSATRBOT=5e-12), dict(AB=2e-12, LS=3e-6)),
    },
    # models with $limit (limited junction voltages, dFdxdVp / dQdxdVp correction terms, origFlag): VBIC 1.3 and Mextram 504
    "vbic13": {
        "npn": ("NPN", {}, {}),
        "pnp_res": ("PNP", dict(TYPE=1, RCX=15.0, RCI=40.0, RBX=10.0, RBI=30.0, RE=2.0, IS=3e-17, CJE=8e-15, CJC=3e-15, TF=5e-12), {}),
    },
    "bjt504va": {
        "npn": ("NPN", {}, {}),
        "res": ("NPN", dict(RCC=20.0, RCV=120.0, RBC=15.0, RBV=80.0, RE=3.0, IS=3e-17, CJE=8e-15, CJC=5e-15, TAUE=3e-12), {}),
    },
    # the variants with a substrate terminal / a thermal node (self heating)
    "vbic13_4t": {
        "npn": ("NPN", {}, {}),
        "pnp_res": ("PNP", dict(TYPE=1, RCX=15.0, RCI=40.0, RBX=10.0, RBI=30.0, RE=2.0, RS=5.0, IS=3e-17, CJE=8e-15, CJC=3e-15, TF=5e-12), {}),
    },
    "bjt504tva": {
        "npn": ("NPN", {}, {}),
        "res": ("NPN", dict(RCC=20.0, RCV=120.0, RBC=15.0, RBV=80.0, RE=3.0, IS=3e-17, CJE=8e-15, CJC=5e-15, TAUE=3e-12, RTH=300.0, CTH=3e-9), {}),
    },
}
# bias windows (uniform node voltages) that keep every model inside its working range
BIAS = {"mvs_2_0_0_etsoi": (-0.6, 1.0), "mvs_2_0_0_hemt": (-0.6, 1.0), "ekv_va": (-1.2, 1.8), "JUNCAP200": (-0.8, 0.6),
        "hic0_full": (0.0, 0.7), "hicumL2va": (0.0, 0.7), "PSP103VA": (0.0, 0.6), "bsim6": (0.0, 0.6), "bsimcmg_110": (0.0, 0.6),
        "DIODE_CMC": (-0.5, 0.5), "vbic13": (0.0, 1.5), "bjt504va": (0.0, 1.5), "vbic13_4t": (0.0, 1.5), "bjt504tva": (0.0, 1.5)}
LIMITED = ("vbic13", "bjt504va", "vbic13_4t", "bjt504tva")


# unknowns that need their own window: V(sf) of the HEMT variant (the Fermi-Dirac fit of the model takes a fractional
# power of a polynomial in V(sf)/phit that is positive only for V(sf) < 0 -- the reference object returns NaN beyond)
OVERRIDE = {"mvs_2_0_0_hemt": [(5, -0.6, -0.02)]}


def bias_vector(model, n, lids_list, rng):
    x = rng.uniform(*BIAS[model], n)
    for pos, lo, hi in OVERRIDE.get(model, []):
        for l in lids_list:
            if l[pos] >= 0:
                x[l[pos]] = rng.uniform(lo, hi)
    return x


def adms_circuit(ref_cls, model, card, n_ext, n_dev=6, seed=0):
    nt = max(4, n_ext)
    c = ref_cls(nt * n_dev)
    mtype, mp, ip = ADMS_CARDS[model][card]
    c.add_dev_model("adms:" + model, "amod", mtype, 1, mp)
    for i in range(n_dev):
        c.add_dev_instance("adms:" + model, "M:%d" % i, "amod", [nt * i + k for k in range(n_ext)], ip)
    c.finalize()
    return c


def outvars_close(got, want, tol, nstore=None, illcond_share=0.0, illcond_tol=0.0):
    """output variables in the store vector: heterogeneous quantities (resistances, capacitances, transit frequencies,
    1 / 0 = inf for a collapsed resistance): non-finite entries must coincide, finite ones agree entry by entry -- with
    nstore given, relative to max(|entry|, 1e-3 * largest entry of the same variable over all instances), which is the
    cancellation floor for variables that are differences.
    illcond_share / illcond_tol (GPU only): a few output variables are ill-conditioned functions of the operating point
    (PSP 103's slots 68-70: even the strict GPU build, <= 1 ulp per operation from the host, moves them by 1.7e-8, the
    fast build by 3e-5, while every other variable sits at 1e-12; the host build of the same statements reproduces all of
    them to the last bits): that share of the variables may deviate up to illcond_tol."""
    return outvars_mismatch(got, want, tol, nstore, illcond_share, illcond_tol) is None


def outvars_mismatch(got, want, tol, nstore=None, illcond_share=0.0, illcond_tol=0.0):
    """None when the output variables agree (see outvars_close), else a description of the first disagreement"""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    fin = np.isfinite(want)
    if not np.array_equal(fin, np.isfinite(got)):
        i = int(np.argmax(fin != np.isfinite(got)))
        return "finite / non-finite pattern differs at %d: got %r want %r" % (i, got[i], want[i])
    # non-finite entries (1 / 0 of a collapsed resistance ...) must coincide as such; inf vs NaN and the sign of an
    # infinity are not compared: the fast division returns NaN for a zero divisor (xb_fastmath.h) where IEEE gives +-inf
    floor = np.full(want.shape, 1e-30)
    if nstore and len(want) % nstore == 0:
        w2 = np.where(fin, np.abs(want), 0.0).reshape(-1, nstore)
        floor = np.maximum(floor, np.broadcast_to(1e-3 * np.max(w2, axis=0, keepdims=True), w2.shape).reshape(-1))
    err = np.where(fin, np.abs(np.where(fin, got, 0.0) - np.where(fin, want, 0.0)) / np.maximum(np.abs(np.where(fin, want, 1.0)), floor), 0.0)
    i = int(np.argmax(err))
    worst = "worst entry %d (variable %s): got %.17g want %.17g err %.3e" % (i, i % nstore if nstore else "?", got[i], want[i], err[i])
    if nstore and len(want) % nstore == 0 and illcond_share > 0.0:
        per_var = err.reshape(-1, nstore).max(axis=0)
        bad = per_var > tol
        ok = np.sum(bad) <= max(1, int(illcond_share * nstore)) and np.all(per_var <= illcond_tol)
        return None if ok else "%d of %d variables beyond %.0e; %s" % (int(np.sum(bad)), nstore, tol, worst)
    return None if np.all(err <= tol) else worst
