# admsXml-generated models of the reference tree that go through the ADMS translator at build time (library:
# xyce_b200/csrc/Makefile; oracle: oracle/Makefile compiles the same reference classes for the parity tests).
# Translatable today (xyce_b200/adms/translate.py): 23 of the reference's 24 generated models -- all but HBT_X, whose
# analog block reads the solver's initial-condition flag vector.  PSP102VA, PSP103TVA, l_utsoi, bsimcmg, bsimcmg_108,
# bsimsoi, bsimsoi450, bsimsoi461, mvsg_cmc, vbic13_4t, bjt504tva translate too (host build checked) and are left out here
# to bound the build time (nvcc needs more than 45 minutes for the 18-unknown mvsg_cmc evaluator).
ADMS_MODELS ?= mvs_2_0_0_etsoi mvs_2_0_0_hemt ekv_va JUNCAP200 hic0_full hicumL2va PSP103VA bsim6 bsimcmg_110 DIODE_CMC vbic13 bjt504va vbic13_4t bjt504tva
