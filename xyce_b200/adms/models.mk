# admsXml-generated models of the reference tree that go through the ADMS translator at build time (library:
# xyce_b200/csrc/Makefile; oracle: oracle/Makefile compiles the same reference classes for the parity tests).
# Translatable today (xyce_b200/adms/translate.py): everything without $limit and without analog functions defined in the
# model file -- also PSP102VA, PSP103TVA, l_utsoi (left out here only to bound the build time).
ADMS_MODELS ?= mvs_2_0_0_etsoi mvs_2_0_0_hemt ekv_va JUNCAP200 hic0_full hicumL2va PSP103VA
