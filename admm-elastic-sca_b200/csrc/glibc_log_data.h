// glibc_log_data.h -- constants of glibc 2.39 log() (sysdeps/ieee754/dbl-64/e_log_data.c, from ARM optimized-routines,
// the table __log_data: ln2hi, ln2lo, poly A[5], poly1 B[11], tab[128]{invc,logc}), read from this image's libm.so.6
// by tools/extract_glibc_log.py.  Needed because the reference prox calls log() (TetForce.cpp:220,241) inside a
// truncated, branchy optimiser: a 1-ulp difference in log flips line-search branches, so the device must return the
// same bits as the host libm the reference links against.  Values are IEEE-754 bit patterns.
#pragma once
namespace admmb {
static const unsigned long long GLIBC_LOG_DATA[274] = {
#include "glibc_log_data.inc"
};
// __exp_data of the same libm (tools/extract_glibc_exp.py): invln2N, shift, negln2hiN, negln2loN, C2..C5, tab[2*128]
static const unsigned long long GLIBC_EXP_DATA[264] = {
#include "glibc_exp_data.inc"
};
} // namespace admmb
