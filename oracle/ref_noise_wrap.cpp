/* extern "C" face of the reference's own noise.cpp (compiled unchanged into
 * oracle/_ref/libref_noise.so by oracle/Makefile).  Test infrastructure: used
 * only to pin oracle/orc_noise.c bit-exactly. */
#include "proland/math/noise.h"
extern "C" {
long ref_lrandom(long *seed) { return proland::lrandom(seed); }
float ref_frandom(long *seed) { return proland::frandom(seed); }
float ref_cnoise2(float x, float y) { return proland::cnoise(x, y); }
}
