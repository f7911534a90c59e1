/*
 * orc_fp.h -- ORACLE (test infrastructure, never shipped): the CANONICAL fp32
 * evaluation order of the GLSL restatement.
 *
 * GLSL 3.30 leaves the contraction of a*b+c implementation-defined (no
 * `precise` qualifier in the reference's shaders), so a restatement has to
 * pick one.  The canonical order is the one a GPU shader compiler emits and
 * the one the CUDA kernels use, so that kernel and oracle agree bit for bit:
 *
 *   R1  dot(a,b)      = fma(a[n-1],b[n-1], ... fma(a[1],b[1], a[0]*b[0]))
 *   R2  M * v (row r) = the same left-to-right fma chain over the columns
 *   R3  a*b + c       = fma(a,b,c)      a - b*c = fma(-b,c,a)
 *   R4  sums of dots (mdot), differences, products: single IEEE operations,
 *       left to right as the shader spells them
 *   R5  sqrt and / are the correctly rounded IEEE operations
 *
 * Building with -DORC_STRICT gives the other admissible reading (no
 * contraction anywhere: every a*b+c is two roundings).  liborc_strict.so is
 * built from the same sources and is used only to MEASURE how far apart the
 * two readings are -- that distance is the stated float tolerance.
 *
 * Integer-deciding maths (cnoise, noise layer selection) never uses these
 * helpers: it is plain non-contracted fp32 like the reference's C++ build.
 */
#ifndef ORC_FP_H
#define ORC_FP_H
#include <math.h>

#ifdef ORC_STRICT
static inline float orc_fma(float a, float b, float c) { return a * b + c; } /* -ffp-contract=off */
#else
static inline float orc_fma(float a, float b, float c) { return fmaf(a, b, c); }
#endif

static inline float orc_dot2(const float *a, const float *b)
{
    return orc_fma(a[1], b[1], a[0] * b[0]);
}
static inline float orc_dot3(const float *a, const float *b)
{
    return orc_fma(a[2], b[2], orc_fma(a[1], b[1], a[0] * b[0]));
}
static inline float orc_dot4(const float *a, const float *b)
{
    return orc_fma(a[3], b[3], orc_fma(a[2], b[2], orc_fma(a[1], b[1], a[0] * b[0])));
}
#endif
