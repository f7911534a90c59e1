/*
 * pl_f2.cuh -- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2, new on sm_100):
 * one instruction issues the same IEEE operation on two independent values.
 * Every hot loop works on PAIRS of texels and spells its maths with these
 * helpers; each half rounds exactly as the scalar operation would, so results
 * stay bit-identical to the oracle.  A packed instruction halves the issue
 * slots, not the register-operand traffic: with three register-pair operands
 * it issues every 3 cycles, with an immediate every 2 (DESIGN.md 3.5,
 * tools/microbench/rf_operands.cu).
 *
 * ptxas caveat (12.9, checked in SASS): a FMUL2 feeding a FADD2 is contracted
 * into FFMA2 even under --fmad=false and despite the .rn modifiers, so code in
 * the canonical order must never spell "RN(a*b) + c" with these helpers unless
 * the product is exact (a power-of-two weight).  Every such place uses the
 * scalar __fmul_rn/__fadd_rn instead (pl_elevation_tile.cuh: upsampleMatrix[2]); an
 * fma chain whose first term is a product (acc = a0*b0; acc = fma(a1,b1,acc))
 * is safe: no instruction computes two products.  The parity tests compare
 * whole tiles bit for bit and would catch any contraction that changed a value.
 */
#ifndef PL_F2_CUH
#define PL_F2_CUH

#include "pl_fpexact.cuh"

namespace plf2 {

typedef float2 F2;

__device__ __forceinline__ F2 bc(float a) { return make_float2(a, a); }   /* broadcast operand (R.F32 in SASS) */
__device__ __forceinline__ F2 neg(F2 a) { return make_float2(-a.x, -a.y); }   /* folds into an operand modifier */
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ F2 mul2(F2 a, F2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ F2 add2(F2 a, F2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ F2 sub2(F2 a, F2 b) { return __fadd2_rn(a, neg(b)); }
__device__ __forceinline__ F2 max2(F2 a, F2 b) { return make_float2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
__device__ __forceinline__ F2 min2(F2 a, F2 b) { return make_float2(fminf(a.x, b.x), fminf(a.y, b.y)); }
__device__ __forceinline__ F2 clamp2(F2 x, float lo, float hi) { return min2(max2(x, bc(lo)), bc(hi)); }

/* the IEEE fast paths of pl_fpexact.cuh, two at a time */
__device__ __forceinline__ F2 rcp_rn2(F2 b)
{
    const F2 r = make_float2(plfp::rcp_seed(b.x), plfp::rcp_seed(b.y));
    const F2 e = fma2(r, neg(b), bc(1.0f));
    return fma2(r, e, r);
}
__device__ __forceinline__ F2 div_rn2(F2 a, F2 b, F2 rb)
{
    const F2 q = mul2(a, rb);
    const F2 rem = fma2(q, neg(b), a);
    return fma2(rb, rem, q);
}
__device__ __forceinline__ F2 sqrt_rn2(F2 x)
{
    const F2 r = make_float2(plfp::rsqrt_seed(fmaxf(x.x, 0x1p-100f)), plfp::rsqrt_seed(fmaxf(x.y, 0x1p-100f)));
    const F2 g = mul2(x, r);
    const F2 h = mul2(r, bc(0.5f));
    const F2 e = fma2(neg(g), g, x);
    return fma2(e, h, g);
}
/* canonical dot products (oracle/orc_fp.h R1): left-to-right fma chains */
__device__ __forceinline__ F2 dot3(F2 a0, F2 a1, F2 a2, F2 b0, F2 b1, F2 b2)
{
    return fma2(a2, b2, fma2(a1, b1, mul2(a0, b0)));
}
__device__ __forceinline__ F2 chain4(F2 a0, F2 a1, F2 a2, F2 a3, float w0, float w1, float w2, float w3)
{
    return fma2(a3, bc(w3), fma2(a2, bc(w2), fma2(a1, bc(w1), mul2(a0, bc(w0)))));
}

}  // namespace plf2
#endif
