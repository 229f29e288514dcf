// Packed complex arithmetic for sm_100a: one complex fp32 number lives in an aligned 64-bit register pair
// (re, im) and is processed by the Blackwell FFMA2 / FADD2 / FMUL2 pipe (PTX fma/add/mul.rn.f32x2).
// ptxas folds the pack / unpack / swap / negate moves written below into FFMA2 operand modifiers
// (R.F32x2.HI_LO, -R.F32x2.LO_HI.NP, R.F32 broadcast), so a complex multiply-accumulate is exactly two
// FFMA2 instructions (checked with cuobjdump -sass; see DESIGN.md "FFMA2").
#pragma once
#include <cuda_runtime.h>

namespace cgptb {

typedef unsigned long long c32;  // (lo = re, hi = im)

__device__ __forceinline__ c32 pk(float re, float im) {
  c32 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(re), "f"(im));
  return r;
}
__device__ __forceinline__ void upk(c32 v, float& re, float& im) { asm("mov.b64 {%0, %1}, %2;" : "=f"(re), "=f"(im) : "l"(v)); }
__device__ __forceinline__ c32 fma2(c32 a, c32 b, c32 c) {
  c32 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ c32 mul2(c32 a, c32 b) {
  c32 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c32 add2(c32 a, c32 b) {
  c32 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c32 bcast(float x) { return pk(x, x); }

// z * i^PH as a pure operand modifier (swap halves / negate lanes)
template <int PH>
__device__ __forceinline__ c32 times_iph(c32 z) {
  float re, im;
  upk(z, re, im);
  if (PH == 0) return z;
  if (PH == 1) return pk(-im, re);
  if (PH == 2) return pk(-re, -im);
  return pk(im, -re);
}

// acc + u * h  (CONJ: conj(u) * h), u given as separate (re, im) scalars
template <bool CONJ>
__device__ __forceinline__ c32 cmac(c32 acc, float ur, float ui, c32 h) {
  acc = fma2(h, bcast(ur), acc);
  // u*h: + ui * (-h.im, h.re) ; conj(u)*h: + ui * (h.im, -h.re)
  return fma2(CONJ ? times_iph<3>(h) : times_iph<1>(h), bcast(ui), acc);
}
template <bool CONJ>
__device__ __forceinline__ c32 cmul(float ur, float ui, c32 h) {
  c32 acc = mul2(h, bcast(ur));
  return fma2(CONJ ? times_iph<3>(h) : times_iph<1>(h), bcast(ui), acc);
}

}  // namespace cgptb
