// Single-precision Wilson / Moebius hopping kernel on the packed FFMA2 pipe (sm_100a).
//
// Same operator, layout and thread mapping as the generic k_dhop (operator.cu): one thread per output
// (4d site, s), the Ls threads of a 4d site share its eight links.  All complex arithmetic is done on
// (re, im) register pairs: a complex multiply-add is two FFMA2, spin projection and reconstruction are FADD2
// with swap / negate operand modifiers, so the kernel issues ~430 packed FP instructions per site instead
// of ~770 scalar ones -- which is what keeps the 5d fp32 stencil under the HBM roofline instead of the
// FP32 issue limit (SURVEY.md section 7 "hard parts").
#include "dslash.cuh"
#include "operator.cuh"
#include "packed.cuh"

namespace cgptb {

__device__ __forceinline__ void load_spinor_c32(const float* __restrict__ base, size_t nsites, size_t site, c32 (&p)[12]) {
  const float4* b = reinterpret_cast<const float4*>(base);
#pragma unroll
  for (int k = 0; k < 6; k++) {
    float4 v = __ldg(b + k * nsites + site);
    p[2 * k] = pk(v.x, v.y);
    p[2 * k + 1] = pk(v.z, v.w);
  }
}

__device__ __forceinline__ void store_spinor_c32(float* __restrict__ base, size_t nsites, size_t site, const c32 (&p)[12]) {
  float4* b = reinterpret_cast<float4*>(base);
#pragma unroll
  for (int k = 0; k < 6; k++) {
    float4 v;
    upk(p[2 * k], v.x, v.y);
    upk(p[2 * k + 1], v.z, v.w);
    b[k * nsites + site] = v;
  }
}

template <int MU, bool FWD, bool DAG>
__device__ __forceinline__ void hop_c32(c32 (&acc)[12], const Geom& g, int x, int y, int z, int t, int i4, int s, int ls,
                                        const float* __restrict__ in, size_t in_stride, const float* __restrict__ links) {
  const int SGN = (FWD != DAG) ? -1 : +1;
  typedef Proj<MU, SGN> P;
  int n4 = neighbor<MU, FWD>(g, x, y, z, t);
  c32 psi[12];
  load_spinor_c32(in, in_stride, (size_t)n4 * ls + s, psi);
  // link elements as (re, im) scalars
  float wr[9], wi[9];
  const float2* lb = reinterpret_cast<const float2*>(links) + ((size_t)i4 * 8 + (FWD ? MU : MU + 4)) * 9;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    float2 v = __ldg(lb + k);
    wr[k] = v.x;
    wi[k] = v.y;
  }
  c32 h[6];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    h[c] = add2(psi[c], times_iph<P::A>(psi[P::J0 * 3 + c]));
    h[3 + c] = add2(psi[3 + c], times_iph<P::B>(psi[P::J1 * 3 + c]));
  }
  c32 chi[6];
#pragma unroll
  for (int sp = 0; sp < 2; sp++) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
      c32 a;
      if (FWD) {  // W h
        a = cmul<false>(wr[r * 3 + 0], wi[r * 3 + 0], h[sp * 3 + 0]);
        a = cmac<false>(a, wr[r * 3 + 1], wi[r * 3 + 1], h[sp * 3 + 1]);
        a = cmac<false>(a, wr[r * 3 + 2], wi[r * 3 + 2], h[sp * 3 + 2]);
      } else {  // W^dag h
        a = cmul<true>(wr[0 * 3 + r], wi[0 * 3 + r], h[sp * 3 + 0]);
        a = cmac<true>(a, wr[1 * 3 + r], wi[1 * 3 + r], h[sp * 3 + 1]);
        a = cmac<true>(a, wr[2 * 3 + r], wi[2 * 3 + r], h[sp * 3 + 2]);
      }
      chi[sp * 3 + r] = a;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; c++) {
    acc[c] = add2(acc[c], chi[c]);
    acc[3 + c] = add2(acc[3 + c], chi[3 + c]);
    acc[6 + c] = add2(acc[6 + c], times_iph<P::C2>(chi[P::K2 * 3 + c]));
    acc[9 + c] = add2(acc[9 + c], times_iph<P::C3>(chi[P::K3 * 3 + c]));
  }
}

template <bool DAG, int LS>
__global__ void __launch_bounds__(128) k_dhop_f32(Geom g, int ls_rt, int p_out, const float* __restrict__ in, size_t in_stride,
                                                  float* __restrict__ out, size_t out_stride, const float* __restrict__ links) {
  const int ls = LS > 0 ? LS : ls_rt;
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (size_t)g.half4 * ls) return;
  int i4 = (int)(tid / ls);
  int s = (int)(tid - (size_t)i4 * ls);
  int x, y, z, t;
  cb_coords(g, p_out, i4, x, y, z, t);
  c32 acc[12];
#pragma unroll
  for (int k = 0; k < 12; k++) acc[k] = 0ull;
  hop_c32<0, true, DAG>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop_c32<0, false, DAG>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop_c32<1, true, DAG>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop_c32<1, false, DAG>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop_c32<2, true, DAG>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop_c32<2, false, DAG>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop_c32<3, true, DAG>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  hop_c32<3, false, DAG>(acc, g, x, y, z, t, i4, s, ls, in, in_stride, links);
  store_spinor_c32(out, out_stride, tid, acc);
}

template <bool DAG>
static void launch(int ls, unsigned blocks, int threads, const Geom& g, int p_out, const float* in, size_t is, float* out, size_t os,
                   const float* links) {
  switch (ls) {
    case 1: k_dhop_f32<DAG, 1><<<blocks, threads, 0, g_stream>>>(g, ls, p_out, in, is, out, os, links); break;
    case 8: k_dhop_f32<DAG, 8><<<blocks, threads, 0, g_stream>>>(g, ls, p_out, in, is, out, os, links); break;
    case 12: k_dhop_f32<DAG, 12><<<blocks, threads, 0, g_stream>>>(g, ls, p_out, in, is, out, os, links); break;
    case 16: k_dhop_f32<DAG, 16><<<blocks, threads, 0, g_stream>>>(g, ls, p_out, in, is, out, os, links); break;
    case 24: k_dhop_f32<DAG, 24><<<blocks, threads, 0, g_stream>>>(g, ls, p_out, in, is, out, os, links); break;
    default: k_dhop_f32<DAG, 0><<<blocks, threads, 0, g_stream>>>(g, ls, p_out, in, is, out, os, links); break;
  }
}

void dhop_half_f32(cgptb_fermion_operator* op, bool dag, const float* pin, size_t in_stride, float* pout, size_t out_stride,
                   int p_out) {
  int ls = op->ls();
  size_t half = (size_t)op->g.half4 * ls;
  int threads = 128;
  unsigned blocks = (unsigned)((half + threads - 1) / threads);
  if (dag)
    launch<true>(ls, blocks, threads, op->g, p_out, pin, in_stride, pout, out_stride, (const float*)op->links[p_out]);
  else
    launch<false>(ls, blocks, threads, op->g, p_out, pin, in_stride, pout, out_stride, (const float*)op->links[p_out]);
  LAUNCH_CHECK();
}

}  // namespace cgptb
