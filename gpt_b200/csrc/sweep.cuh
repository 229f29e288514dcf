// Fifth-dimension operators of the Moebius action as register-resident sweeps (shared by the stand-alone
// kernel in sweep.cu and the epilogue of the fused Dslash in dslash_f32.cu).
//
// Every s-direction operator of the action is, per chirality, a cyclic bidiagonal matrix
//     B = d + f S ,   S = lower shift (couples s-1) with corner S[0][Ls-1] = -m      (P+ of S5, P- of S5^dag)
//                     or upper shift (couples s+1) with corner S[Ls-1][0] = -m       (P- of S5, P+ of S5^dag)
// (M5D shape lib/cgpt/lib/foundation/mobius_with_vector_field.h:36-96).  Applying B costs
// 2 flops per element; solving B y = x is a forward sweep plus a rank-one correction for the corner
// (the LDU sweep of mobius_with_vector_field.h:185-302 in closed form):
//     u_0 = x_0/d, u_s = (x_s - f u_{s-1})/d ;  w_0 = f m/d, w_s = -(f/d) w_{s-1} ;
//     y_L = u_L / (1 - w_L) ;  y_s = u_s + y_L w_s .
#pragma once
#include <vector>
#include "operator.cuh"

namespace cgptb {

static const int MAXLS = 32;

template <typename T>
struct SweepStage {
  int solve;      // 0: y = B x ; 1: y = B^-1 x
  int dir[2];     // per chirality block (0: upper spin components = P+, 1: lower = P-): 0 lower shift, 1 upper shift
  T d, f;         // B = d + f S
  T id, fid;      // 1/d, f/d
  T m[2];         // corner masses per chirality
  T g[2];         // 1 / (1 - w_last) per chirality
  T w[2][MAXLS];  // corner response per chirality
};

template <typename T>
struct SweepParams {
  int nstages;
  SweepStage<T> st[2];
};

template <typename T>
struct VecOf;
template <>
struct VecOf<float> {
  typedef float4 type;
  static const int NB = 6;   // 16-byte blocks per spinor
  static const int NBU = 3;  // blocks of the upper (P+) chirality
};
template <>
struct VecOf<double> {
  typedef double2 type;
  static const int NB = 12;
  static const int NBU = 6;
};

// a*x + y with scalar a
__device__ __forceinline__ float4 vfma(float a, float4 x, float4 y) {
  return make_float4(fmaf(a, x.x, y.x), fmaf(a, x.y, y.y), fmaf(a, x.z, y.z), fmaf(a, x.w, y.w));
}
__device__ __forceinline__ double2 vfma(double a, double2 x, double2 y) { return make_double2(fma(a, x.x, y.x), fma(a, x.y, y.y)); }
__device__ __forceinline__ float4 vmul(float a, float4 x) { return make_float4(a * x.x, a * x.y, a * x.z, a * x.w); }
__device__ __forceinline__ double2 vmul(double a, double2 x) { return make_double2(a * x.x, a * x.y); }

template <typename T, int LS>
__device__ __forceinline__ void run_stage(const SweepStage<T>& st, int chi, typename VecOf<T>::type (&x)[LS]) {
  typedef typename VecOf<T>::type V;
  const int L = LS - 1;
  const int dir = st.dir[chi];
  if (!st.solve) {
    V first = x[0], last = x[L];
    if (dir == 0) {  // y_s = d x_s + f x_{s-1}, y_0 = d x_0 - f m x_L
#pragma unroll
      for (int s = L; s >= 1; s--) x[s] = vfma(st.f, x[s - 1], vmul(st.d, x[s]));
      x[0] = vfma(-st.f * st.m[chi], last, vmul(st.d, first));
    } else {  // y_s = d x_s + f x_{s+1}, y_L = d x_L - f m x_0
#pragma unroll
      for (int s = 0; s < L; s++) x[s] = vfma(st.f, x[s + 1], vmul(st.d, x[s]));
      x[L] = vfma(-st.f * st.m[chi], first, vmul(st.d, last));
    }
  } else {
    if (dir == 0) {
      x[0] = vmul(st.id, x[0]);
#pragma unroll
      for (int s = 1; s <= L; s++) x[s] = vfma(-st.fid, x[s - 1], vmul(st.id, x[s]));
      V yl = vmul(st.g[chi], x[L]);
      x[L] = yl;
#pragma unroll
      for (int s = 0; s < L; s++) x[s] = vfma(st.w[chi][s], yl, x[s]);
    } else {
      x[L] = vmul(st.id, x[L]);
#pragma unroll
      for (int s = L - 1; s >= 0; s--) x[s] = vfma(-st.fid, x[s + 1], vmul(st.id, x[s]));
      V y0 = vmul(st.g[chi], x[0]);
      x[0] = y0;
#pragma unroll
      for (int s = 1; s <= L; s++) x[s] = vfma(st.w[chi][s], y0, x[s]);
    }
  }
}

// the sweep of one (component block k, site l) row held in shared memory: row[s], s = 0..LS-1
// (STRIDE: distance of consecutive s in units of the 16-byte vector type)
template <typename T, int LS, int STRIDE = 1>
__device__ __forceinline__ void sweep_row(const SweepParams<T>& P, int k, typename VecOf<T>::type* row) {
  typename VecOf<T>::type x[LS];
#pragma unroll
  for (int s = 0; s < LS; s++) x[s] = row[s * STRIDE];
  int chi = k < VecOf<T>::NBU ? 0 : 1;
  run_stage<T, LS>(P.st[0], chi, x);
  if (P.nstages > 1) run_stage<T, LS>(P.st[1], chi, x);
#pragma unroll
  for (int s = 0; s < LS; s++) row[s * STRIDE] = x[s];
}

// stage description for B = d + f S5 (dag: S5^dag), solve or apply
template <typename T>
static SweepStage<T> make_stage(int ls, bool solve, bool dag, double d, double f, double mp, double mm) {
  SweepStage<T> st;
  st.solve = solve ? 1 : 0;
  // S5: P+ couples s-1 (lower), P- couples s+1 (upper); the dagger swaps them
  st.dir[0] = dag ? 1 : 0;
  st.dir[1] = dag ? 0 : 1;
  st.d = (T)d;
  st.f = (T)f;
  st.id = (T)(1.0 / d);
  st.fid = (T)(f / d);
  double m[2] = {mp, mm};
  for (int chi = 0; chi < 2; chi++) {
    st.m[chi] = (T)m[chi];
    std::vector<double> w(ls);
    int L = ls - 1;
    if (st.dir[chi] == 0) {
      w[0] = f * m[chi] / d;
      for (int s = 1; s <= L; s++) w[s] = -(f / d) * w[s - 1];
      st.g[chi] = (T)(1.0 / (1.0 - w[L]));
    } else {
      w[L] = f * m[chi] / d;
      for (int s = L - 1; s >= 0; s--) w[s] = -(f / d) * w[s + 1];
      st.g[chi] = (T)(1.0 / (1.0 - w[0]));
    }
    for (int s = 0; s < MAXLS; s++) st.w[chi][s] = s < ls ? (T)w[s] : (T)0;
  }
  return st;
}

// mode: SWEEP_T = (b + c S5)(bee - cee S5)^-1, SWEEP_TDAG its adjoint, SWEEP_MINV(DAG) = (bee - cee S5)^-1 (adjoint)
template <typename T>
static bool make_sweep_params(const cgptb_fermion_operator* op, int mode, SweepParams<T>& P) {
  int ls = op->Ls;
  double b = op->p.b, c = op->p.c, mp = op->p.mass_plus, mm = op->p.mass_minus;
  double bee = b * (4.0 - op->p.M5) + 1.0, cee = 1.0 - c * (4.0 - op->p.M5);
  switch (mode) {
    case SWEEP_T:
      P.nstages = 2;
      P.st[0] = make_stage<T>(ls, true, false, bee, -cee, mp, mm);
      P.st[1] = make_stage<T>(ls, false, false, b, c, mp, mm);
      return true;
    case SWEEP_TDAG:
      P.nstages = 2;
      P.st[0] = make_stage<T>(ls, false, true, b, c, mp, mm);
      P.st[1] = make_stage<T>(ls, true, true, bee, -cee, mp, mm);
      return true;
    case SWEEP_MINV:
    case SWEEP_MINVDAG:
      P.nstages = 1;
      P.st[0] = make_stage<T>(ls, true, mode == SWEEP_MINVDAG, bee, -cee, mp, mm);
      P.st[1] = P.st[0];
      return true;
    default:
      return false;
  }
}

}  // namespace cgptb
