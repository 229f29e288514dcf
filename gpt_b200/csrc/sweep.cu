// Stand-alone fifth-dimension sweep kernel (see sweep.cuh for the algebra).
//
// Spinor fields are stored in 32-byte blocks (common.cuh); a 16-byte "unit" k is half `k & 1` of block `k >> 1`.
// One thread owns one 16-byte unit of one 4d site for ALL s (Ls values in registers) and runs a short
// chain of stages, e.g. T = (b + c S5)(bee - cee S5)^-1 = "Meooe5D o MooeeInv" in a single pass over the field:
// 48 reals of traffic per site instead of 96 (+ the dense Ls x Ls product) of the unfused kernels.
//
// Pure HBM streaming, so it is written as a persistent kernel: CTAs loop over tiles of NSB sites, tile i+1 is fetched with
// bulk copies (cp.async.bulk global -> shared, one per (32-byte component plane, site): the Ls x 32 B of a site are contiguous;
// completion on an mbarrier, no register staging) while tile i is swept in registers and written back.  (Until round 2 the
// tile was fetched with 16-byte cp.async: LDGSTS does not merge the two halves of a 32-byte sector, 2.4 GB came through L2 for
// 1.2 GB of data and the kernel ran at 4.6 TB/s.)
#include <stdlib.h>
#include <string.h>
#include "sweep.cuh"

namespace cgptb {

// d[0..1] += conj(a) * b summed over the complex numbers of a 16-byte unit, d[2] += |b|^2, in double
__device__ __forceinline__ void vdot_acc(float4 a, float4 b, double (&d)[3]) {
  d[0] += (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z + (double)a.w * b.w;
  d[1] += (double)a.x * b.y - (double)a.y * b.x + (double)a.z * b.w - (double)a.w * b.z;
  d[2] += (double)b.x * b.x + (double)b.y * b.y + (double)b.z * b.z + (double)b.w * b.w;
}
__device__ __forceinline__ void vdot_acc(double2 a, double2 b, double (&d)[3]) {
  d[0] += a.x * b.x + a.y * b.y;
  d[1] += a.x * b.y - a.y * b.x;
  d[2] += b.x * b.x + b.y * b.y;
}

// CG vector update fused in front of the sweep (cg.py:91-95 + the first factor of the next matrix application):
//   psi += a p ; p = b p + r ; out = T p         (in = p is updated in place)
template <typename T>
struct UpdateArgs {
  T a, b;
  const double* ab_dev;  // if set: a = ab_dev[0], b = ab_dev[1] (computed on the device by the CG)
  const T* r;
  T* psi;
  T* p;
};

// Epilogue behind the sweep (the last factor of Mpc^dag inside CG, schur_complement_two.py:100-112 + cg.py:83):
//   out = z - S in ; partial[cta] = (re, im) <dotp, out>, |out|^2 in double
template <typename T>
struct SubDotArgs {
  const T* z;
  const T* dotp;   // may be 0: no reduction
  double* partial; // [ctas][3]
};

__device__ __forceinline__ uint32_t sw_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sw_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; spin++) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();
  }
}

template <typename T, int LS, int NSB, bool UPD, bool EPI = false>
__global__ void __launch_bounds__(NSB* VecOf<T>::NB, 2) k_s_sweep(size_t n4, const T* __restrict__ in, T* __restrict__ out,
                                                               size_t stride, SweepParams<T> P, int ntiles, UpdateArgs<T> upd,
                                                               SubDotArgs<T> epi = SubDotArgs<T>()) {
  typedef typename VecOf<T>::type V;
  constexpr int NB = VecOf<T>::NB;
  constexpr int NPL = NB / 2;            // 32-byte component planes
  constexpr int ROWB = LS * 32 + 16;     // bytes of one (plane, site) row in shared memory: Ls x 32 B + 16 B, so that the
                                         // 16-byte column reads of 8 consecutive sites fall into different banks
  constexpr int BUFB = NPL * NSB * ROWB;
  constexpr int NT = NSB * NB;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sm = smem_raw;
  const uint32_t bar0 = sw_smem_u32(smem_raw + 2 * BUFB);
  V* gout = reinterpret_cast<V*>(out);
  // element (16-byte unit k = 2 * plane + half, site l, slice s) of a buffer
  auto at = [&](unsigned char* buf, int k, int l, int s) -> V& {
    return *reinterpret_cast<V*>(buf + ((k >> 1) * NSB + l) * ROWB + s * 32 + (k & 1) * 16);
  };
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto prefetch = [&](int tile, int b) {
    size_t site0 = (size_t)tile * NSB;
    int nloc = (int)((n4 - site0) < (size_t)NSB ? (n4 - site0) : NSB);
    const uint32_t bar = bar0 + 8 * b;
    // the buffer was read and written through the generic proxy by the previous tile: order that before the bulk copies
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(NPL * nloc * LS * 32)) : "memory");
    if (EPI) {
      // z and dotp of that tile are read element-wise after the sweep: pull them into L2 now (no registers held)
#pragma unroll
      for (int it = 0; it < LS; it++) {
        int idx = threadIdx.x + it * NT;
        int q = idx >> 1;
        int kb = q / (NSB * LS), rem = q - kb * (NSB * LS);
        if (!(idx & 1) && rem / LS < nloc) {  // one prefetch per 32-byte sector
          size_t o = ((size_t)kb * stride + site0 * LS + rem) << 1;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const V*>(epi.z) + o));
          if (epi.dotp) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const V*>(epi.dotp) + o));
        }
      }
    }
    for (int r = threadIdx.x; r < NPL * NSB; r += NT) {
      const int kb = r / NSB, l = r - kb * NSB;
      if (l < nloc) {
        const T* src = in + ((size_t)kb * stride + (site0 + l) * LS) * (32 / sizeof(T));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         sw_smem_u32(sm + (size_t)b * BUFB + (kb * NSB + l) * ROWB)),
                     "l"(src), "r"((uint32_t)(LS * 32)), "r"(bar)
                     : "memory");
      }
    }
  };

  if (UPD && upd.ab_dev) {
    upd.a = (T)upd.ab_dev[0];
    upd.b = (T)upd.ab_dev[1];
  }
  double dsum[3] = {0.0, 0.0, 0.0};
  int tile = blockIdx.x;
  if (tile < ntiles) prefetch(tile, 0);
  int cur = 0;
  uint32_t phase[2] = {0, 0};
  for (; tile < ntiles; tile += gridDim.x, cur ^= 1) {
    unsigned char* buf = sm + (size_t)cur * BUFB;
    int next = tile + gridDim.x;
    if (next < ntiles) prefetch(next, cur ^ 1);
    sw_mbar_wait(bar0 + 8 * cur, phase[cur]);
    phase[cur] ^= 1u;
    size_t site0 = (size_t)tile * NSB;
    int nloc = (int)((n4 - site0) < (size_t)NSB ? (n4 - site0) : NSB);
    if (UPD) {
      const V* gr = reinterpret_cast<const V*>(upd.r);
      V* gpsi = reinterpret_cast<V*>(upd.psi);
      V* gp = reinterpret_cast<V*>(upd.p);
      constexpr int CH = 4;  // loads in flight per thread and field: CH x 16 bytes
#pragma unroll
      for (int it0 = 0; it0 < LS; it0 += CH) {
        V rv[CH], sv[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) {
          int it = it0 + c;
          if (it < LS) {
            int idx = threadIdx.x + it * NT;
            int half = idx & 1, q = idx >> 1;
            int kb = q / (NSB * LS), rem = q - kb * (NSB * LS);
            int l = rem / LS;
            if (l < nloc) {
              size_t o = (((size_t)kb * stride + site0 * LS + rem) << 1) + half;
              rv[c] = __ldcs(gr + o);
              sv[c] = __ldcs(gpsi + o);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < CH; c++) {
          int it = it0 + c;
          if (it < LS) {
            int idx = threadIdx.x + it * NT;
            int half = idx & 1, q = idx >> 1;
            int kb = q / (NSB * LS), rem = q - kb * (NSB * LS);
            int k = 2 * kb + half;
            int l = rem / LS, s = rem - l * LS;
            if (l < nloc) {
              size_t o = (((size_t)kb * stride + site0 * LS + rem) << 1) + half;
              V pv = at(buf, k, l, s);
              __stcs(gpsi + o, vfma(upd.a, pv, sv[c]));
              V pn = vfma(upd.b, pv, rv[c]);
              __stcs(gp + o, pn);
              at(buf, k, l, s) = pn;
            }
          }
        }
      }
      __syncthreads();
    }
    {
      int l = threadIdx.x % NSB, k = threadIdx.x / NSB;
      if (l < nloc) sweep_row<T, LS, 2>(P, k, &at(buf, k, l, 0));
    }
    __syncthreads();
    if (EPI) {
      // out = z - (swept), partial sums of <dotp, out> and |out|^2: the loads of z and dotp go out CH at a time
      constexpr int CH = 4;
      const V* gz = reinterpret_cast<const V*>(epi.z);
      const V* gd = reinterpret_cast<const V*>(epi.dotp);
#pragma unroll
      for (int it0 = 0; it0 < LS; it0 += CH) {
        V zv[CH], dv[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) {
          int it = it0 + c;
          if (it < LS) {
            int idx = threadIdx.x + it * NT;
            int half = idx & 1, q = idx >> 1;
            int kb = q / (NSB * LS), rem = q - kb * (NSB * LS);
            int l = rem / LS;
            if (l < nloc) {
              size_t o = (((size_t)kb * stride + site0 * LS + rem) << 1) + half;
              zv[c] = __ldcs(gz + o);
              if (gd) dv[c] = __ldcs(gd + o);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < CH; c++) {
          int it = it0 + c;
          if (it < LS) {
            int idx = threadIdx.x + it * NT;
            int half = idx & 1, q = idx >> 1;
            int kb = q / (NSB * LS), rem = q - kb * (NSB * LS);
            int k = 2 * kb + half;
            int l = rem / LS, s = rem - l * LS;
            if (l < nloc) {
              size_t o = (((size_t)kb * stride + site0 * LS + rem) << 1) + half;
              V r = vfma((T)-1, at(buf, k, l, s), zv[c]);
              if (gd) vdot_acc(dv[c], r, dsum);
              __stcs(gout + o, r);
            }
          }
        }
      }
    } else {
#pragma unroll
      for (int it = 0; it < LS; it++) {
        int idx = threadIdx.x + it * NT;
        int half = idx & 1, q = idx >> 1;
        int kb = q / (NSB * LS), rem = q - kb * (NSB * LS);
        int k = 2 * kb + half;
        int l = rem / LS, s = rem - l * LS;
        if (l < nloc) __stcs(gout + (((size_t)kb * stride + site0 * LS + rem) << 1) + half, at(buf, k, l, s));
      }
    }
    __syncthreads();  // buf is refilled by the prefetch of the next iteration
  }
  if (EPI && epi.dotp) {
    __shared__ double red[96];
    block_reduce<3>(dsum, red);
    if (threadIdx.x == 0) {
      epi.partial[blockIdx.x * 3 + 0] = dsum[0];
      epi.partial[blockIdx.x * 3 + 1] = dsum[1];
      epi.partial[blockIdx.x * 3 + 2] = dsum[2];
    }
  }
}

// sites per CTA: two buffers must fit into ~100 KB so that two CTAs share an SM
template <typename T, int LS>
constexpr int sweep_nsb() {
  int n = 32;
  while (n > 1 && (size_t)2 * VecOf<T>::NB * n * (LS + 1) * 16 > 100 * 1024) n /= 2;
  return n;
}

// two tile buffers of (NB / 2 planes) x NSB rows of Ls x 32 B + 16 B, two mbarriers, alignment slack
template <typename T, int LS>
constexpr size_t sweep_smem() {
  return (size_t)2 * (VecOf<T>::NB / 2) * sweep_nsb<T, LS>() * (LS * 32 + 16) + 128;
}

template <typename T, int LS>
static int sweep_blocks(size_t n4) {
  constexpr int NSB = sweep_nsb<T, LS>();
  constexpr size_t smem = sweep_smem<T, LS>();
  int ntiles = (int)((n4 + NSB - 1) / NSB);
  int per_sm = (int)(200 * 1024 / smem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int blocks = sm_count() * per_sm;
  if (blocks > ntiles) blocks = ntiles;
  return blocks;
}

template <typename T, int LS, bool UPD, bool EPI = false>
static void launch_sweep(size_t n4, const T* in, T* out, size_t stride, const SweepParams<T>& P, const UpdateArgs<T>& upd,
                         const SubDotArgs<T>& epi = SubDotArgs<T>()) {
  constexpr int NSB = sweep_nsb<T, LS>();
  constexpr size_t smem = sweep_smem<T, LS>();
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(k_s_sweep<T, LS, NSB, UPD, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int ntiles = (int)((n4 + NSB - 1) / NSB);
  int blocks = sweep_blocks<T, LS>(n4);
  k_s_sweep<T, LS, NSB, UPD, EPI><<<blocks, NSB * VecOf<T>::NB, smem, g_stream>>>(n4, in, out, stride, P, ntiles, upd, epi);
}

// out = z - S in with optional <dotp, out>, |out|^2; returns the number of CTAs (= rows of partial) or 0 if Ls has no kernel
template <typename T>
static int launch_sweep_epi(int ls, size_t n4, const T* in, T* out, size_t stride, const SweepParams<T>& P, const SubDotArgs<T>& epi) {
  UpdateArgs<T> none;
  memset(&none, 0, sizeof(none));
  switch (ls) {
    case 4: launch_sweep<T, 4, false, true>(n4, in, out, stride, P, none, epi); return sweep_blocks<T, 4>(n4);
    case 6: launch_sweep<T, 6, false, true>(n4, in, out, stride, P, none, epi); return sweep_blocks<T, 6>(n4);
    case 8: launch_sweep<T, 8, false, true>(n4, in, out, stride, P, none, epi); return sweep_blocks<T, 8>(n4);
    case 12: launch_sweep<T, 12, false, true>(n4, in, out, stride, P, none, epi); return sweep_blocks<T, 12>(n4);
    case 16: launch_sweep<T, 16, false, true>(n4, in, out, stride, P, none, epi); return sweep_blocks<T, 16>(n4);
    case 24: launch_sweep<T, 24, false, true>(n4, in, out, stride, P, none, epi); return sweep_blocks<T, 24>(n4);
    default: return 0;
  }
}

template <typename T, bool UPD>
static bool launch_sweep_ls(int ls, size_t n4, const T* in, T* out, size_t stride, const SweepParams<T>& P, const UpdateArgs<T>& upd) {
  switch (ls) {
    case 4: launch_sweep<T, 4, UPD>(n4, in, out, stride, P, upd); break;
    case 6: launch_sweep<T, 6, UPD>(n4, in, out, stride, P, upd); break;
    case 8: launch_sweep<T, 8, UPD>(n4, in, out, stride, P, upd); break;
    case 12: launch_sweep<T, 12, UPD>(n4, in, out, stride, P, upd); break;
    case 16: launch_sweep<T, 16, UPD>(n4, in, out, stride, P, upd); break;
    case 24: launch_sweep<T, 24, UPD>(n4, in, out, stride, P, upd); break;
    default: return false;
  }
  return true;
}

template <typename T>
static bool sweep_t(cgptb_fermion_operator* op, int mode, const cgptb_lattice* in, cgptb_lattice* out) {
  int ls = op->Ls;
  SweepParams<T> P;
  if (!make_sweep_params<T>(op, mode, P)) return false;
  size_t n4 = in->sites / ls;
  const T* pin = (const T*)in->data;
  T* pout = (T*)out->data;
  UpdateArgs<T> none;
  memset(&none, 0, sizeof(none));
  if (!launch_sweep_ls<T, false>(ls, n4, pin, pout, in->sites, P, none)) return false;
  LAUNCH_CHECK();
  return true;
}

__global__ void k_nop() {}

// psi += a p ; p = b p + r ; t = T p   in one pass (CG update + first factor of the next Mpc)
template <typename T>
static bool cg_update_t(cgptb_fermion_operator* op, double a, double b, cgptb_lattice* p, const cgptb_lattice* r, cgptb_lattice* psi,
                        cgptb_lattice* t, const double* ab_dev) {
  SweepParams<T> P;
  if (!make_sweep_params<T>(op, SWEEP_T, P)) return false;
  UpdateArgs<T> upd;
  upd.a = (T)a;
  upd.b = (T)b;
  upd.ab_dev = ab_dev;
  upd.r = (const T*)r->data;
  upd.psi = (T*)psi->data;
  upd.p = (T*)p->data;
  if (!launch_sweep_ls<T, true>(op->Ls, p->sites / op->Ls, (const T*)p->data, (T*)t->data, p->sites, P, upd)) return false;
  LAUNCH_CHECK();
  static int nop = getenv("CGPTB_UPD_NOP") ? atoi(getenv("CGPTB_UPD_NOP")) : 0;
  if (nop == 1) k_nop<<<1, 32, 0, g_stream>>>();
  if (nop == 2) k_nop<<<sm_count() * 2, 1024, 0, g_stream>>>();
  return true;
}

bool op_cg_update_sweep(cgptb_fermion_operator* op, double a, double b, cgptb_lattice* p, const cgptb_lattice* r, cgptb_lattice* psi,
                        cgptb_lattice* t, const double* ab_dev) {
  if (op->type != CGPTB_MOBIUS || !(op->Ls == 4 || op->Ls == 6 || op->Ls == 8 || op->Ls == 12 || op->Ls == 16 || op->Ls == 24)) return false;
  CGPTB_ASSERT(same_shape(p, r) && same_shape(p, psi) && same_shape(p, t));
  bool ok = op->prec == CGPTB_SINGLE ? cg_update_t<float>(op, a, b, p, r, psi, t, ab_dev) : cg_update_t<double>(op, a, b, p, r, psi, t, ab_dev);
  if (ok) t->cb = p->cb;
  return ok;
}

bool sweep_supported(int ls);
double* blas_partial_scratch(int nblocks);  // blas.cu
void blas_finalize(int nblocks, int ncomp, const double* partial, double* host_out);

// out = z - S in (S = the fused fifth-dimension operator `mode`); dot (optional, 3 doubles): re, im of <dotp, out> and |out|^2,
// global sums inside the solver.  false if this operator / Ls has no sweep kernel.
template <typename T>
static bool sweep_sub_dot_t(cgptb_fermion_operator* op, int mode, const cgptb_lattice* in, const cgptb_lattice* z, cgptb_lattice* out,
                            const cgptb_lattice* dotp, double* dot) {
  SweepParams<T> P;
  if (!make_sweep_params<T>(op, mode, P)) return false;
  const int ls = op->Ls;
  const size_t n4 = in->sites / ls;
  SubDotArgs<T> epi;
  epi.z = (const T*)z->data;
  epi.dotp = dot ? (const T*)dotp->data : 0;
  epi.partial = dot ? blas_partial_scratch(sm_count() * 4) : 0;
  const int ctas = launch_sweep_epi<T>(ls, n4, (const T*)in->data, (T*)out->data, in->sites, P, epi);
  if (!ctas) return false;
  LAUNCH_CHECK();
  if (dot) blas_finalize(ctas, 3, epi.partial, dot);
  return true;
}

bool op_s_sweep_sub_dot(cgptb_fermion_operator* op, int mode, const cgptb_lattice* in, const cgptb_lattice* z, cgptb_lattice* out,
                        const cgptb_lattice* dotp, double* dot) {
  if (op->type != CGPTB_MOBIUS || op->zmobius || !sweep_supported(op->Ls)) return false;
  CGPTB_ASSERT(same_shape(in, z) && same_shape(in, out) && (!dot || same_shape(in, dotp)) && in->data != out->data);
  bool ok = op->prec == CGPTB_SINGLE ? sweep_sub_dot_t<float>(op, mode, in, z, out, dotp, dot) : sweep_sub_dot_t<double>(op, mode, in, z, out, dotp, dot);
  if (ok) out->cb = in->cb;
  return ok;
}

bool sweep_supported(int ls) { return ls == 4 || ls == 6 || ls == 8 || ls == 12 || ls == 16 || ls == 24; }

// returns false if this Ls has no sweep kernel (the caller falls back to the tridiagonal + dense kernels)
bool op_s_sweep(cgptb_fermion_operator* op, int mode, const cgptb_lattice* in, cgptb_lattice* out) {
  op->check_field(in);
  op->check_field(out);
  CGPTB_ASSERT(op->type == CGPTB_MOBIUS && in->sites == out->sites);
  if (op->zmobius || !sweep_supported(op->Ls)) return false;  // zMoebius: complex coefficients, dense kernel instead
  bool ok = op->prec == CGPTB_SINGLE ? sweep_t<float>(op, mode, in, out) : sweep_t<double>(op, mode, in, out);
  if (ok) out->cb = in->cb;
  return ok;
}

}  // namespace cgptb
