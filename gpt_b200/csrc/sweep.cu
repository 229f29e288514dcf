// Stand-alone fifth-dimension sweep kernel (see sweep.cuh for the algebra).
//
// Spinor fields are stored in 32-byte blocks (common.cuh); a 16-byte "unit" k is half `k & 1` of block `k >> 1`.
// One thread owns one 16-byte unit of one 4d site for ALL s (Ls values in registers) and runs a short
// chain of stages, e.g. T = (b + c S5)(bee - cee S5)^-1 = "Meooe5D o MooeeInv" in a single pass over the field:
// 48 reals of traffic per site instead of 96 (+ the dense Ls x Ls product) of the unfused kernels.
//
// Pure HBM streaming, so it is written as a persistent kernel: CTAs loop over tiles of NSB sites, tile i+1 is
// fetched with cp.async (global -> shared, no register staging, s fastest = fully coalesced) while tile i is swept
// in registers and written back.
#include <string.h>
#include "sweep.cuh"

namespace cgptb {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}

// CG vector update fused in front of the sweep (cg.py:91-95 + the first factor of the next matrix application):
//   psi += a p ; p = b p + r ; out = T p         (in = p is updated in place)
template <typename T>
struct UpdateArgs {
  T a, b;
  const T* r;
  T* psi;
  T* p;
};

template <typename T, int LS, int NSB, bool UPD>
__global__ void __launch_bounds__(NSB* VecOf<T>::NB, 2) k_s_sweep(size_t n4, const T* __restrict__ in, T* __restrict__ out,
                                                               size_t stride, SweepParams<T> P, int ntiles, UpdateArgs<T> upd) {
  typedef typename VecOf<T>::type V;
  constexpr int NB = VecOf<T>::NB;
  constexpr int PITCH = LS + 1;  // vectors per (block, site) row in shared memory: conflict-free column reads
  constexpr int BUF = NB * NSB * PITCH;
  constexpr int NT = NSB * NB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V* sm = reinterpret_cast<V*>(smem_raw);
  const V* gin = reinterpret_cast<const V*>(in);
  V* gout = reinterpret_cast<V*>(out);

  auto prefetch = [&](int tile, V* buf) {
    size_t site0 = (size_t)tile * NSB;
    int nloc = (int)((n4 - site0) < (size_t)NSB ? (n4 - site0) : NSB);
#pragma unroll
    for (int it = 0; it < LS; it++) {
      int idx = threadIdx.x + it * NT;
      int half = idx & 1, q = idx >> 1;
      int kb = q / (NSB * LS), rem = q - kb * (NSB * LS);
      int k = 2 * kb + half;
      int l = rem / LS, s = rem - l * LS;
      if (l < nloc) cp_async16(buf + (k * NSB + l) * PITCH + s, gin + (((size_t)kb * stride + site0 * LS + rem) << 1) + half);
    }
    asm volatile("cp.async.commit_group;");
  };

  int tile = blockIdx.x;
  if (tile < ntiles) prefetch(tile, sm);
  int cur = 0;
  for (; tile < ntiles; tile += gridDim.x, cur ^= 1) {
    V* buf = sm + cur * BUF;
    int next = tile + gridDim.x;
    if (next < ntiles) {
      prefetch(next, sm + (cur ^ 1) * BUF);
      asm volatile("cp.async.wait_group 1;");
    } else {
      asm volatile("cp.async.wait_group 0;");
    }
    __syncthreads();
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
              V pv = buf[(k * NSB + l) * PITCH + s];
              __stcs(gpsi + o, vfma(upd.a, pv, sv[c]));
              V pn = vfma(upd.b, pv, rv[c]);
              __stcs(gp + o, pn);
              buf[(k * NSB + l) * PITCH + s] = pn;
            }
          }
        }
      }
      __syncthreads();
    }
    {
      int l = threadIdx.x % NSB, k = threadIdx.x / NSB;
      if (l < nloc) sweep_row<T, LS>(P, k, buf + (k * NSB + l) * PITCH);
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < LS; it++) {
      int idx = threadIdx.x + it * NT;
      int half = idx & 1, q = idx >> 1;
      int kb = q / (NSB * LS), rem = q - kb * (NSB * LS);
      int k = 2 * kb + half;
      int l = rem / LS, s = rem - l * LS;
      if (l < nloc) __stcs(gout + (((size_t)kb * stride + site0 * LS + rem) << 1) + half, buf[(k * NSB + l) * PITCH + s]);
    }
    __syncthreads();  // buf is refilled by the prefetch of the next iteration
  }
}

// sites per CTA: two buffers must fit into ~100 KB so that two CTAs share an SM
template <typename T, int LS>
constexpr int sweep_nsb() {
  int n = 32;
  while (n > 1 && (size_t)2 * VecOf<T>::NB * n * (LS + 1) * 16 > 100 * 1024) n /= 2;
  return n;
}

template <typename T, int LS, bool UPD>
static void launch_sweep(size_t n4, const T* in, T* out, size_t stride, const SweepParams<T>& P, const UpdateArgs<T>& upd) {
  constexpr int NSB = sweep_nsb<T, LS>();
  constexpr size_t smem = (size_t)2 * VecOf<T>::NB * NSB * (LS + 1) * 16;
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(k_s_sweep<T, LS, NSB, UPD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int ntiles = (int)((n4 + NSB - 1) / NSB);
  int per_sm = (int)(200 * 1024 / smem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int blocks = sm_count() * per_sm;
  if (blocks > ntiles) blocks = ntiles;
  k_s_sweep<T, LS, NSB, UPD><<<blocks, NSB * VecOf<T>::NB, smem, g_stream>>>(n4, in, out, stride, P, ntiles, upd);
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

// psi += a p ; p = b p + r ; t = T p   in one pass (CG update + first factor of the next Mpc)
template <typename T>
static bool cg_update_t(cgptb_fermion_operator* op, double a, double b, cgptb_lattice* p, const cgptb_lattice* r, cgptb_lattice* psi,
                        cgptb_lattice* t) {
  SweepParams<T> P;
  if (!make_sweep_params<T>(op, SWEEP_T, P)) return false;
  UpdateArgs<T> upd;
  upd.a = (T)a;
  upd.b = (T)b;
  upd.r = (const T*)r->data;
  upd.psi = (T*)psi->data;
  upd.p = (T*)p->data;
  if (!launch_sweep_ls<T, true>(op->Ls, p->sites / op->Ls, (const T*)p->data, (T*)t->data, p->sites, P, upd)) return false;
  LAUNCH_CHECK();
  return true;
}

bool op_cg_update_sweep(cgptb_fermion_operator* op, double a, double b, cgptb_lattice* p, const cgptb_lattice* r, cgptb_lattice* psi,
                        cgptb_lattice* t) {
  if (op->type != CGPTB_MOBIUS || !(op->Ls == 4 || op->Ls == 6 || op->Ls == 8 || op->Ls == 12 || op->Ls == 16 || op->Ls == 24)) return false;
  CGPTB_ASSERT(same_shape(p, r) && same_shape(p, psi) && same_shape(p, t));
  bool ok = op->prec == CGPTB_SINGLE ? cg_update_t<float>(op, a, b, p, r, psi, t) : cg_update_t<double>(op, a, b, p, r, psi, t);
  if (ok) t->cb = p->cb;
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
