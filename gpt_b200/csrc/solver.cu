// Fused fast paths above the opcode interface: the even-odd Schur complement and CG on its normal equation.
// Same arithmetic, order of operations and reduction precision as the Python they replace:
//   lib/gpt/algorithms/preconditioner/schur_complement_two.py:87-112  (_N, _N_dag)
//   lib/gpt/algorithms/preconditioner/normal_equation.py:44-45        (Mpc^dag Mpc)
//   lib/gpt/algorithms/inverter/cg.py:47-112                          (CG loop)
//
// Moebius, single precision, one GPU: with T = (b + c S5)(bee - cee S5)^-1 = Meooe5D o MooeeInv
//   Mpc     x = x - Dhop T Dhop T x          = 1 sweep kernel + 2 Dslash kernels with fused epilogues
//   Mpc^dag x = x - T^dag Dhop^dag T^dag Dhop^dag x  = 2 Dslash kernels with fused epilogues
// (the epilogue applies T / T^dag to the stencil result while it is still on chip, subtracts from x and, for the
// last factor inside CG, accumulates <p, A p>), i.e. 5 passes over the field per Mpc^dag Mpc instead of 18.
#include <stdlib.h>
#include "operator.cuh"

namespace cgptb {

bool dhop_fusable(const cgptb_fermion_operator* op);  // dslash_f32.cu
int dhop_tile_blocks(const cgptb_fermion_operator* op);
void dhop_half_f32_fused(cgptb_fermion_operator* op, bool dag, const float* pin, size_t in_stride, float* pout, size_t out_stride,
                         int p_out, int sweep_mode, const float* z, size_t z_stride, const float* dotp, size_t dot_stride,
                         double* partial);
bool sweep_supported(int ls);  // sweep.cu
bool op_cg_update_sweep(cgptb_fermion_operator* op, double a, double b, cgptb_lattice* p, const cgptb_lattice* r, cgptb_lattice* psi,
                        cgptb_lattice* t, const double* ab_dev = 0);
double* blas_axpy_norm2_dev(cgptb_lattice* r, const double* scal_dev, double mult, const cgptb_lattice* x, const cgptb_lattice* y);  // blas.cu
extern bool g_reduce_to_device;
extern double* g_reduce_dev_ptr;

// Scalars of the CG on the device (cg.py:70-112): the loop runs without a host round trip per iteration; the host reads `done`
// and `iter` every few iterations.  After convergence a = 0, so the iterations that were already queued leave psi and r alone.
struct CgState {
  double a, b;  // (a, b) contiguous: the fused update kernel reads them as ab_dev[0..1]
  double c, rsq;
  int done, iter;
};
// mode 0: red[0] = <p, A p>  ->  a = c / d
// mode 1: red[0] = |r|^2     ->  b = cp / c ; history[iter++] = |cp| ; done if |cp| <= rsq ; c = cp
__global__ void k_cg_step(CgState* S, double* hist, int mode, const double* red) {
  if (mode == 0) {
    S->a = S->done ? 0.0 : S->c / red[0];
  } else {
    const double cp = red[0];
    if (S->done) {
      S->b = 1.0;
    } else {
      S->b = cp / S->c;
      const double res = fabs(cp);
      hist[S->iter] = res;
      S->iter++;
      if (res <= S->rsq) S->done = 1;
      S->c = cp;
    }
  }
}
bool op_s_sweep_sub_dot(cgptb_fermion_operator* op, int mode, const cgptb_lattice* in, const cgptb_lattice* z, cgptb_lattice* out,
                        const cgptb_lattice* dotp, double* dot);
void blas_finalize(int nblocks, int ncomp, const double* partial, double* host_out);  // blas.cu
double* blas_partial_scratch(int nblocks);

static void fused_dhop(cgptb_fermion_operator* op, bool dag, const cgptb_lattice* in, cgptb_lattice* out, int sweep_mode,
                       const cgptb_lattice* z, const cgptb_lattice* dotp, double* partial) {
  out->cb = 1 - in->cb;
  dhop_half_f32_fused(op, dag, (const float*)in->data, in->sites, (float*)out->data, out->sites, out->cb, sweep_mode,
                      z ? (const float*)z->data : 0, z ? z->sites : 0, dotp ? (const float*)dotp->data : 0,
                      dotp ? dotp->sites : 0, partial);
}

// o = i - Meooe MooeeInv Meooe MooeeInv i      (dag: o = i - MooeeInv^dag Meooe^dag MooeeInv^dag Meooe^dag i)
// dot (optional, 3 doubles): re<dotp,o>, im<dotp,o>, |o|^2 (global sums inside the solver)
// t_in (optional): T in, already computed by the fused CG update (non-dag only)
void op_schur_two(cgptb_fermion_operator* op, bool dag, const cgptb_lattice* in, cgptb_lattice* out, const cgptb_lattice* dotp,
                  double* dot, const cgptb_lattice* t_in = 0) {
  CGPTB_ASSERT(in->cb != CGPTB_FULL && in->data != out->data);
  op->check_field(in);
  op->check_field(out);
  int D = in->cb, C = 1 - in->cb;
  cgptb_lattice* td = op->tmp(1, D);
  cgptb_lattice* tc0 = op->tmp(2, C);
  cgptb_lattice* tc1 = op->tmp(3, C);
  static int no_sweep = getenv("CGPTB_NO_SWEEP") ? 1 : 0;
  bool done_dot = false;
  if (op->type == CGPTB_MOBIUS && !op->zmobius && !no_sweep && sweep_supported(op->Ls)) {
    if (dhop_fusable(op)) {
      double* partial = dot ? blas_partial_scratch(dhop_tile_blocks(op)) : 0;
      if (!dag) {
        if (t_in)
          td = const_cast<cgptb_lattice*>(t_in);
        else
          op_s_sweep(op, SWEEP_T, in, td);                                 // T x
        fused_dhop(op, false, td, tc1, SWEEP_T, 0, 0, 0);                  // T Dhop (.)         D -> C
        fused_dhop(op, false, tc1, out, -1, in, dot ? dotp : 0, partial);  // x - Dhop (.)       C -> D
      } else {
        fused_dhop(op, true, in, tc1, SWEEP_TDAG, 0, 0, 0);                       // T^dag Dhop^dag x           D -> C
        fused_dhop(op, true, tc1, out, SWEEP_TDAG, in, dot ? dotp : 0, partial);  // x - T^dag Dhop^dag (.)     C -> D
      }
      if (dot) {
        blas_finalize(dhop_tile_blocks(op), 3, partial, dot);
        done_dot = true;
      }
    } else {
      // Meooe MooeeInv = Dhop o T applied as one register-resident sweep + the plain stencil (the TMA sweep kernel where it
      // applies: decomposed lattices, compressed links); the last sweep of Mpc^dag subtracts from x and accumulates <p, A p>
      if (!dag) {
        if (t_in)
          td = const_cast<cgptb_lattice*>(t_in);
        else
          op_s_sweep(op, SWEEP_T, in, td);
        op_dhop(op, false, td, tc0);
        op_s_sweep(op, SWEEP_T, tc0, tc1);
        op_dhop(op, false, tc1, out);
        blas_axpy(out, -1.0, 0.0, out, in);
      } else {
        op_dhop(op, true, in, tc0);
        op_s_sweep(op, SWEEP_TDAG, tc0, tc1);
        op_dhop(op, true, tc1, td);
        if (!op_s_sweep_sub_dot(op, SWEEP_TDAG, td, in, out, dotp, dot)) CGPTB_ERR("sweep kernel missing for Ls=%d", op->Ls);
        if (dot) done_dot = true;
      }
    }
  } else {
    if (!dag) {
      op_mooee(op, true, false, false, in, td);    // DD^-1
      op_meooe(op, false, td, tc0);                // CD
      op_mooee(op, true, false, false, tc0, tc1);  // CC^-1
      op_meooe(op, false, tc1, out);               // DC
    } else {
      op_meooe(op, true, in, tc0);                 // DC^dag
      op_mooee(op, true, true, false, tc0, tc1);   // CC^-dag
      op_meooe(op, true, tc1, td);                 // CD^dag
      op_mooee(op, true, true, false, td, out);    // DD^-dag
    }
    blas_axpy(out, -1.0, 0.0, out, in);            // gpt.axpy(o_d, -1.0, o_d, i_d)
  }
  if (dot && !done_dot) {
    double a2;
    if (cgptb_lattice_inner_product_norm2(dotp, out, dot, &a2)) CGPTB_ERR("%s", cgptb_last_error());
    dot[2] = 0.0;
  }
}

}  // namespace cgptb

using namespace cgptb;

extern "C" {

int cgptb_apply_schur_two(cgptb_fermion_operator* op, int dag, const cgptb_lattice* in, cgptb_lattice* out) {
  CGPTB_API_BEGIN
  op_schur_two(op, dag != 0, in, out, 0, 0, 0);
  CGPTB_API_END
}

int cgptb_cg_eo2_ne(cgptb_fermion_operator* op, cgptb_lattice* psi, const cgptb_lattice* src, double eps, int maxiter,
                    double* history, int* iterations, int* converged) {
  CGPTB_API_BEGIN
  op->check_field(psi);
  op->check_field(src);
  CGPTB_ASSERT(src->cb != CGPTB_FULL && psi->sites == src->sites && psi->data != src->data);
  psi->cb = src->cb;
  *iterations = 0;
  *converged = 0;
  // work fields: kept in the operator between solves (allocating and freeing five fields of the lattice's size per solve costs
  // tens to hundreds of milliseconds in cudaFree alone and showed as a "slow first solve" / bimodal time per iteration)
  static int no_upd = getenv("CGPTB_NO_FUSED_UPDATE") ? 1 : 0;
  static int no_sweep_cg = getenv("CGPTB_NO_SWEEP") ? 1 : 0;
  bool fuse_update = !no_upd && !no_sweep_cg && op->type == CGPTB_MOBIUS && !op->zmobius && sweep_supported(op->Ls);
  cgptb_lattice* w[5];
  for (int i = 0; i < (fuse_update ? 5 : 4); i++) {
    if (!op->cg_half[i] && cgptb_create_lattice(&op->cg_half[i], op->dims4, op->Ls, op->prec, CGPTB_OT_VSPINCOLOR, src->cb)) CGPTB_ERR("%s", cgptb_last_error());
    op->cg_half[i]->cb = src->cb;
    w[i] = op->cg_half[i];
  }
  cgptb_lattice *p = w[0], *mmp = w[1], *r = w[2], *v = w[3], *tp = fuse_update ? w[4] : 0;

  struct GlobalSums {  // reductions inside the solver are global sums (cg.py gets them via grid.globalsum)
    GlobalSums() { g_reduce_global = true; }
    ~GlobalSums() { g_reduce_global = false; }
  } global_sums;
  // o = Mpc^dag Mpc i ; d3 (optional) = <i, o> fused into the last kernel
  bool have_tp = false;  // tp = T p from the fused update of the previous iteration
  auto mat = [&](cgptb_lattice* o, const cgptb_lattice* i, double* d3) {
    op_schur_two(op, false, i, v, 0, 0, (have_tp && i == p) ? tp : 0);
    op_schur_two(op, true, v, o, d3 ? i : 0, d3, 0);
  };
  double n2;
  mat(mmp, psi, 0);
  blas_axpy(r, -1.0, 0.0, mmp, src);
  blas_copy(p, r);
  if (cgptb_lattice_norm2(p, &n2)) CGPTB_ERR("%s", cgptb_last_error());
  double cp = n2;
  if (cgptb_lattice_norm2(src, &n2)) CGPTB_ERR("%s", cgptb_last_error());
  double ssq = n2;
  if (ssq == 0.0) {
    // cg.py:67-69: a zero source gives psi = 0 and returns silently -- not a convergence failure
    blas_zero(psi);
    *converged = 1;
    return 0;
  }
  double rsq = eps * eps * ssq;
  // device-resident scalars (Moebius with the fused update + sweep kernels; CGPTB_CG_HOST=1: the host loop below)
  static int cg_host = getenv("CGPTB_CG_HOST") ? 1 : 0;
  static int cg_chunk = getenv("CGPTB_CG_CHUNK") ? atoi(getenv("CGPTB_CG_CHUNK")) : 6;
  if (fuse_update && !cg_host && maxiter > 0) {
    CgState* S = 0;
    double* hist = 0;
    CUDA_CHECK(cudaMalloc(&S, sizeof(CgState)));
    CUDA_CHECK(cudaMalloc(&hist, sizeof(double) * (size_t)maxiter));
    struct Free {
      void *a, *b;
      ~Free() {
        cudaFree(a);
        cudaFree(b);
      }
    } fr{S, hist};
    CgState h;
    h.a = h.b = 0.0;
    h.c = cp;
    h.rsq = rsq;
    h.done = 0;
    h.iter = 0;
    CUDA_CHECK(cudaMemcpyAsync(S, &h, sizeof(h), cudaMemcpyHostToDevice, g_stream));
    CUDA_CHECK(cudaStreamSynchronize(g_stream));  // h is a stack object
    int k = 0;
    while (k < maxiter) {
      const int chunk = cg_chunk < 1 ? 1 : (maxiter - k < cg_chunk ? maxiter - k : cg_chunk);
      for (int j = 0; j < chunk; j++) {
        double dummy[3];
        g_reduce_to_device = true;
        g_reduce_dev_ptr = 0;
        try {
          mat(mmp, p, dummy);  // <p, mmp> stays on the device
        } catch (...) {
          g_reduce_to_device = false;
          throw;
        }
        g_reduce_to_device = false;
        if (!g_reduce_dev_ptr) CGPTB_ERR("device CG: the matrix application left no reduction on the device");
        k_cg_step<<<1, 1, 0, g_stream>>>(S, hist, 0, g_reduce_dev_ptr);
        LAUNCH_CHECK();
        double* cp_dev = blas_axpy_norm2_dev(r, &S->a, -1.0, mmp, r);
        k_cg_step<<<1, 1, 0, g_stream>>>(S, hist, 1, cp_dev);
        LAUNCH_CHECK();
        if (!op_cg_update_sweep(op, 0.0, 0.0, p, r, psi, tp, &S->a)) CGPTB_ERR("device CG: fused update kernel missing");
        have_tp = true;
      }
      k += chunk;
      CUDA_CHECK(cudaMemcpyAsync(&h, S, sizeof(h), cudaMemcpyDeviceToHost, g_stream));
      CUDA_CHECK(cudaStreamSynchronize(g_stream));
      if (h.done) break;
    }
    *iterations = h.iter;
    *converged = h.done;
    if (history && h.iter > 0) CUDA_CHECK(cudaMemcpy(history, hist, sizeof(double) * (size_t)h.iter, cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaStreamSynchronize(g_stream));
    return 0;
  }
  for (int k = 0; k < maxiter; k++) {
    double c = cp;
    double ip[3];
    mat(mmp, p, ip);  // d = <p, mmp>.real
    double d = ip[0];
    double a = c / d;
    if (cgptb_lattice_axpy_norm2(r, -a, 0.0, mmp, r, &cp)) CGPTB_ERR("%s", cgptb_last_error());
    double b = cp / c;
    static int cg_timing = getenv("CGPTB_CG_TIMING") ? 1 : 0;  // measurement aid: device time of the fused update kernel
    static cudaEvent_t tev[2];
    static double tacc = 0;
    static int tn = 0;
    if (cg_timing) {
      if (!tn && tacc == 0) {
        cudaEventCreate(&tev[0]);
        cudaEventCreate(&tev[1]);
      }
      cudaEventRecord(tev[0], g_stream);
    }
    const bool fused_upd = fuse_update && op_cg_update_sweep(op, a, b, p, r, psi, tp);
    if (cg_timing) {
      cudaEventRecord(tev[1], g_stream);
      cudaEventSynchronize(tev[1]);
      float ms = 0;
      cudaEventElapsedTime(&ms, tev[0], tev[1]);
      tacc += ms;
      if (++tn % 50 == 0) fprintf(stderr, "[cg timing] update step: %.3f ms (fused %d, %d calls)\n", tacc / tn, (int)fused_upd, tn);
    }
    if (fused_upd) {
      have_tp = true;  // psi += a p ; p = b p + r ; tp = T p in one pass
    } else {
      double ca[2] = {a, 0.0};
      const cgptb_lattice* pp[1] = {p};
      blas_lc(psi, 1, 1, ca, pp);  // psi += a p
      blas_axpy(p, b, 0.0, p, r);  // p = b p + r
      have_tp = false;
    }
    double res = fabs(cp);
    if (history) history[k] = res;
    *iterations = k + 1;
    if (res <= rsq) {
      *converged = 1;
      break;
    }
  }
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
  CGPTB_API_END
}
}
