// Fused fast paths above the opcode interface: the even-odd Schur complement and CG on its normal equation.
// Same arithmetic, order of operations and reduction precision as the Python they replace:
//   lib/gpt/algorithms/preconditioner/schur_complement_two.py:87-112  (_N, _N_dag)
//   lib/gpt/algorithms/preconditioner/normal_equation.py:44-45        (Mpc^dag Mpc)
//   lib/gpt/algorithms/inverter/cg.py:47-112                          (CG loop)
#include "operator.cuh"

namespace cgptb {

// o = i - Meooe MooeeInv Meooe MooeeInv i      (dag: o = i - MooeeInv^dag Meooe^dag MooeeInv^dag Meooe^dag i)
void op_schur_two(cgptb_fermion_operator* op, bool dag, const cgptb_lattice* in, cgptb_lattice* out) {
  CGPTB_ASSERT(in->cb != CGPTB_FULL && in->data != out->data);
  int D = in->cb, C = 1 - in->cb;
  cgptb_lattice* td = op->tmp(1, D);
  cgptb_lattice* tc0 = op->tmp(2, C);
  cgptb_lattice* tc1 = op->tmp(3, C);
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

}  // namespace cgptb

using namespace cgptb;

extern "C" {

int cgptb_apply_schur_two(cgptb_fermion_operator* op, int dag, const cgptb_lattice* in, cgptb_lattice* out) {
  CGPTB_API_BEGIN
  op_schur_two(op, dag != 0, in, out);
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
  cgptb_lattice *p = 0, *mmp = 0, *r = 0, *v = 0;
  cgptb_lattice** all[4] = {&p, &mmp, &r, &v};
  struct Guard {
    cgptb_lattice*** a;
    ~Guard() {
      for (int i = 0; i < 4; i++)
        if (*a[i]) cgptb_delete_lattice(*a[i]);
    }
  } guard{all};
  for (int i = 0; i < 4; i++)
    if (cgptb_create_lattice(all[i], op->dims4, op->Ls, op->prec, CGPTB_OT_VSPINCOLOR, src->cb)) CGPTB_ERR("%s", cgptb_last_error());

  struct GlobalSums {  // reductions inside the solver are global sums (cg.py gets them via grid.globalsum)
    GlobalSums() { g_reduce_global = true; }
    ~GlobalSums() { g_reduce_global = false; }
  } global_sums;
  auto mat = [&](cgptb_lattice* o, const cgptb_lattice* i) {
    op_schur_two(op, false, i, v);
    op_schur_two(op, true, v, o);
  };
  double n2;
  mat(mmp, psi);
  blas_axpy(r, -1.0, 0.0, mmp, src);
  blas_copy(p, r);
  if (cgptb_lattice_norm2(p, &n2)) CGPTB_ERR("%s", cgptb_last_error());
  double cp = n2;
  if (cgptb_lattice_norm2(src, &n2)) CGPTB_ERR("%s", cgptb_last_error());
  double ssq = n2;
  if (ssq == 0.0) {
    blas_zero(psi);
    return 0;
  }
  double rsq = eps * eps * ssq;
  for (int k = 0; k < maxiter; k++) {
    double c = cp;
    mat(mmp, p);
    double ip[3];
    const cgptb_lattice* l[1] = {p};
    const cgptb_lattice* rr[1] = {mmp};
    if (cgptb_lattice_rank_inner_product(l, 1, rr, 1, ip)) CGPTB_ERR("%s", cgptb_last_error());
    double d = ip[0];
    double a = c / d;
    if (cgptb_lattice_axpy_norm2(r, -a, 0.0, mmp, r, &cp)) CGPTB_ERR("%s", cgptb_last_error());
    double b = cp / c;
    double ca[2] = {a, 0.0};
    const cgptb_lattice* pp[1] = {p};
    blas_lc(psi, 1, 1, ca, pp);  // psi += a p
    blas_axpy(p, b, 0.0, p, r);  // p = b p + r
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
