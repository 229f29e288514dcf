/*
  cgpt_b200.h -- C ABI of libcgpt_b200.so: the B200-native replacement for the fermion-operator hot path of
  GPT's `cgpt` module (lehner/gpt).  Every entry point names the cgpt export it replaces
  (file:line relative to the reference checkout).

  Conventions (mirroring cgpt):
    * handles are opaque pointers; create_* returns an owning handle, delete_* frees it;
      operator inputs are borrowed for the call only, except U which is copied at create/update.
    * every call returns 0 on success and non-zero on failure; cgptb_last_error() then returns the message
      (cgpt: C++ throw std::string -> PyExc_RuntimeError, lib/cgpt/lib/exception.h:23-39).
    * calls are made from one host thread per GPU; work is enqueued on the library stream and the call
      returns without synchronising unless it has to hand a number back to the host.
    * host-side field layout ("GPT order"): site index lexicographic with dimension 0 fastest (5d grids:
      s is dimension 0), checkerboarded fields hold the sites of that parity in the same order; per site the
      tensor elements row-major, complex interleaved (numpy complex64 / complex128).  This is the order of
      `lattice[:]` in GPT (lib/cgpt/lib/lattice/implementation.h:246-281).
*/
#ifndef CGPT_B200_H
#define CGPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cgptb_lattice cgptb_lattice;
typedef struct cgptb_fermion_operator cgptb_fermion_operator;

enum { CGPTB_SINGLE = 0, CGPTB_DOUBLE = 1 };
enum { CGPTB_EVEN = 0, CGPTB_ODD = 1, CGPTB_FULL = 2 };
/* object types (lib/gpt/core/object_type): complex singlet, colour matrix, spin-colour vector */
enum { CGPTB_OT_SINGLET = 1, CGPTB_OT_MCOLOR = 9, CGPTB_OT_VSPINCOLOR = 12 };
enum { CGPTB_WILSON_CLOVER = 0, CGPTB_MOBIUS = 1 };

/* opcodes: lib/cgpt/lib/operators/register.h:2-20 */
enum {
  CGPTB_OP_M = 2001, CGPTB_OP_Mdag = 2002, CGPTB_OP_Meooe = 2003, CGPTB_OP_MeooeDag = 2004,
  CGPTB_OP_Mooee = 2005, CGPTB_OP_MooeeDag = 2006, CGPTB_OP_MooeeInv = 2007, CGPTB_OP_MooeeInvDag = 2008,
  CGPTB_OP_Mdiag = 2009, CGPTB_OP_Dminus = 2010, CGPTB_OP_DminusDag = 2011,
  CGPTB_OP_ImportPhysicalFermionSource = 2012, CGPTB_OP_ImportUnphysicalFermion = 2013,
  CGPTB_OP_ExportPhysicalFermionSolution = 2014, CGPTB_OP_ExportPhysicalFermionSource = 2015,
  CGPTB_OP_Dhop = 3001, CGPTB_OP_DhopEO = 3002, CGPTB_OP_DhopDag = 4001, CGPTB_OP_DhopEODag = 4002
};

/* ---- runtime ------------------------------------------------------------------------------------ */
/* cgpt.init (lib/cgpt/lib/init.cc:26-103): select the CUDA device, create the library stream.       */
int cgptb_init(int device);
const char* cgptb_last_error(void);
/* cgpt.accelerator_barrier (lib/cgpt/lib/util.cc:335-338)                                            */
int cgptb_accelerator_barrier(void);
/* make the library enqueue on an externally owned cudaStream_t (e.g. torch's current stream); 0 = own */
int cgptb_set_stream(void* cuda_stream);
void* cgptb_get_stream(void);
/* CUDA-event stopwatch on the library stream (cgpt.time is host wall clock, lib/cgpt/lib/time.cc:79-81) */
int cgptb_timer_start(void);
int cgptb_timer_stop(double* milliseconds);
int cgptb_device_info(int* sm_count, size_t* total_mem, int* cc_major, int* cc_minor);
/* number of kernels this library has launched since init (for bench.py's gpu_launches) */
uint64_t cgptb_launch_count(void);

/* ---- processor grid (one process per GPU; GPT's --mpi X.Y.Z.T, lib/gpt/core/grid.py:77-94) ---------------- */
/* rank 0 creates the 128-byte NCCL id, the host layer broadcasts it (torch.distributed / MPI), every rank then
   calls cgptb_comm_init with the processor grid mpi = {1, Y, Z, T} (x cannot be split; s never is).  From then
   on lattice extents passed to this library are LOCAL extents and fermion operators exchange halos.        */
int cgptb_comm_unique_id(char* id128);
int cgptb_comm_init(int rank, int world, const int mpi[4], const char* id128);
int cgptb_comm_finalize(void);
int cgptb_comm_info(int* rank, int* world, int pgrid[4], int pcoor[4]);
/* cgpt.grid_globalsum (lib/cgpt/lib/grid.cc:119-160): in-place sum over ranks of n doubles in host memory */
int cgptb_comm_globalsum(double* host, int n);

/* ---- lattices (storage seam; cgpt.create_lattice & co., lib/cgpt/lib/lattice.cc:42-210) ----------- */
/* dims4 = local x,y,z,t extents (all even); Ls = 0 for a 4d grid, else the extent of the 5th dimension
   (grid dimension 0, never checkerboarded: lib/gpt/core/grid.py:31-37); cb = CGPTB_EVEN/ODD for a field
   on the red-black grid, CGPTB_FULL otherwise.                                                        */
int cgptb_create_lattice(cgptb_lattice** out, const int dims4[4], int Ls, int precision, int otype, int cb);
/* wrap externally owned device memory (e.g. a torch tensor's data_ptr) of cgptb_lattice_bytes() bytes */
int cgptb_create_lattice_view(cgptb_lattice** out, const int dims4[4], int Ls, int precision, int otype, int cb,
                              void* device_ptr);
int cgptb_delete_lattice(cgptb_lattice* l);
/* gpt.pack of n 4d fields <-> one 5d field whose fifth dimension is the list index (matrix_operator.packed(),
   lib/gpt/core/operator/matrix_operator.py:200-239); unpack != 0: 5d -> the n 4d fields */
int cgptb_lattice_pack_rhs(cgptb_lattice* l5, cgptb_lattice* const* l4, int n, int unpack);
size_t cgptb_lattice_bytes(const cgptb_lattice* l);
size_t cgptb_lattice_sites(const cgptb_lattice* l);
void* cgptb_lattice_device_ptr(cgptb_lattice* l);
/* geometry of a lattice (what cgpt_Lattice_base::get_grid / to_decl expose, lib/cgpt/lib/lattice/implementation.h:50-54): lets the
   host layer allocate "a lattice like this one" (cgpt.eval with dst = None) without tracking grids itself */
int cgptb_lattice_info(const cgptb_lattice* l, int dims4[4], int* Ls, int* precision, int* otype, int* cb);
/* cgpt.lattice_get_checkerboard / lattice_change_checkerboard (lattice.cc:189-210) */
int cgptb_lattice_get_checkerboard(const cgptb_lattice* l);
int cgptb_lattice_change_checkerboard(cgptb_lattice* l, int cb);
/* cgpt.lattice_set_to_number with 0 (`lattice[:] = 0`), lattice.cc:86-99 */
int cgptb_lattice_set_to_zero(cgptb_lattice* l);
/* `lattice[:] = ndarray` / `ndarray = lattice[:]` through cgpt.lattice_memory_view
   (lattice/implementation.h:246-281): host buffer in GPT order, nbytes = sites*otype*2*sizeof(real)   */
int cgptb_lattice_import(cgptb_lattice* l, const void* host, size_t nbytes);
int cgptb_lattice_export(const cgptb_lattice* l, void* host, size_t nbytes);
/* same, but the buffer is DEVICE memory in GPT order (no PCIe copy) */
int cgptb_lattice_import_device(cgptb_lattice* l, const void* dev, size_t nbytes);
int cgptb_lattice_export_device(const cgptb_lattice* l, void* dev, size_t nbytes);
/* cgpt.copy / cgpt.convert (lib/cgpt/lib/transform.cc:41-54,128-141) */
int cgptb_lattice_copy(cgptb_lattice* dst, const cgptb_lattice* src);
int cgptb_lattice_convert(cgptb_lattice* dst, const cgptb_lattice* src);
/* cgpt.lattice_pick_checkerboard / lattice_set_checkerboard (lattice.cc:164-187) */
int cgptb_lattice_pick_checkerboard(int cb, cgptb_lattice* half, const cgptb_lattice* full);
int cgptb_lattice_set_checkerboard(cgptb_lattice* full, const cgptb_lattice* half);

/* ---- vector kernels ------------------------------------------------------------------------------- */
/* cgpt.lattice_axpy: r = a*x + y, a complex cast to field precision
   (transform.cc:211-246, foundation/transform.h:233-248).  No synchronisation (accelerator_forNB).   */
int cgptb_lattice_axpy(cgptb_lattice* r, double a_re, double a_im, const cgptb_lattice* x, const cgptb_lattice* y);
/* g.axpy_norm2 (lib/gpt/core/transform.py:151-153) fused: r = a*x + y ; *norm2 = |r|^2 (double accumulate) */
int cgptb_lattice_axpy_norm2(cgptb_lattice* r, double a_re, double a_im, const cgptb_lattice* x,
                             const cgptb_lattice* y, double* norm2);
/* cgpt.lattice_rank_inner_product (transform.cc:143-178; foundation/reduce.h:76-253): result[i*n_right+j] =
   sum conj(left_i) right_j accumulated in complex double for both precisions (reduce.h:129); rank-local. */
int cgptb_lattice_rank_inner_product(const cgptb_lattice* const* left, int n_left, const cgptb_lattice* const* right,
                                     int n_right, double* result_re_im);
/* cgpt.lattice_norm2 (transform.cc:199-209) */
int cgptb_lattice_norm2(const cgptb_lattice* a, double* norm2);
/* cgpt.lattice_inner_product_norm2 (transform.cc:180-197): ip = <a,b>, a2 = |a|^2 in one pass */
int cgptb_lattice_inner_product_norm2(const cgptb_lattice* a, const cgptb_lattice* b, double* ip_re_im, double* a2);
/* cgpt.eval restricted to linear combinations: dst (+)= sum_i c_i a_i
   (lib/cgpt/lib/eval.cc:323-365, expression/linear_combination_implementation.h:154-192)               */
int cgptb_lattice_lc(cgptb_lattice* dst, int accumulate, int n, const double* coef_re_im,
                     const cgptb_lattice* const* a);
/* cgpt.linear_combination (lib/cgpt/lib/basis.cc:146-173, foundation/basis.h:21-97):
   r_i = sum_k Qt[i*n_basis+k] basis_k, Qt complex128 row-major                                          */
int cgptb_linear_combination(cgptb_lattice* const* r, int n_r, const cgptb_lattice* const* basis, int n_basis,
                             const double* Qt_re_im);
/* lattice *= a  (cgpt.eval with a single scaled term onto itself) */
int cgptb_lattice_scale(cgptb_lattice* l, double a_re, double a_im);
/* g.slice(g.trace(a * g.adj(b)), 3) for spin-colour vectors: out[t] = sum_{x,y,z} <b(x,t), a(x,t)> (complex),
   the building block of the pion correlator in README.md:163-166; out has 2*dims4[3] doubles            */
int cgptb_lattice_slice_inner_product(const cgptb_lattice* b, const cgptb_lattice* a, double* out_re_im);

/* ---- fermion operators (lib/cgpt/lib/operators.cc:34-107) ------------------------------------------ */
typedef struct {
  /* Wilson-clover: lib/cgpt/lib/operators/wilson_clover.h:28-39 */
  double mass, csw_r, csw_t, cF, xi_0, nu;
  int isAnisotropic;
  /* Moebius: lib/cgpt/lib/operators/mobius.h:42-51 */
  double mass_plus, mass_minus, M5, b, c;
  int Ls;
  /* both: 4 complex boundary phases (re,im) */
  double boundary_phases[8];
  /* wilson_twisted_mass (lib/cgpt/lib/operators/wilson_twisted_mass.h): Mooee = (4 + mass) + i mu gamma_5; 0 = untwisted */
  double mu;
  /* zmobius (lib/cgpt/lib/operators/zmobius.h:20-56): n_omega = Ls complex omega_s (re,im); 0 = plain Moebius */
  int n_omega;
  double omega[2 * 64];
  /* extension (no counterpart in the reference's parameter tables): 12 = keep only the first two rows of every SU(3) link in
     the stencil's link tables and rebuild the third in registers (row_2 = f conj(row_0 x row_1), f = the U(1) factor the stored
     link carries: -c_mu/2 and the boundary phase); 0 = all 18 reals */
  int link_compression;
} cgptb_fermion_params;

/* cgpt.create_fermion_operator(optype, prec, params): U = 4 colour-matrix lattices on the full 4d grid */
int cgptb_create_fermion_operator(cgptb_fermion_operator** out, int optype, int precision,
                                  const cgptb_fermion_params* params, const cgptb_lattice* const U[4]);
/* cgpt.update_fermion_operator: re-import the gauge field */
int cgptb_update_fermion_operator(cgptb_fermion_operator* op, const cgptb_lattice* const U[4]);
/* cgpt.set_mass_fermion_operator (operators/implementation.h:21-45) */
int cgptb_set_mass_fermion_operator(cgptb_fermion_operator* op, const cgptb_fermion_params* params);
int cgptb_delete_fermion_operator(cgptb_fermion_operator* op);
/* cgpt.apply_fermion_operator(op, opcode, src, dst) -- note (src, dst) order (operators.cc:96-107) */
int cgptb_apply_fermion_operator(cgptb_fermion_operator* op, int opcode, const cgptb_lattice* src, cgptb_lattice* dst);

/* The same call on HOST buffers (full fields in GPT order, operator precision, nbytes each): replaces the sequence
   lattice[:] = src (cgpt.lattice_import_view, lib/gpt/core/lattice.py:229-260) ; cgpt.apply_fermion_operator ;
   dst = lattice[:] (cgpt.lattice_export_view).  For Dhop / DhopDag in single precision the upload, the stencil and
   the download are pipelined over slabs of time slices (gpt_b200/csrc/hostpipe.cu); every other opcode runs
   import -> apply -> export.  Synchronous.                                                               */
int cgptb_apply_fermion_operator_host(cgptb_fermion_operator* op, int opcode, const void* src_host, void* dst_host, size_t nbytes);

/* d = G s for a constant 4 x 4 complex spin matrix G (row-major, (re,im)): g.gamma[...] * field (lib/gpt/core/gamma.py:28-80) */
int cgptb_lattice_spin_matrix(cgptb_lattice* d, const cgptb_lattice* s, const double* m_re_im);
/* gpt.scale_per_coordinate(d, s, a, dim) (lib/gpt/core/transform.py:210-214, cgpt.lattice_scale_per_coordinate): d = a[x_dim] s,
   a = n complex factors (re,im), dim counts the fifth dimension as 0 on 5d lattices */
int cgptb_lattice_scale_per_coordinate(cgptb_lattice* d, const cgptb_lattice* s, const double* a_re_im, int n, int dim);

/* work schedule of the TMA sweep kernel (gpt_b200/csrc/dslash_tma.cu) as a table, host only: item i = out6[6 i ..] = (chunk group,
   x/2 origin, y origin, z origin, first time slice, number of time slices); item i runs on CTA i % grid.  For tests / tuning. */
int cgptb_debug_tma_schedule(const int dims4[4], int Ls, int grid, int chunks_per_cta, int sched, int trl, int t_begin, int t_count,
                             int* out6, int max_items, int* n_items);

/* ---- generic matrix-vector stencils: cgpt.stencil_matrix_vector_create / _execute / _delete
   (lib/cgpt/lib/stencil.cc:41-58,101-121 and stencil/matrix_vector.h:20-290; used by benchmarks/stencil.py:91-145 and
   tests/core/stencil.py:168-215).  points = n_points shifts of 4 ints; code line i = code_ints[5 i ..] = (target, accumulate
   or -1, source, source_point, number of factors) with weight (re, im); its factors follow each other in `factors` as
   (matrix field index, point, adjoint) and are applied right to left:
     vector[target](x) = weight M_1 ... M_n vector[source](x + p_source) [+ vector[accumulate](x)].
   Lines of a block of code_parallel_block_size run in order, blocks are independent.  Colour-matrix and (spin-)colour-vector
   fields on the full 4d lattice of one rank.                                                               */
typedef struct cgptb_stencil_mv cgptb_stencil_mv;
int cgptb_stencil_matrix_vector_create(cgptb_stencil_mv** out, const int dims4[4], int precision, int n_points, const int* points,
                                       int n_code, const int* code_ints, const double* weights_re_im, const int* factors,
                                       int code_parallel_block_size, int local, int matrix_parity, int vector_parity);
int cgptb_stencil_matrix_vector_execute(cgptb_stencil_mv* s, const cgptb_lattice* const* matrix_fields, int n_m,
                                        cgptb_lattice* const* vector_fields, int n_v, int fast_osites);
int cgptb_stencil_matrix_vector_delete(cgptb_stencil_mv* s);

/* ---- random numbers: cgpt.create_random(engine, seed) / cgpt.random_sample(rng, params) / cgpt.delete_random
   (lib/cgpt/lib/random.cc:38-101, random/engine.h:64-125).  Same streams as the reference: RANLUX24 lanes seeded by
   SHA-256, one generator per 2^4 block of sites, values drawn in double.  engine: "vectorized_ranlux24_389_64" (default
   of gpt.random) or "vectorized_ranlux24_24_64".  grid_key tells grid objects apart: the reference keeps one set of
   generators per (rng, grid object), so two grids of equal shape each restart from the seed.                        */
typedef struct cgptb_random cgptb_random;
enum { CGPTB_DIST_NORMAL = 0, CGPTB_DIST_CNORMAL = 1, CGPTB_DIST_UNIFORM_REAL = 2, CGPTB_DIST_UNIFORM_INT = 3, CGPTB_DIST_ZN = 4 };
int cgptb_create_random(cgptb_random** out, const char* engine, const char* seed);
int cgptb_delete_random(cgptb_random* r);
/* p0, p1 = (mu, sigma) for normal / cnormal, (min, max) for uniform_real / uniform_int, (n, -) for zn */
int cgptb_random_sample_scalar(cgptb_random* r, int dist, double p0, double p1, double out[2]);
/* fill a host array out[site][nel][re,im] (site lexicographic, dimension 0 fastest; nd <= 5, dimension 0 of a 5d grid
   is the unblocked fifth dimension); ldims / gdims / lstart: local extents, global extents, global origin of the local block */
int cgptb_random_sample_host(cgptb_random* r, uint64_t grid_key, int nd, const int* ldims, const int* gdims, const int* lstart,
                             int nel, int dist, double p0, double p1, double* out);
/* the same into a (full) lattice */
int cgptb_random_sample(cgptb_random* r, uint64_t grid_key, cgptb_lattice* l, int dist, double p0, double p1);
/* g.qcd.gauge.random(grid, rng, scale) (lib/gpt/qcd/gauge/create.py:66-71): U_mu = exp(i scale sum_a u_a T_a), u_a ~ U[-1/2,1/2) */
int cgptb_random_su3_links(cgptb_random* r, uint64_t grid_key, cgptb_lattice* const U[4], double scale);

/* g.qcd.gauge.plaquette(U) (lib/gpt/qcd/gauge/stencil/plaquette.py:23-44) and the NERSC link trace (lib/gpt/core/io/nersc_io.py:238-247):
   out[0] = <Re tr P_{mu nu}> / Nc over the six planes, out[1] = <Re tr U_mu> / Nc over the four directions */
int cgptb_gauge_plaquette(const cgptb_lattice* const U[4], double out[2]);

/* NERSC gauge configurations (lib/gpt/core/io/nersc_io.py:146-199): the data part of the file -> four link lattices on
   the device: byte order (cgpt.munge_byte_order), third-row reconstruction for 4D_SU3_GAUGE (cgpt.munge_reconstruct_third_row,
   lib/cgpt/lib/munge.h:21-39), [site][mu] -> [mu][site] (cgpt.munge_inner_outer), checksum (cgpt.util_nersc_checksum,
   lib/cgpt/lib/checksums/nersc.h:19-33) in one kernel.  rows = 2 (4D_SU3_GAUGE) or 3 (4D_SU3_GAUGE_3x3)        */
int cgptb_nersc_munge(const void* raw_host, size_t nbytes, int float_size, int big_endian, int rows, cgptb_lattice* const U[4],
                      unsigned int* checksum);

/* ---- fused fast paths (same results as the opcode sequences they replace) ---------------------------- */
/* Mpc / Mpc^dag of schur_complement_two (lib/gpt/algorithms/preconditioner/schur_complement_two.py:87-112):
   o = i - Meooe MooeeInv Meooe MooeeInv i ; tmp = 2 work fields of the same shape                       */
int cgptb_apply_schur_two(cgptb_fermion_operator* op, int dag, const cgptb_lattice* in, cgptb_lattice* out);
/* inv.cg on Mpc^dag Mpc (lib/gpt/algorithms/inverter/cg.py:47-112 on normal_equation.py:44-45) run entirely on the
   device: same update order, reductions in double, residual test every iteration; history receives
   |r|^2 per iteration (cg.history), *iterations the count.                                               */
int cgptb_cg_eo2_ne(cgptb_fermion_operator* op, cgptb_lattice* psi, const cgptb_lattice* src, double eps,
                    int maxiter, double* history, int* iterations, int* converged);

#ifdef __cplusplus
}
#endif
#endif
