// Fermion operator object of libcgpt_b200 (the thing a cgpt fermion-operator handle points to,
// lib/cgpt/lib/operators/base.h, implementation.h:47-110).
#pragma once
#include <math.h>
#include "common.cuh"

namespace cgptb {
struct LinkCoef {
  double w[4];    // -c_mu / 2
  double ph[8];   // boundary phases (re,im)
  int goff[4];    // global offset of the local lattice (multi-GPU decomposition)
  int gL[4];      // global extents
  int open_bc;    // open boundary conditions in time: U_t on the global slices 0, T-2, T-1 is zero (see import_gauge_t)
};
}  // namespace cgptb

struct cgptb_fermion_operator {
  int type = 0, prec = 0;
  int dims4[4] = {0, 0, 0, 0};
  int Ls = 0;
  cgptb::Geom g;
  cgptb_fermion_params p;
  void* links[2] = {0, 0};      // per output parity: [half4][8][9] complex, -c_mu/2 and phases folded in
  void* links_pad[2] = {0, 0};  // the same links as [2 halves][site][304 B] for the TMA sweep kernel (dslash_tma.cu), built lazily
  bool links_pad_valid = false;
  bool compress = false;        // two-row link compression (cgptb_fermion_params::link_compression == 12)
  void* links_c[2] = {0, 0};    // per output parity: [half4][8][rows 0 and 1 (12 reals), U(1) factor f (2 reals)]
  bool open_bc = false;         // boundary_phases[3] == 0: results vanish on the global time slices 0 and T-1
  bool has_clover = false;
  void* clov[2] = {0, 0};       // per parity: [72][half4] reals
  void* clov_inv[2] = {0, 0};
  void* s_coef = 0;             // Moebius tridiagonal tables [3 kinds][2 dag][5][Ls]
  void* s_inv = 0;              // dense MooeeInv blocks [P+, P-, P+^T, P-^T][Ls][Ls]
  bool zmobius = false;         // complex, s-dependent coefficients (p.n_omega > 0)
  void* z_tab = 0;              // zMoebius: dense complex blocks [kind 5][dag 2][P+, P-][Ls][Ls][re, im]
  cgptb_lattice* tmp_full[4] = {0, 0, 0, 0};
  cgptb_lattice* tmp_half[4] = {0, 0, 0, 0};
  cgptb_lattice* cg_half[5] = {0, 0, 0, 0, 0};  // work fields of cgptb_cg_eo2_ne (p, mmp, r, v, T p), kept between solves
  // multi-GPU decomposition (halo.cu); g.comm_mask marks the split directions
  int goff[4] = {0, 0, 0, 0};   // global coordinate of the local origin
  int gL[4] = {0, 0, 0, 0};     // global extents
  void* halo_send[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};  // [mu][lo,hi] projected faces of one parity
  void* halo_recv[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  void* ghost_links[4] = {0, 0, 0, 0};  // U_mu on the high face of the rank-mu neighbour (double)
  void* p2p = 0;                        // peer-to-peer halo state (halo.cu: HaloP2P) when the ranks can map each other's memory

  int ls() const { return Ls > 0 ? Ls : 1; }
  void check_field(const cgptb_lattice* l) const;
  cgptb_lattice* tmp(int i, int cb);
  void setup_mobius_tables();
  void import_gauge(const cgptb_lattice* const U[4]);
};

namespace cgptb {
void apply_open_boundaries(const cgptb_fermion_operator* op, cgptb_lattice* l);
void op_dhop(cgptb_fermion_operator* op, bool dag, const cgptb_lattice* in, cgptb_lattice* out);
void op_meooe(cgptb_fermion_operator* op, bool dag, const cgptb_lattice* in, cgptb_lattice* out);
void op_mooee(cgptb_fermion_operator* op, bool inverse, bool dag, bool acc, const cgptb_lattice* in, cgptb_lattice* out);
void op_s_tridiag(cgptb_fermion_operator* op, int kind, bool dag, bool acc, const cgptb_lattice* in, cgptb_lattice* out);
void op_s_dense(cgptb_fermion_operator* op, bool dag, const cgptb_lattice* in, cgptb_lattice* out);
// zMoebius: kind 0 = b_s + c_s S5, 1 = Mooee, 2 = MooeeInv, 3 = 1 - c_s (4 - M5), 4 = -c_s (the two pieces of Dminus)
enum { ZK_A = 0, ZK_EE = 1, ZK_EEINV = 2, ZK_DM1 = 3, ZK_DM2 = 4, ZK_COUNT = 5 };
void op_s_z(cgptb_fermion_operator* op, int zkind, bool dag, bool acc, const cgptb_lattice* in, cgptb_lattice* out);
void op_apply(cgptb_fermion_operator* op, int opcode, const cgptb_lattice* src, cgptb_lattice* dst);
// sweep.cu: fused fifth-dimension operators; T = (b + c S5)(bee - cee S5)^-1 = Meooe5D o MooeeInv
enum { SWEEP_T = 0, SWEEP_TDAG = 1, SWEEP_MINV = 2, SWEEP_MINVDAG = 3 };
bool op_s_sweep(cgptb_fermion_operator* op, int mode, const cgptb_lattice* in, cgptb_lattice* out);
// dslash_tma.cu
bool dhop_tma_usable(const cgptb_fermion_operator* op);
void dhop_tma_release(cgptb_fermion_operator* op);
void dhop_half_f32_tma(cgptb_fermion_operator* op, bool dag, const float* pin, size_t in_stride, float* pout, size_t out_stride,
                       int p_out, int t_begin = 0, int t_count = 0);
// halo.cu
void halo_setup(cgptb_fermion_operator* op, const cgptb_lattice* const U[4]);
void halo_begin(cgptb_fermion_operator* op, bool dag, int p_out, const void* in, size_t in_stride);
void halo_end(cgptb_fermion_operator* op, bool dag, int p_out, void* out, size_t out_stride);
void halo_release(cgptb_fermion_operator* op);
bool halo_is_p2p(const cgptb_fermion_operator* op);
}  // namespace cgptb
