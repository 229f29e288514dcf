/*
 * ORACLE (test infrastructure only) -- plain C / OpenMP restatement of the Wilson hopping term
 *
 *   Dhop psi(x) = -1/2 sum_mu c_mu [ (1 - g_mu) V_mu(x) psi(x+mu) + (1 + g_mu) V_mu^dag(x-mu) psi(x-mu) ]
 *
 * as stated by the reference in lib/gpt/qcd/fermion/reference/wilson_clover.py:182-200 (gamma basis
 * lib/gpt/core/gamma.py:28-41), applied to every s-slice of a 5d field like Grid's WilsonFermion5D::Dhop
 * (opcode 3001, lib/cgpt/lib/operators/register.h:17).  Used (a) by tests/ to cross-check the numpy oracle
 * on larger volumes and (b) by bench.py as the CPU baseline ("CPU restatement, Grid unavailable"): the
 * reference's own Grid build cannot be produced here.  Never linked into the product.
 *
 * Fields are in GPT order: site = s + Ls*(x + Lx*(y + Ly*(z + Lz*t))), 12 complex (spin*3+colour) per site,
 * interleaved re/im.  Links: V[mu][x4][3][3] complex with the boundary phases already applied.
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC (oracle/Makefile).
 */
#include <stddef.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* (1 + sgn*gamma_mu) psi: upper half-spinor h_k = psi_k + ph[k] * psi_{j[k]}, k = 0,1;
   lower components reconstructed as psi_2 = r2 * h_{k2}, psi_3 = r3 * h_{k3}.
   Phases are powers of i: 0:+1 1:+i 2:-1 3:-i.  Rows for sgn = -1; sgn = +1 adds 2 to every phase. */
static const int PJ[4][2] = {{3, 2}, {3, 2}, {2, 3}, {2, 3}};
static const int PP[4][2] = {{3, 3}, {0, 2}, {3, 1}, {2, 2}};
static const int RK[4][2] = {{1, 0}, {1, 0}, {0, 1}, {0, 1}};
static const int RP[4][2] = {{1, 1}, {2, 0}, {1, 3}, {2, 2}};

#define DEFINE_DHOP(NAME, REAL)                                                                                   \
  static inline void NAME##_mulph(int ph, REAL re, REAL im, REAL* ore, REAL* oim) {                               \
    switch (ph & 3) {                                                                                             \
      case 0: *ore = re; *oim = im; break;                                                                        \
      case 1: *ore = -im; *oim = re; break;                                                                       \
      case 2: *ore = -re; *oim = -im; break;                                                                      \
      default: *ore = im; *oim = -re; break;                                                                      \
    }                                                                                                             \
  }                                                                                                               \
  /* acc += recon( W(^dag) proj psi ) * w */                                                                      \
  static inline void NAME##_hop(REAL* acc, const REAL* psi, const REAL* W, int adj, int mu, int sgn, REAL w) {    \
    REAL h[12], chi[12];                                                                                          \
    int shift = sgn > 0 ? 2 : 0;                                                                                  \
    for (int k = 0; k < 2; k++)                                                                                   \
      for (int c = 0; c < 3; c++) {                                                                               \
        REAL ar, ai;                                                                                              \
        NAME##_mulph(PP[mu][k] + shift, psi[(PJ[mu][k] * 3 + c) * 2], psi[(PJ[mu][k] * 3 + c) * 2 + 1], &ar, &ai); \
        h[(k * 3 + c) * 2] = psi[(k * 3 + c) * 2] + ar;                                                           \
        h[(k * 3 + c) * 2 + 1] = psi[(k * 3 + c) * 2 + 1] + ai;                                                   \
      }                                                                                                           \
    for (int k = 0; k < 2; k++)                                                                                   \
      for (int r = 0; r < 3; r++) {                                                                               \
        REAL sr = 0, si = 0;                                                                                      \
        for (int c = 0; c < 3; c++) {                                                                             \
          REAL wr, wi;                                                                                            \
          if (!adj) { wr = W[(r * 3 + c) * 2]; wi = W[(r * 3 + c) * 2 + 1]; }                                     \
          else { wr = W[(c * 3 + r) * 2]; wi = -W[(c * 3 + r) * 2 + 1]; }                                         \
          REAL hr = h[(k * 3 + c) * 2], hi = h[(k * 3 + c) * 2 + 1];                                              \
          sr += wr * hr - wi * hi;                                                                                \
          si += wr * hi + wi * hr;                                                                                \
        }                                                                                                         \
        chi[(k * 3 + r) * 2] = sr * w;                                                                            \
        chi[(k * 3 + r) * 2 + 1] = si * w;                                                                        \
      }                                                                                                           \
    for (int c = 0; c < 3; c++) {                                                                                 \
      REAL r, i;                                                                                                  \
      acc[(0 * 3 + c) * 2] += chi[(0 * 3 + c) * 2];                                                               \
      acc[(0 * 3 + c) * 2 + 1] += chi[(0 * 3 + c) * 2 + 1];                                                       \
      acc[(1 * 3 + c) * 2] += chi[(1 * 3 + c) * 2];                                                               \
      acc[(1 * 3 + c) * 2 + 1] += chi[(1 * 3 + c) * 2 + 1];                                                       \
      NAME##_mulph(RP[mu][0] + shift, chi[(RK[mu][0] * 3 + c) * 2], chi[(RK[mu][0] * 3 + c) * 2 + 1], &r, &i);     \
      acc[(2 * 3 + c) * 2] += r;                                                                                  \
      acc[(2 * 3 + c) * 2 + 1] += i;                                                                              \
      NAME##_mulph(RP[mu][1] + shift, chi[(RK[mu][1] * 3 + c) * 2], chi[(RK[mu][1] * 3 + c) * 2 + 1], &r, &i);     \
      acc[(3 * 3 + c) * 2] += r;                                                                                  \
      acc[(3 * 3 + c) * 2 + 1] += i;                                                                              \
    }                                                                                                             \
  }                                                                                                               \
  void NAME(const int* dims, int Ls, const REAL* V, const double* coef, const REAL* in, REAL* out, int dag) {     \
    const long Lx = dims[0], Ly = dims[1], Lz = dims[2], Lt = dims[3];                                            \
    const long V4 = Lx * Ly * Lz * Lt;                                                                            \
    const long ls = Ls > 0 ? Ls : 1;                                                                              \
    _Pragma("omp parallel for schedule(static)") for (long x4 = 0; x4 < V4; x4++) {                               \
      long c[4] = {x4 % Lx, (x4 / Lx) % Ly, (x4 / (Lx * Ly)) % Lz, x4 / (Lx * Ly * Lz)};                          \
      const long L[4] = {Lx, Ly, Lz, Lt};                                                                         \
      const long stride[4] = {1, Lx, Lx * Ly, Lx * Ly * Lz};                                                      \
      for (long s = 0; s < ls; s++) {                                                                             \
        REAL acc[24];                                                                                             \
        memset(acc, 0, sizeof(acc));                                                                              \
        for (int mu = 0; mu < 4; mu++) {                                                                          \
          long xp = x4 + (c[mu] + 1 == L[mu] ? -(L[mu] - 1) : 1) * stride[mu];                                    \
          long xm = x4 + (c[mu] == 0 ? (L[mu] - 1) : -1) * stride[mu];                                            \
          REAL w = (REAL)(-0.5 * coef[mu]);                                                                       \
          /* forward: (1 - g) V(x) psi(x+mu); backward: (1 + g) V^dag(x-mu) psi(x-mu); dag swaps the signs */      \
          NAME##_hop(acc, in + (xp * ls + s) * 24, V + ((size_t)mu * V4 + x4) * 18, 0, mu, dag ? +1 : -1, w);     \
          NAME##_hop(acc, in + (xm * ls + s) * 24, V + ((size_t)mu * V4 + xm) * 18, 1, mu, dag ? -1 : +1, w);     \
        }                                                                                                         \
        memcpy(out + (x4 * ls + s) * 24, acc, sizeof(acc));                                                       \
      }                                                                                                           \
    }                                                                                                             \
  }

DEFINE_DHOP(oracle_dhop_f, float)
DEFINE_DHOP(oracle_dhop_d, double)

/* The same hopping term on a LIST of 4d sites only (all s): out[(i*ls + s)*24 ...] for sites[i].  bench.py uses it to
   check the GPU result on >= 1e5 sampled sites of the full-size lattice (and of every rank's block of a split
   lattice, padded with the neighbours' boundary slices) without paying for a full CPU application.             */
#define DEFINE_DHOP_SITES(NAME, HOP, REAL)                                                                        \
  void NAME(const int* dims, int Ls, const REAL* V, const double* coef, const REAL* in, long n, const long* sites, \
            REAL* out, int dag) {                                                                                 \
    const long Lx = dims[0], Ly = dims[1], Lz = dims[2], Lt = dims[3];                                            \
    const long V4 = Lx * Ly * Lz * Lt;                                                                            \
    const long ls = Ls > 0 ? Ls : 1;                                                                              \
    _Pragma("omp parallel for schedule(static)") for (long i = 0; i < n; i++) {                                   \
      const long x4 = sites[i];                                                                                   \
      long c[4] = {x4 % Lx, (x4 / Lx) % Ly, (x4 / (Lx * Ly)) % Lz, x4 / (Lx * Ly * Lz)};                          \
      const long L[4] = {Lx, Ly, Lz, Lt};                                                                         \
      const long stride[4] = {1, Lx, Lx * Ly, Lx * Ly * Lz};                                                      \
      for (long s = 0; s < ls; s++) {                                                                             \
        REAL acc[24];                                                                                             \
        memset(acc, 0, sizeof(acc));                                                                              \
        for (int mu = 0; mu < 4; mu++) {                                                                          \
          long xp = x4 + (c[mu] + 1 == L[mu] ? -(L[mu] - 1) : 1) * stride[mu];                                    \
          long xm = x4 + (c[mu] == 0 ? (L[mu] - 1) : -1) * stride[mu];                                            \
          REAL w = (REAL)(-0.5 * coef[mu]);                                                                       \
          HOP(acc, in + (xp * ls + s) * 24, V + ((size_t)mu * V4 + x4) * 18, 0, mu, dag ? +1 : -1, w);            \
          HOP(acc, in + (xm * ls + s) * 24, V + ((size_t)mu * V4 + xm) * 18, 1, mu, dag ? -1 : +1, w);            \
        }                                                                                                         \
        memcpy(out + (i * ls + s) * 24, acc, sizeof(acc));                                                        \
      }                                                                                                           \
    }                                                                                                             \
  }

DEFINE_DHOP_SITES(oracle_dhop_sites_f, oracle_dhop_f_hop, float)
DEFINE_DHOP_SITES(oracle_dhop_sites_d, oracle_dhop_d_hop, double)

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
