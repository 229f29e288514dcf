// GPT's parallel random number generator, rebuilt for libcgpt_b200: cgptb_create_random / cgptb_random_sample* replace
// cgpt.create_random / cgpt.random_sample / cgpt.delete_random (lib/cgpt/lib/random.cc:38-101), so that gpt_b200 draws the
// SAME gauge fields and sources as the reference from the same seed string -- the precondition of every parity statement
// ("identical random gauge fields and sources").
//
// Like the reference's, this is host code (the generator is a bit-serial integer recurrence with 6 KB of state per
// lattice block; the reference runs it on the CPU threads as well): OpenMP over the blocks, values produced in double in
// GPT order, then handed to the device through cgptb_lattice_import.  What is reproduced, file:line in /root/reference:
//   RANLUX24 (r = 24, s = 10, luxury p = 389 or 24)                      lib/cgpt/lib/random/ranlux.h:27-104
//   SHA-256 seeding, 64 virtual lanes advancing in lock-step              lib/cgpt/lib/random/vector.h:19-107
//   bit reservoir, uniform / Box-Muller normal / cnormal / zn             lib/cgpt/lib/random/distribution.h:19-175
//   one generator per 2^4 block of sites, block seed, fill order          lib/cgpt/lib/random/parallel.h:21-318
//   generators persist per (engine, grid)                                 lib/cgpt/lib/random/engine.h:37-131
//   gauge.random: su(3) generators, A = scale sum_a u_a T_a, exp(iA)      lib/gpt/qcd/gauge/create.py:66-71,
//       lib/gpt/core/object_type/su_n.py:208-240, lib/gpt/core/foundation/lattice/matrix/exp.py:167-219
#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <complex>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "common.cuh"

namespace cgptb {
namespace rng {

// ---- SHA-256 (FIPS 180-4) ---------------------------------------------------------------------------------
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void sha256(const uint8_t* data, size_t len, uint32_t h[8]) {
  static const uint32_t H0[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  memcpy(h, H0, sizeof(H0));
  std::vector<uint8_t> m(data, data + len);
  m.push_back(0x80);
  while (m.size() % 64 != 56) m.push_back(0);
  uint64_t bits = (uint64_t)len * 8;
  for (int i = 7; i >= 0; i--) m.push_back((uint8_t)(bits >> (8 * i)));
  for (size_t off = 0; off < m.size(); off += 64) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
      w[i] = ((uint32_t)m[off + 4 * i] << 24) | ((uint32_t)m[off + 4 * i + 1] << 16) | ((uint32_t)m[off + 4 * i + 2] << 8) | m[off + 4 * i + 3];
    for (int i = 16; i < 64; i++) {
      uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
      uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
      uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
      uint32_t ch = (e & f) ^ (~e & g);
      uint32_t t1 = hh + S1 + ch + K256[i] + w[i];
      uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
      uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
      uint32_t t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
}

// ---- RANLUX24, NL lanes in lock-step (NL = 1: the scalar generator that seeds the lanes) --------------------
static const uint32_t RL_B = 1u << 24, RL_MASK = RL_B - 1;
enum { RL_R = 24, RL_S = 10 };

template <int NL>
struct Ranlux {
  uint32_t x[RL_R][NL];
  uint32_t c[NL];
  int offset, discard, p;

  // seed[j][lane]: x = seed & (b-1), carry = (seed[0] == 0)       (ranlux.h:59-65)
  void seed(const uint32_t (*s)[NL], int luxury) {
    for (int j = 0; j < RL_R; j++)
      for (int l = 0; l < NL; l++) x[j][l] = s[j][l] & RL_MASK;
    for (int l = 0; l < NL; l++) c[l] = s[0][l] == 0 ? 1 : 0;
    offset = 0;
    discard = RL_R - 1;
    p = luxury;
  }
  inline void step(uint32_t* out) {
    offset = (offset + 1) % RL_R;
    const uint32_t* xs = x[(RL_S - 1 - offset + RL_R) % RL_R];
    const uint32_t* xr = x[(RL_R - 1 - offset + RL_R) % RL_R];
    uint32_t* x0 = x[(2 * RL_R - 1 - offset) % RL_R];
    for (int l = 0; l < NL; l++) {
      uint32_t d = xs[l] - xr[l] - c[l];  // wraps below zero exactly when a borrow is needed
      c[l] = d >= RL_B ? 1 : 0;
      d &= RL_MASK;
      x0[l] = d;
    }
    if (out)
      for (int l = 0; l < NL; l++) out[l] = x0[l];
  }
  inline void next(uint32_t* out) {
    if (++discard == RL_R) {
      discard = 0;
      for (int i = 0; i < p - RL_R; i++) step(0);
    }
    step(out);
  }
};

// ---- cgpt_random<vectorized ranlux, uint64_t>: bit reservoir + distributions ---------------------------------
struct Generator {
  Ranlux<64> v;
  uint32_t buffer[64];
  int nbuffer;
  uint64_t state;
  int nbits;
  bool has_stack;
  double stack;

  Generator(const std::vector<uint64_t>& seed, int luxury) {
    // three SHA-256 digests of (seed ++ [idx]) as little-endian uint64 -> 24 big-endian words (vector.h:38-58)
    uint32_t words[RL_R][1];
    std::vector<uint64_t> s(seed);
    s.push_back(0);
    for (int idx = 0; idx < 3; idx++) {
      s.back() = idx;
      uint32_t h[8];
      sha256(reinterpret_cast<const uint8_t*>(s.data()), s.size() * 8, h);
      for (int k = 0; k < 8; k++) words[8 * idx + k][0] = h[k];
    }
    Ranlux<1> sr;
    sr.seed(words, luxury);
    // lane L is seeded with the scalar outputs [24 L, 24 L + 24)   (vector.h:62-72)
    static thread_local uint32_t vs[RL_R][64];
    for (int lane = 0; lane < 64; lane++)
      for (int j = 0; j < RL_R; j++) sr.next(&vs[j][lane]);
    v.seed(vs, luxury);
    populate_buffer();
    state = 0;
    nbits = 0;
    has_stack = false;
    stack = 0.0;
    populate();
  }
  void populate_buffer() {
    v.next(buffer);
    nbuffer = 0;
  }
  inline uint32_t word() {
    if (nbuffer == 64) populate_buffer();
    return buffer[nbuffer++];
  }
  inline void populate() {
    state = (state << 24) + word();
    nbits += 24;
    if (nbits > 64) nbits = 64;
  }
  inline uint64_t get_bits(int bits) {
    while (bits > nbits) populate();
    const uint64_t base = (uint64_t)1 << bits;
    const uint64_t res = state & (base - 1);
    state >>= bits;
    nbits -= bits;
    return res;
  }
  inline double get_double() { return (double)get_bits(53) / (double)((uint64_t)1 << 53); }
  uint64_t get_uniform_int(uint64_t mx) {
    if (mx == 0) return 0;
    int bits = 0;
    for (uint64_t n = mx; n > 1; n >>= 1) bits++;
    bits += 1;
    for (;;) {
      uint64_t r = get_bits(bits);
      if (r <= mx) return r;
    }
  }
  double get_normal() {
    if (has_stack) {
      has_stack = false;
      return stack;
    }
    const double eps = std::numeric_limits<double>::min(), two_pi = 2.0 * 3.14159265358979323846;
    double u1, u2;
    do {
      u1 = get_double();
      u2 = get_double();
    } while (u1 <= eps);
    const double rad = ::sqrt(-2.0 * ::log(u1));
    stack = rad * ::sin(two_pi * u2);
    has_stack = true;
    return rad * ::cos(two_pi * u2);
  }
  // one complex sample of a distribution (distribution.h:117-175)
  inline void sample(int dist, double p0, double p1, double& re, double& im) {
    switch (dist) {
      case CGPTB_DIST_NORMAL:
        re = get_normal() * p1 + p0;
        im = 0.0;
        break;
      case CGPTB_DIST_CNORMAL:
        im = get_normal() * p1 + p0;  // imaginary part first
        re = get_normal() * p1 + p0;
        break;
      case CGPTB_DIST_UNIFORM_REAL:
        re = get_double() * (p1 - p0) + p0;
        im = 0.0;
        break;
      case CGPTB_DIST_UNIFORM_INT:
        re = (double)((long)get_uniform_int((uint64_t)((long)p1 - (long)p0)) + (long)p0);
        im = 0.0;
        break;
      default: {  // zn, n = p0
        const double n = p0;
        const double k = (double)get_uniform_int((uint64_t)(n - 1));
        std::complex<double> z = std::exp(std::complex<double>(k, 0.0) * std::complex<double>(0.0, 2.0 * M_PI / n));
        re = z.real();
        im = z.imag();
      }
    }
  }
};

struct GridRng {
  std::vector<std::unique_ptr<Generator>> gen;  // one per block of this rank
  int nd = 0;
  std::vector<int> ldims, block_dim, reduced_dim, block_size;  // block_size: 2 for blocked dimensions, 1 otherwise
  long nred = 0;
};

}  // namespace rng
}  // namespace cgptb

struct cgptb_random {
  std::string seed;
  int luxury;
  std::unique_ptr<cgptb::rng::Generator> scalar;
  std::map<uint64_t, cgptb::rng::GridRng> grids;
};

namespace cgptb {
namespace rng {

static std::vector<uint64_t> str_seed(const std::string& s) {
  std::vector<uint64_t> r;
  for (unsigned char ch : s) r.push_back(ch);
  return r;
}

// generators of a grid: blocks of 2 in every 4d direction, the fifth dimension (dimension 0 of a 5d grid) unblocked
static GridRng& grid_rng(cgptb_random* r, uint64_t key, int nd, const int* ldims, const int* gdims, const int* lstart) {
  auto it = r->grids.find(key);
  if (it != r->grids.end()) {
    GridRng& G = it->second;
    bool same = G.nd == nd;
    for (int j = 0; same && j < nd; j++) same = G.ldims[j] == ldims[j];
    if (!same) CGPTB_ERR("random: grid key %llu was used with another geometry before", (unsigned long long)key);
    return G;
  }
  if (nd < 1 || nd > 5) CGPTB_ERR("random: Nd = %d is not supported", nd);
  GridRng& G = r->grids[key];
  G.nd = nd;
  G.ldims.assign(ldims, ldims + nd);
  std::vector<bool> blocked(nd, true);
  if (nd == 5) blocked[0] = false;
  long blocks = 1;
  G.nred = 1;
  for (int j = 0; j < nd; j++) {
    if (blocked[j]) {
      if (gdims[j] % 2 || ldims[j] % 2 || lstart[j] % 2) CGPTB_ERR("random: extent %d of dimension %d is not a multiple of the block size 2", ldims[j], j);
      G.block_dim.push_back(ldims[j] / 2);
      G.reduced_dim.push_back(2);
      G.block_size.push_back(2);
    } else {
      G.block_dim.push_back(1);
      G.reduced_dim.push_back(ldims[j]);
      G.block_size.push_back(1);
    }
    blocks *= G.block_dim[j];
    G.nred *= G.reduced_dim[j];
  }
  // block seed = characters of the seed string ++ fdimensions ++ gdimensions ++ [global block index, dimension 0 most
  // significant]   (engine.h:88-98, parallel.h:74-89)
  std::vector<uint64_t> base = str_seed(r->seed);
  for (int j = 0; j < nd; j++) base.push_back((uint64_t)gdims[j]);
  for (int j = 0; j < nd; j++) base.push_back((uint64_t)gdims[j]);
  G.gen.resize(blocks);
  const int luxury = r->luxury;
#pragma omp parallel for schedule(dynamic, 16)
  for (long idx = 0; idx < blocks; idx++) {
    long rem = idx;
    uint64_t t = 0;
    std::vector<int> bc(nd);
    for (int j = 0; j < nd; j++) {
      bc[j] = (int)(rem % G.block_dim[j]);
      rem /= G.block_dim[j];
    }
    for (int j = 0; j < nd; j++)
      if (blocked[j]) t = t * (uint64_t)(gdims[j] / 2) + (uint64_t)(bc[j] + lstart[j] / 2);
    std::vector<uint64_t> s(base);
    s.push_back(t);
    G.gen[idx].reset(new Generator(s, luxury));
  }
  return G;
}

// out[site][e] (re, im), site = lexicographic local index with dimension 0 fastest; every block generator fills the
// sites of its block in lexicographic order (dimension 0 fastest), the nel elements of a site one after the other
template <typename TO>
static void sample_grid(GridRng& G, int nel, int dist, double p0, double p1, TO* out) {
  const int nd = G.nd;
  const long blocks = (long)G.gen.size();
#pragma omp parallel for schedule(dynamic, 16)
  for (long idx = 0; idx < blocks; idx++) {
    Generator& gen = *G.gen[idx];
    int bc[5], rc[5];
    long rem = idx;
    for (int j = 0; j < nd; j++) {
      bc[j] = (int)(rem % G.block_dim[j]);
      rem /= G.block_dim[j];
    }
    for (long ridx = 0; ridx < G.nred; ridx++) {
      long rr = ridx;
      for (int j = 0; j < nd; j++) {
        rc[j] = (int)(rr % G.reduced_dim[j]);
        rr /= G.reduced_dim[j];
      }
      long flat = 0, stride = 1;
      for (int j = 0; j < nd; j++) {
        const int c = bc[j] * G.block_size[j] + rc[j];
        flat += (long)c * stride;
        stride *= G.ldims[j];
      }
      TO* o = out + (size_t)flat * nel * 2;
      for (int e = 0; e < nel; e++) {
        double re, im;
        gen.sample(dist, p0, p1, re, im);  // always drawn in double, then cast (engine.h:104-105)
        o[2 * e] = (TO)re;
        o[2 * e + 1] = (TO)im;
      }
    }
  }
}

static void lattice_geometry(const cgptb_lattice* l, int& nd, int ldims[5], int gdims[5], int lstart[5]) {
  // A checkerboarded lattice is sampled like the reference samples a GridRedBlackCartesian (lib/cgpt/lib/random/parallel.h:
  // 27-128): the generators' blocks and the fill order follow the REDUCED coordinates (x/2, y, z, t) -- _ldimensions and
  // _gdimensions of such a grid have the checkerboarded extent halved --, whatever parity the lattice is labelled with; the
  // sample of reduced site (x/2, y, z, t) lands on the stored site with that x/2, which is this library's half-lattice order.
  const bool half = l->cb != CGPTB_FULL;
  if (half && (l->dims4[0] / 2) % 2) CGPTB_ERR("random: a checkerboarded lattice needs an x extent that is a multiple of 4 (2^4 blocks of the reduced lattice)");
  nd = 0;
  if (l->Ls > 0) {
    ldims[0] = gdims[0] = l->Ls;
    lstart[0] = 0;
    nd = 1;
  }
  for (int mu = 0; mu < 4; mu++, nd++) {
    const int d = half && mu == 0 ? l->dims4[mu] / 2 : l->dims4[mu];
    ldims[nd] = d;
    gdims[nd] = d * g_comm.pgrid[mu];
    lstart[nd] = d * g_comm.pcoor[mu];
  }
}

static void import_doubles(cgptb_lattice* l, const std::vector<double>& v) {
  if (l->prec == CGPTB_DOUBLE) {
    if (cgptb_lattice_import(l, v.data(), v.size() * sizeof(double))) throw Error{g_error};
  } else {
    std::vector<float> f(v.size());
    for (size_t i = 0; i < v.size(); i++) f[i] = (float)v[i];  // generated in double, then cast (engine.h:104-105)
    if (cgptb_lattice_import(l, f.data(), f.size() * sizeof(float))) throw Error{g_error};
  }
  CUDA_CHECK(cudaStreamSynchronize(g_stream));  // the staging copy reads the host vector
}

// su(3) generators in GPT's order and normalisation tr T_a T_a = 1/2 (su_n.py:208-240)
static void su3_generators(std::complex<double> T[8][3][3]) {
  int n = 0;
  for (int i = 0; i < 3; i++)
    for (int j = i + 1; j < 3; j++) {
      std::complex<double> a[3][3] = {};
      a[i][j] = 1.0;
      a[j][i] = 1.0;
      memcpy(T[n++], a, sizeof(a));
      std::complex<double> b[3][3] = {};
      b[i][j] = std::complex<double>(0, -1);
      b[j][i] = std::complex<double>(0, 1);
      memcpy(T[n++], b, sizeof(b));
      if (j == i + 1) {
        std::complex<double> d[3][3] = {};
        for (int l = 0; l < j; l++) d[l][l] = 1.0;
        d[j][j] = -(double)j;
        memcpy(T[n++], d, sizeof(d));
      }
    }
  for (int a = 0; a < 8; a++) {
    std::complex<double> tr = 0;
    for (int i = 0; i < 3; i++)
      for (int k = 0; k < 3; k++) tr += T[a][i][k] * T[a][k][i];
    const double nrm = ::sqrt(tr.real() * 2.0);
    for (int i = 0; i < 3; i++)
      for (int k = 0; k < 3; k++) T[a][i][k] /= nrm;
  }
}

typedef std::complex<double> cd;
static inline void mm3(const cd* a, const cd* b, cd* c) {
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 3; k++) c[3 * i + k] = a[3 * i] * b[k] + a[3 * i + 1] * b[3 + k] + a[3 * i + 2] * b[6 + k];
}

}  // namespace rng
}  // namespace cgptb

using namespace cgptb;
using namespace cgptb::rng;

extern "C" {

int cgptb_create_random(cgptb_random** out, const char* engine, const char* seed) {
  CGPTB_API_BEGIN
  int luxury = 0;
  if (!strcmp(engine, "vectorized_ranlux24_389_64"))
    luxury = 389;
  else if (!strcmp(engine, "vectorized_ranlux24_24_64"))
    luxury = 24;
  else
    CGPTB_ERR("Unknown rng engine type: %s", engine);
  // torchrun exports OMP_NUM_THREADS=1 to every rank; the generators are host code, so give each rank its share of the cores
  // (one process per GPU: cores / LOCAL_WORLD_SIZE) unless the user asked for more than one thread explicitly
  {
    const char* lws = getenv("LOCAL_WORLD_SIZE");
    const char* ont = getenv("OMP_NUM_THREADS");
    if (lws && atoi(lws) >= 1 && (!ont || atoi(ont) <= 1)) {
      int share = omp_get_num_procs() / atoi(lws);
      if (share > omp_get_max_threads()) omp_set_num_threads(share);
    }
  }
  cgptb_random* r = new cgptb_random;
  r->seed = seed;
  r->luxury = luxury;
  r->scalar.reset(new Generator(str_seed(r->seed), luxury));
  *out = r;
  CGPTB_API_END
}

int cgptb_delete_random(cgptb_random* r) {
  delete r;
  return 0;
}

int cgptb_random_sample_scalar(cgptb_random* r, int dist, double p0, double p1, double out[2]) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(r && dist >= 0 && dist <= CGPTB_DIST_ZN);
  r->scalar->sample(dist, p0, p1, out[0], out[1]);
  CGPTB_API_END
}

int cgptb_random_sample_host(cgptb_random* r, uint64_t grid_key, int nd, const int* ldims, const int* gdims, const int* lstart,
                             int nel, int dist, double p0, double p1, double* out) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(r && nel > 0 && dist >= 0 && dist <= CGPTB_DIST_ZN);
  GridRng& G = grid_rng(r, grid_key, nd, ldims, gdims, lstart);
  sample_grid(G, nel, dist, p0, p1, out);
  CGPTB_API_END
}

int cgptb_random_sample(cgptb_random* r, uint64_t grid_key, cgptb_lattice* l, int dist, double p0, double p1) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(r && l && dist >= 0 && dist <= CGPTB_DIST_ZN);
  int nd, ldims[5], gdims[5], lstart[5];
  lattice_geometry(l, nd, ldims, gdims, lstart);
  GridRng& G = grid_rng(r, grid_key, nd, ldims, gdims, lstart);
  const size_t n = l->sites * (size_t)l->otype * 2;
  if (l->prec == CGPTB_DOUBLE) {
    std::vector<double> v(n);
    sample_grid(G, l->otype, dist, p0, p1, v.data());
    import_doubles(l, v);
  } else {
    std::vector<float> v(n);
    sample_grid(G, l->otype, dist, p0, p1, v.data());
    if (cgptb_lattice_import(l, v.data(), n * sizeof(float))) throw Error{g_error};
    CUDA_CHECK(cudaStreamSynchronize(g_stream));  // the staging copy reads the host vector
  }
  CGPTB_API_END
}

int cgptb_random_su3_links(cgptb_random* r, uint64_t grid_key, cgptb_lattice* const U[4], double scale) {
  CGPTB_API_BEGIN
  CGPTB_ASSERT(r && U);
  for (int mu = 0; mu < 4; mu++) CGPTB_ASSERT(U[mu] && U[mu]->otype == 9 && U[mu]->Ls == 0 && same_shape(U[mu], U[0]));
  int nd, ldims[5], gdims[5], lstart[5];
  lattice_geometry(U[0], nd, ldims, gdims, lstart);
  GridRng& G = grid_rng(r, grid_key, nd, ldims, gdims, lstart);
  const size_t sites = U[0]->sites;
  const bool single = U[0]->prec == CGPTB_SINGLE;
  double gsites = (double)sites;
  for (int mu = 0; mu < 4; mu++) gsites *= g_comm.pgrid[mu];
  cd T[8][3][3];
  su3_generators(T);
  std::vector<double> ca(sites * 2), A(sites * 18), out(sites * 18);
  for (int mu = 0; mu < 4; mu++) {
    std::fill(A.begin(), A.end(), 0.0);
    for (int a = 0; a < 8; a++) {
      sample_grid(G, 1, CGPTB_DIST_UNIFORM_REAL, -0.5, 0.5, ca.data());
      // A += scale * ca * T_a in the precision of the lattice (lib/gpt/core/random.py:136-140)
#pragma omp parallel for
      for (long i = 0; i < (long)sites; i++) {
        for (int k = 0; k < 9; k++) {
          const cd t = T[a][k / 3][k % 3];
          if (single) {
            const std::complex<float> c((float)ca[2 * i], 0.f), tf((float)t.real(), (float)t.imag());
            const std::complex<float> acc((float)A[18 * i + 2 * k], (float)A[18 * i + 2 * k + 1]);
            const std::complex<float> v = acc + (std::complex<float>((float)scale, 0.f) * c) * tf;
            A[18 * i + 2 * k] = v.real();
            A[18 * i + 2 * k + 1] = v.imag();
          } else {
            const cd v = cd(A[18 * i + 2 * k], A[18 * i + 2 * k + 1]) + (cd(scale, 0.0) * cd(ca[2 * i], 0.0)) * t;
            A[18 * i + 2 * k] = v.real();
            A[18 * i + 2 * k + 1] = v.imag();
          }
        }
      }
    }
    // U = exp(i A): scaling by the lattice-wide norm, Taylor series to order 19, repeated squaring (exp.py:174-213)
    double n2 = 0.0;
#pragma omp parallel for reduction(+ : n2)
    for (long i = 0; i < (long)(sites * 18); i++) n2 += A[i] * A[i];
    if (g_comm.active) {
      if (cgptb_comm_globalsum(&n2, 1)) throw Error{g_error};
    }
    const double n = ::sqrt(n2) / gsites, maxn = 0.01;
    int ns = 0;
    if (n > maxn) ns = (int)::log2(n / maxn);
    const double sc = ::ldexp(1.0, -ns);
#pragma omp parallel for
    for (long i = 0; i < (long)sites; i++) {
      cd x[9], o[9], xn[9], t[9];
      for (int k = 0; k < 9; k++) x[k] = cd(0.0, 1.0) * cd(A[18 * i + 2 * k], A[18 * i + 2 * k + 1]) * sc;
      for (int k = 0; k < 9; k++) {
        xn[k] = x[k];
        o[k] = (k % 4 == 0 ? cd(1.0) : cd(0.0)) + x[k];
      }
      double nfac = 1.0;
      for (int j = 2; j < 20; j++) {
        nfac /= j;
        mm3(xn, x, t);
        for (int k = 0; k < 9; k++) {
          xn[k] = t[k];
          o[k] += xn[k] * nfac;
        }
      }
      for (int j = 0; j < ns; j++) {
        mm3(o, o, t);
        for (int k = 0; k < 9; k++) o[k] = t[k];
      }
      for (int k = 0; k < 9; k++) {
        out[18 * i + 2 * k] = o[k].real();
        out[18 * i + 2 * k + 1] = o[k].imag();
      }
    }
    import_doubles(U[mu], out);
  }
  CGPTB_API_END
}
}
