// BP simple-update gate application on the device: the main CONSUMER of the BP messages.
//
// Restates /root/reference/src/apply/apply_operators.jl for a BATCH of vertex-disjoint gates (one Trotter layer):
//   * one-site gate  (:226-244)  A_v <- O A_v, optionally divided by the norm of the gauged tensor
//   * two-site gate  (:246-283)  gauges X_w from the incoming boundary messages (gram_eigh_full_with_pinv, :250-253)
//                                -> gauged tensors (:255-256) -> compact QR against the external legs (:258-259)
//                                -> gate on R_1 R_2 (:260) -> truncated SVD (:261) -> optional S / |S| (:262-264)
//                                -> sqrt(S) split (:265-268) -> Q R and inverse gauges (:270-271)
//                                -> new messages diag(S) on both directions of the gate edge (:273-277)
// The reference does this for ONE gate per call on the host (MatrixAlgebraKit qr_compact / svd_trunc,
// TensorAlgebra gram_eigh_full); here one CTA owns one gate and a launch covers a whole layer of disjoint gates.
//
// Numerical building blocks (all written against the `Team` abstraction below so that the SAME code runs as a CUDA
// block on the device and as a single sequential "lane" on the host -- tests/native/apply_host.cu compiles this
// header for the host and checks every stage against numpy without a GPU; the product path never runs it there):
//   * jacobi_cols     one-sided (Hestenes) Jacobi on the columns of an m x n matrix, round-robin pair schedule, one
//                     warp per column pair: B V = U S.  Used for the SVD of the gated bond matrix AND for the
//                     eigen-decomposition of the (Hermitian PSD) messages (G V = V D; the sign of every eigenvalue is
//                     recovered from the Rayleigh quotient so that slightly indefinite messages are treated like
//                     numpy.linalg.eigh would).
//   * householder_qr  LAPACK geqr2-style reflectors (complex: zlarfg convention), one warp per trailing column.
//   * apply_q         Q Y from the stored reflectors, one warp per column of Y (columns are independent).
//   * mode_product    T[.., g, ..] = sum_l X[g, l] T[.., l, ..] on one external leg of the matrix view.
// All matrices are column-major.  Work space lives in global memory (L2 resident per gate); this is the first,
// correctness-oriented version of the path: see DESIGN.md §4.9 for its traffic/flop budget and what a tuned version does.
#pragma once
#include "bpx_common.cuh"

namespace bpx {
namespace applyk {

constexpr int MAXZ = BPX_MAX_DEGREE;
constexpr double EPS = 2.220446049250313e-16;
constexpr int MAX_JACOBI_SWEEPS = 60;
constexpr int NT = 256;  // threads per CTA on the device

// ---- the team: a CTA on the device, one sequential lane on the host -----------------------------------------
#ifndef BPX_HOST_TEAM_SYNC
#define BPX_HOST_TEAM_SYNC()
#endif
#ifndef BPX_HOST_LANES  // the race-check harness can also run several lanes per warp as threads
#define BPX_HOST_LANES 1
#define BPX_HOST_WARP_SUM(team, x) (x)
#define BPX_HOST_SYNCWARP(team)
#define BPX_HOST_ATOMIC_ADD(p, v) (*(p) += (v))
#endif
#ifndef BPX_FLAG_SET
#define BPX_FLAG_SET(p) (*(p) = 1)  // the race-check harness makes this a relaxed atomic store for ThreadSanitizer
#endif
struct Team {
  int lane, wid, nw;  // lane in the warp, warp in the team, warps in the team
  __host__ __device__ __forceinline__ int lanes() const {
#ifdef __CUDA_ARCH__
    return 32;
#else
    return BPX_HOST_LANES;
#endif
  }
  __host__ __device__ __forceinline__ int tid() const { return wid * lanes() + lane; }
  __host__ __device__ __forceinline__ int nt() const { return nw * lanes(); }
  __host__ __device__ __forceinline__ void sync() const {
#ifdef __CUDA_ARCH__
    __syncthreads();
#else
    BPX_HOST_TEAM_SYNC();  // empty unless a host test harness runs several "warps" as threads (tests/native/apply_race_check.cu)
#endif
  }
  // warp-wide sums (every lane gets the result)
  __host__ __device__ __forceinline__ double sum(double x) const {
#ifdef __CUDA_ARCH__
    x = warp_sum_d(x);
#else
    x = BPX_HOST_WARP_SUM(*this, x);
#endif
    return x;
  }
  template <typename T>
  __host__ __device__ __forceinline__ T sum_t(T x) const {
#ifdef __CUDA_ARCH__
    x = warp_sum<T>(x);
#else
    x = BPX_HOST_WARP_SUM(*this, x);
#endif
    return x;
  }
};

// ---- scalar helpers -----------------------------------------------------------------------------------------
template <typename T>
__host__ __device__ __forceinline__ T from_real(double r);
template <>
__host__ __device__ __forceinline__ double from_real<double>(double r) { return r; }
template <>
__host__ __device__ __forceinline__ c64 from_real<c64>(double r) { return make_c64(r, 0.0); }
__host__ __device__ __forceinline__ double real_of(double a) { return a; }
__host__ __device__ __forceinline__ double real_of(c64 a) { return a.re; }
__host__ __device__ __forceinline__ double imag_of(double) { return 0.0; }
__host__ __device__ __forceinline__ double imag_of(c64 a) { return a.im; }
__host__ __device__ __forceinline__ double scal(double a, double r) { return a * r; }
__host__ __device__ __forceinline__ c64 scal(c64 a, double r) { return make_c64(a.re * r, a.im * r); }
template <typename T>
__host__ __device__ __forceinline__ T sub(T a, T b) { return Elem<T>::add(a, scal(b, -1.0)); }

// ---- descriptors (built on the host, one per gate) ----------------------------------------------------------
struct Side {
  int64_t site_off;       // element offset of A_v in the device site buffer
  int64_t n;              // elements of A_v = d * prod(dim)
  int64_t rows;           // matrix view: prod of the external link dims (all link dims for a one-site gate)
  int64_t in_msg[MAXZ];   // element offsets of the messages w_i -> v in the message set (slot order)
  int32_t dim[MAXZ];      // link dims, slot order
  int32_t z, d;
  int32_t bond_slot;      // leg towards the other vertex of the gate; -1 for a one-site gate
  int32_t cols;           // d * chi_bond  (d for a one-site gate)
  int32_t nref;           // min(rows, cols): rows of R
  int32_t pad_;
};

struct GateDesc {
  Side s[2];
  int64_t msg12, msg21;   // element offsets of the two messages on the gate edge
  int64_t op_off;         // element offset of the operator in the packed operator buffer
  int64_t ws_off;         // element offset of this gate's work space
  int32_t nsides;         // 1 or 2
  int32_t chi_b;          // bond dimension (stays the leg's dimension; the kept rank k is zero-padded up to it)
  int32_t k;              // kept rank: min(max_rank, chi_b, m, n)
  int32_t sv_row_p1;      // 1 + row of sv_out that receives this gate's singular values (0: the gate's index in the launch)
};

__host__ __device__ inline int64_t gauge_elems(const Side& s) {  // per external leg: B (-> X), V (-> Xinv), eigenvalues
  int64_t t = 0;
  for (int i = 0; i < s.z; ++i)
    if (i != s.bond_slot) t += 2 * (int64_t)s.dim[i] * s.dim[i] + s.dim[i];
  return t;
}

// work-space layout of one gate, in elements of T (identical on host and device)
struct Layout {
  int64_t buf[2][2];   // two ping-pong tensors per side
  int64_t tau[2];      // Householder scalars
  int64_t gauge[2];    // per-leg gauge matrices
  int64_t r[2];        // R factors, nref x cols
  int64_t theta[2];    // bond matrix before / after the gate, m x n
  int64_t vs;          // right singular vectors, n x n
  int64_t sig;         // singular values (as doubles; n T-slots reserved)
  int64_t order;       // sorted column order (as int32; n T-slots reserved)
  int64_t total;
};

__host__ __device__ inline Layout layout_of(const GateDesc& g) {
  Layout L;
  int64_t o = 0;
  for (int a = 0; a < 2; ++a) {
    const bool live = a < g.nsides;
    const Side& s = g.s[a];
    for (int b = 0; b < 2; ++b) { L.buf[a][b] = o; o += live ? s.n : 0; }
    L.tau[a] = o; o += live ? s.cols : 0;
    L.gauge[a] = o; o += live ? gauge_elems(s) : 0;
    L.r[a] = o; o += live ? (int64_t)s.nref * s.cols : 0;
  }
  int64_t m = 0, n = 0;
  if (g.nsides == 2) {
    m = (int64_t)g.s[0].nref * g.s[0].d;
    n = (int64_t)g.s[1].nref * g.s[1].d;
  }
  L.theta[0] = o; o += m * n;
  L.theta[1] = o; o += m * n;
  L.vs = o; o += n * n;
  L.sig = o; o += n;
  L.order = o; o += n;
  L.total = o;
  return L;
}

// Fill the derived fields of a side (host).  dims / in_msg / z / d / bond_slot / site_off must be set.
inline void finish_side(Side& s, int chi_b) {
  int64_t rows = 1, n = s.d;
  for (int i = 0; i < s.z; ++i) {
    n *= s.dim[i];
    if (i != s.bond_slot) rows *= s.dim[i];
  }
  s.n = n;
  s.rows = rows;
  s.cols = s.bond_slot >= 0 ? s.d * chi_b : s.d;
  s.nref = (int32_t)(rows < (int64_t)s.cols ? rows : (int64_t)s.cols);
}

// ---- one-sided Jacobi -----------------------------------------------------------------------------------------
// Orthogonalise the columns of B (m x n, leading dimension m) by plane rotations from the right, accumulated in V
// (n x n, initialised to the identity here): on return B_in V = B_out with mutually orthogonal columns.
// Round-robin schedule: n' = n rounded up to even players, n' - 1 steps of n'/2 disjoint pairs per sweep, one warp
// per pair (lanes stride the rows).  `flag` is one int visible to the whole team (shared memory on the device).
// STABLE_PHASE: use the rotation [[c, s ph], [-s conj(ph), c]], which tends to the identity as s -> 0; the default
// (version 1, verified on the B200) multiplies column q by conj(ph) -- equally valid, but near convergence ph is the phase
// of rounding noise, so the phases of the singular vectors (a bond gauge) depend on the order of the reductions.
template <typename T, bool STABLE_PHASE = false>
__host__ __device__ void jacobi_cols(const Team& tm, T* B, int m, int n, T* V, int* flag) {
  using E = Elem<T>;
  const int L = tm.lanes();
  for (int64_t i = tm.tid(); i < (int64_t)n * n; i += tm.nt()) V[i] = from_real<T>((i % n) == (i / n) ? 1.0 : 0.0);
  tm.sync();
  if (n < 2) return;
  const int np = (n + 1) & ~1;
  const double tol2 = (double)m * EPS * EPS;  // |g|^2 <= m eps^2 a b: converged pair (LAPACK xGESVJ's sqrt(m) eps)
  // columns whose norm falls below n eps |B|_F are numerically zero (the null space of a wide or rank-deficient
  // matrix): they are left alone -- rotating them only chases rounding noise down to the underflow range.
  // Every warp derives the threshold redundantly (identical arithmetic, no cross-warp reduction).
  double fro2 = 0.0;
  for (int64_t i = tm.lane; i < (int64_t)m * n; i += L) fro2 += E::abs2(B[i]);
  fro2 = tm.sum(fro2);
  const double zero2 = (double)n * n * EPS * EPS * fro2;
  tm.sync();
  for (int sweep = 0; sweep < MAX_JACOBI_SWEEPS; ++sweep) {
    if (tm.tid() == 0) *flag = 0;
    tm.sync();
    for (int step = 0; step < np - 1; ++step) {
      for (int idx = tm.wid; idx < np / 2; idx += tm.nw) {
        int p, q;
        if (idx == 0) {
          p = np - 1;
          q = step;
        } else {
          p = (step + idx) % (np - 1);
          q = (step - idx + (np - 1)) % (np - 1);
        }
        if (p > q) { const int t = p; p = q; q = t; }
        if (q >= n) continue;  // the bye of an odd n
        T* bp = B + (int64_t)p * m;
        T* bq = B + (int64_t)q * m;
        double a = 0.0, b = 0.0;
        T g = E::zero();
        for (int r = tm.lane; r < m; r += L) {
          const T x = bp[r], y = bq[r];
          a += E::abs2(x);
          b += E::abs2(y);
          g = E::fma(E::conj(x), y, g);
        }
        a = tm.sum(a);
        b = tm.sum(b);
        g = tm.template sum_t<T>(g);
        const double g2 = E::abs2(g);
        if (!(g2 > tol2 * a * b) || !(a > zero2) || !(b > zero2)) continue;  // warp-uniform
        const double ga = sqrt(g2);
        const T ph = scal(g, 1.0 / ga);                // g / |g|
        const double zeta = (b - a) / (2.0 * ga);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        const T cph = E::conj(ph);
        T* vp = V + (int64_t)p * n;
        T* vq = V + (int64_t)q * n;
        if constexpr (STABLE_PHASE) {
          // [p', q'] = [p, q] J,  J = [[c, s ph], [-s conj(ph), c]]  (unitary, -> 1 as s -> 0)
          const T sph = scal(ph, s), scph = scal(cph, s);
          for (int r = tm.lane; r < m; r += L) {
            const T x = bp[r], y = bq[r];
            bp[r] = sub(scal(x, c), E::mul(y, scph));
            bq[r] = E::add(E::mul(x, sph), scal(y, c));
          }
          for (int r = tm.lane; r < n; r += L) {
            const T x = vp[r], y = vq[r];
            vp[r] = sub(scal(x, c), E::mul(y, scph));
            vq[r] = E::add(E::mul(x, sph), scal(y, c));
          }
        } else {
          // [p', q'] = [p, q] J,  J = [[c, s], [-s conj(ph), c conj(ph)]]  (unitary)
          for (int r = tm.lane; r < m; r += L) {
            const T x = bp[r], y = E::mul(bq[r], cph);
            bp[r] = sub(scal(x, c), scal(y, s));
            bq[r] = E::add(scal(x, s), scal(y, c));
          }
          for (int r = tm.lane; r < n; r += L) {
            const T x = vp[r], y = E::mul(vq[r], cph);
            vp[r] = sub(scal(x, c), scal(y, s));
            vq[r] = E::add(scal(x, s), scal(y, c));
          }
        }
        if (tm.lane == 0) BPX_FLAG_SET(flag);  // several warps may store the same 1: benign
      }
      tm.sync();
    }
    const int f = *flag;
    tm.sync();
    if (!f) break;
  }
}

// ---- gauges from one message (gram_eigh_full_with_pinv, apply_operators.jl:250-253) ---------------------------
// msg[bra, ket] chi x chi.  On return X[g, l] = sqrt(d_g) conj(V[l, g]) (X^H X = G) in `bx`, Xinv[l, g] = V[l, g] /
// sqrt(d_g) in `vx` (0 for eigenvalues at or below chi eps d_max), eigenvalues in ev.
template <typename T>
__host__ __device__ void gauge_from_message(const Team& tm, const T* msg, int chi, T* bx, T* vx, double* ev, int* flag) {
  using E = Elem<T>;
  for (int i = tm.tid(); i < chi * chi; i += tm.nt()) {
    const int r = i % chi, c = i / chi;
    bx[i] = scal(E::add(msg[r + c * chi], E::conj(msg[c + r * chi])), 0.5);  // Hermitian part
  }
  tm.sync();
  jacobi_cols<T>(tm, bx, chi, chi, vx, flag);
  // eigenvalue j = Re(v_j^H G v_j), G v_j = column j of B
  for (int j = tm.tid(); j < chi; j += tm.nt()) {
    T acc = E::zero();
    for (int r = 0; r < chi; ++r) acc = E::fma(E::conj(vx[r + j * chi]), bx[r + j * chi], acc);
    ev[j] = real_of(acc);
  }
  tm.sync();
  double dmax = 0.0;
  for (int j = 0; j < chi; ++j) dmax = ev[j] > dmax ? ev[j] : dmax;
  const double cut = EPS * chi * dmax;
  for (int i = tm.tid(); i < chi * chi; i += tm.nt()) {
    const int g = i % chi, l = i / chi;  // X[g, l]
    const double d = ev[g];
    bx[i] = d > cut ? scal(E::conj(vx[l + g * chi]), sqrt(d)) : E::zero();
  }
  tm.sync();
  for (int i = tm.tid(); i < chi * chi; i += tm.nt()) {
    const int g = i / chi;  // Xinv[l, g]: scale column g of V
    const double d = ev[g];
    vx[i] = d > cut ? scal(vx[i], 1.0 / sqrt(d)) : E::zero();
  }
  tm.sync();
}

// ---- tensor <-> matrix view -----------------------------------------------------------------------------------
// P[row, col]: row = external link legs in slot order (first fastest), col = s + d * bond index (s for one-site).
// Decompose a canonical element index of A_v[s, l_0..l_{z-1}] into (row, col).
__host__ __device__ __forceinline__ void split_index(const Side& sd, int64_t i, int64_t& row, int& col) {
  int s = (int)(i % sd.d);
  int64_t rest = i / sd.d, stride = 1;
  row = 0;
  int bond = 0;
  for (int k = 0; k < sd.z; ++k) {
    const int l = (int)(rest % sd.dim[k]);
    rest /= sd.dim[k];
    if (k == sd.bond_slot) {
      bond = l;
    } else {
      row += l * stride;
      stride *= sd.dim[k];
    }
  }
  col = s + sd.d * bond;
}

template <typename T>
__host__ __device__ void tensor_to_matrix(const Team& tm, const Side& sd, const T* a, T* p) {
  for (int64_t i = tm.tid(); i < sd.n; i += tm.nt()) {
    int64_t row;
    int col;
    split_index(sd, i, row, col);
    p[row + sd.rows * col] = a[i];
  }
  tm.sync();
}

// out[.., g, ..] = sum_l x[g + chi l] in[.., l, ..] on the leg with row stride `st` and dimension chi
template <typename T>
__host__ __device__ void mode_product(const Team& tm, const T* in, T* out, int64_t rows, int ncols, int64_t st, int chi,
                                      const T* x) {
  using E = Elem<T>;
  const int64_t total = rows * ncols;
  for (int64_t i = tm.tid(); i < total; i += tm.nt()) {
    const int64_t row = i % rows, col = i / rows;
    const int64_t lo = row % st, g = (row / st) % chi, hi = row / (st * chi);
    const T* src = in + col * rows + hi * st * chi + lo;
    T acc = E::zero();
    for (int l = 0; l < chi; ++l) acc = E::fma(x[g + (int64_t)chi * l], src[l * st], acc);
    out[i] = acc;
  }
  tm.sync();
}

// ---- Householder QR (geqr2 / larfg conventions) -----------------------------------------------------------------
// P is rows x cols (ld rows).  On return the upper trapezoid holds R, the columns below the diagonal the reflector
// vectors (v_j = 1 implicit), tau[j] the scalars: P_in = H_0 ... H_{nr-1} R, H_j = I - tau_j v_j v_j^H.
// Returns the number of reflectors nr = min(cols, rows - 1) (tau of trivial reflectors is 0).
template <typename T>
__host__ __device__ int householder_qr(const Team& tm, T* P, int64_t rows, int cols, T* tau) {
  using E = Elem<T>;
  const int L = tm.lanes();
  const int nr = (int)((rows - 1) < (int64_t)cols ? (rows - 1) : (int64_t)cols);
  for (int j = 0; j < nr; ++j) {
    T* cj = P + (int64_t)j * rows;
    // every warp derives the reflector redundantly (identical arithmetic), no cross-warp reduction needed
    double xn2 = 0.0;
    for (int64_t r = j + 1 + tm.lane; r < rows; r += L) xn2 += E::abs2(cj[r]);
    xn2 = tm.sum(xn2);
    const T alpha = cj[j];
    const double ar = real_of(alpha), ai = imag_of(alpha);
    T tj = E::zero(), scale = E::zero();
    double beta = ar;
    const bool trivial = (xn2 == 0.0 && ai == 0.0);
    if (!trivial) {
      beta = sqrt(ar * ar + ai * ai + xn2);
      if (ar >= 0.0) beta = -beta;
      tj = E::add(from_real<T>((beta - ar) / beta), scal(sub(alpha, from_real<T>(ar)), -1.0 / beta));  // ((beta-ar)/beta, -ai/beta)
      scale = E::div(from_real<T>(1.0), sub(alpha, from_real<T>(beta)));
      const T ctau = E::conj(tj);
      for (int c = j + 1 + tm.wid; c < cols; c += tm.nw) {
        T* cc = P + (int64_t)c * rows;
        T w = E::zero();
        for (int64_t r = j + 1 + tm.lane; r < rows; r += L) w = E::fma(E::conj(E::mul(cj[r], scale)), cc[r], w);
        w = tm.template sum_t<T>(w);
        w = E::add(w, cc[j]);
        const T f = E::mul(ctau, w);
#ifdef __CUDA_ARCH__
        __syncwarp();  // every lane has read cc[j] before lane 0 overwrites it
#else
        BPX_HOST_SYNCWARP(tm);
#endif
        for (int64_t r = j + 1 + tm.lane; r < rows; r += L) cc[r] = sub(cc[r], E::mul(f, E::mul(cj[r], scale)));
        if (tm.lane == 0) cc[j] = sub(cc[j], f);
      }
    }
    tm.sync();
    if (!trivial) {
      for (int64_t r = j + 1 + tm.tid(); r < rows; r += tm.nt()) cj[r] = E::mul(cj[r], scale);
      if (tm.tid() == 0) cj[j] = from_real<T>(beta);
    }
    if (tm.tid() == 0) tau[j] = tj;
    tm.sync();
  }
  return nr;
}

// Y <- H_0 ... H_{nr-1} Y  (= Q Y); Y is rows x ncols.  Columns are independent: one warp per column, no syncs inside.
template <typename T>
__host__ __device__ void apply_q(const Team& tm, const T* P, int64_t rows, int nr, const T* tau, T* Y, int ncols) {
  using E = Elem<T>;
  const int L = tm.lanes();
  for (int c = tm.wid; c < ncols; c += tm.nw) {
    T* y = Y + (int64_t)c * rows;
    for (int j = nr - 1; j >= 0; --j) {
      const T tj = tau[j];
      if (E::is_zero(tj)) continue;
      const T* v = P + (int64_t)j * rows;
      T w = E::zero();
      for (int64_t r = j + 1 + tm.lane; r < rows; r += L) w = E::fma(E::conj(v[r]), y[r], w);
      w = tm.template sum_t<T>(w);
      w = E::add(w, y[j]);
      const T f = E::mul(tj, w);
#ifdef __CUDA_ARCH__
      __syncwarp();
#else
      BPX_HOST_SYNCWARP(tm);
#endif
      for (int64_t r = j + 1 + tm.lane; r < rows; r += L) y[r] = sub(y[r], E::mul(f, v[r]));
      if (tm.lane == 0) y[j] = sub(y[j], f);
#ifdef __CUDA_ARCH__
      __syncwarp();
#else
      BPX_HOST_SYNCWARP(tm);
#endif
    }
  }
  tm.sync();
}

// ---- gauge one side: tensor -> matrix view with every external leg multiplied by its X ---------------------------
// Returns the index (0 / 1) of the ping-pong buffer that holds the gauged matrix.  Gauge matrices stay in `gz`
// (per external leg, slot order: X (chi^2), Xinv (chi^2), eigenvalues (chi)).
template <typename T>
__host__ __device__ int gauge_side(const Team& tm, const Side& sd, const T* a, const T* msgs, T* buf0, T* buf1, T* gz,
                                   int* flag) {
  tensor_to_matrix<T>(tm, sd, a, buf0);
  int cur = 0;
  int64_t st = 1;
  T* g = gz;
  for (int i = 0; i < sd.z; ++i) {
    if (i == sd.bond_slot) continue;
    const int chi = sd.dim[i];
    T* bx = g;
    T* vx = g + (int64_t)chi * chi;
    double* ev = reinterpret_cast<double*>(g + 2 * (int64_t)chi * chi);
    gauge_from_message<T>(tm, msgs + sd.in_msg[i], chi, bx, vx, ev, flag);
    mode_product<T>(tm, cur ? buf1 : buf0, cur ? buf0 : buf1, sd.rows, sd.cols, st, chi, bx);
    cur ^= 1;
    st *= chi;
    g += 2 * (int64_t)chi * chi + chi;
  }
  return cur;
}

// inverse gauges on every external leg: T[.., l, ..] = sum_g Xinv[l, g] T[.., g, ..]; returns the buffer index
template <typename T>
__host__ __device__ int ungauge_side(const Team& tm, const Side& sd, T* buf0, T* buf1, int cur, int ncols, const T* gz) {
  int64_t st = 1;
  const T* g = gz;
  for (int i = 0; i < sd.z; ++i) {
    if (i == sd.bond_slot) continue;
    const int chi = sd.dim[i];
    const T* vx = g + (int64_t)chi * chi;
    mode_product<T>(tm, cur ? buf1 : buf0, cur ? buf0 : buf1, sd.rows, ncols, st, chi, vx);
    cur ^= 1;
    st *= chi;
    g += 2 * (int64_t)chi * chi + chi;
  }
  return cur;
}

// ---- one two-site gate -----------------------------------------------------------------------------------------
// op[o1, o2, i1, i2] column-major (1 = first side).  sv_out: chi_b doubles (kept singular values, zero padded).
template <typename T>
__host__ __device__ void run_two_site(const Team& tm, const GateDesc& gd, T* sites, T* msgs, const T* ops, T* ws,
                                      double* sv_out, int normalize, int* flag) {
  using E = Elem<T>;
  const Layout L = layout_of(gd);
  T* w = ws + gd.ws_off;
  int pbuf[2], nr[2];
  for (int a = 0; a < 2; ++a) {
    const Side& sd = gd.s[a];
    T* b0 = w + L.buf[a][0];
    T* b1 = w + L.buf[a][1];
    pbuf[a] = gauge_side<T>(tm, sd, sites + sd.site_off, msgs, b0, b1, w + L.gauge[a], flag);
    T* P = pbuf[a] ? b1 : b0;
    nr[a] = householder_qr<T>(tm, P, sd.rows, sd.cols, w + L.tau[a]);
    T* R = w + L.r[a];
    for (int i = tm.tid(); i < sd.nref * sd.cols; i += tm.nt()) {
      const int q = i % sd.nref, c = i / sd.nref;
      R[i] = q <= c ? P[q + sd.rows * c] : E::zero();
    }
    tm.sync();
  }
  const Side& s1 = gd.s[0];
  const Side& s2 = gd.s[1];
  const int d1 = s1.d, d2 = s2.d, n1 = s1.nref, n2 = s2.nref, chi = gd.chi_b;
  const int m = n1 * d1, n = n2 * d2;
  const T* R1 = w + L.r[0];
  const T* R2 = w + L.r[1];
  T* th0 = w + L.theta[0];
  T* th1 = w + L.theta[1];
  // theta[(q1, s1), (q2, s2)] = sum_b R1[q1, (s1, b)] R2[q2, (s2, b)]      (apply_operators.jl:260, R_v1 * R_v2)
  for (int i = tm.tid(); i < m * n; i += tm.nt()) {
    const int row = i % m, col = i / m;
    const int q1 = row % n1, x1 = row / n1, q2 = col % n2, x2 = col / n2;
    T acc = E::zero();
    for (int b = 0; b < chi; ++b) acc = E::fma(R1[q1 + n1 * (x1 + d1 * b)], R2[q2 + n2 * (x2 + d2 * b)], acc);
    th0[i] = acc;
  }
  tm.sync();
  // the gate on the two site legs (ITensorBase.apply)
  const T* op = ops + gd.op_off;
  const int dd = d1 * d2;
  for (int i = tm.tid(); i < m * n; i += tm.nt()) {
    const int row = i % m, col = i / m;
    const int q1 = row % n1, o1 = row / n1, q2 = col % n2, o2 = col / n2;
    T acc = E::zero();
    for (int x2 = 0; x2 < d2; ++x2)
      for (int x1 = 0; x1 < d1; ++x1)
        acc = E::fma(op[o1 + d1 * o2 + dd * (x1 + d1 * x2)], th0[(q1 + n1 * x1) + m * (q2 + n2 * x2)], acc);
    th1[i] = acc;
  }
  tm.sync();
  // SVD by one-sided Jacobi: th1 Vs = U diag(sigma)
  T* Vs = w + L.vs;
  double* sig = reinterpret_cast<double*>(w + L.sig);
  int32_t* order = reinterpret_cast<int32_t*>(w + L.order);
  jacobi_cols<T>(tm, th1, m, n, Vs, flag);
  for (int j = tm.tid(); j < n; j += tm.nt()) {
    double a = 0.0;
    for (int r = 0; r < m; ++r) a += E::abs2(th1[r + m * j]);
    sig[j] = sqrt(a);
  }
  tm.sync();
  if (tm.tid() == 0) {  // descending order (n <= a few dozen: insertion sort; ties keep the column order)
    for (int j = 0; j < n; ++j) {
      int pos = j;
      while (pos > 0 && sig[order[pos - 1]] < sig[j]) {
        order[pos] = order[pos - 1];
        --pos;
      }
      order[pos] = j;
    }
  }
  tm.sync();
  const int k = gd.k;
  double nrm = 1.0;
  if (normalize) {
    double a = 0.0;
    for (int j = 0; j < k; ++j) a += sig[order[j]] * sig[order[j]];
    nrm = a > 0.0 ? sqrt(a) : 1.0;
  }
  // new R factors into the free ping-pong buffers: Y1[q1, (s1, kk)] = U[(q1, s1), j] sqrt(s_j), Y2[q2, (s2, kk)] =
  // sqrt(s_j) conj(Vs[(q2, s2), j]); rows >= nref are zero                                       (:265-268)
  T* Y[2];
  for (int a = 0; a < 2; ++a) {
    const Side& sd = gd.s[a];
    Y[a] = w + L.buf[a][pbuf[a] ^ 1];
    const int64_t tot = sd.rows * (int64_t)(sd.d * k);
    const int na = sd.nref, da = sd.d;
    for (int64_t i = tm.tid(); i < tot; i += tm.nt()) {
      const int64_t q = i % sd.rows, c = i / sd.rows;
      T v = E::zero();
      if (q < na) {
        const int x = (int)(c % da), kk = (int)(c / da), j = order[kk];
        const double sj = sig[j], snew = sj / nrm;
        if (sj > 0.0) {
          if (a == 0)
            v = scal(th1[(q + na * x) + (int64_t)m * j], sqrt(snew) / sj);
          else
            v = scal(E::conj(Vs[(q + na * x) + (int64_t)n * j]), sqrt(snew));
        }
      }
      Y[a][i] = v;
    }
  }
  tm.sync();
  for (int a = 0; a < 2; ++a) {
    const Side& sd = gd.s[a];
    T* b0 = w + L.buf[a][0];
    T* b1 = w + L.buf[a][1];
    const T* P = pbuf[a] ? b1 : b0;
    const int ncols = sd.d * k;
    apply_q<T>(tm, P, sd.rows, nr[a], w + L.tau[a], Y[a], ncols);                        // Q_v * R_v         (:270-271)
    const int fin = ungauge_side<T>(tm, sd, b0, b1, pbuf[a] ^ 1, ncols, w + L.gauge[a]);  // inverse gauges
    const T* F = fin ? b1 : b0;
    T* out = sites + sd.site_off;
    for (int64_t i = tm.tid(); i < sd.n; i += tm.nt()) {
      int64_t row;
      int col;
      split_index(sd, i, row, col);
      out[i] = (col / sd.d) < k ? F[row + sd.rows * col] : E::zero();   // kept rank zero-padded up to chi_b
    }
  }
  // new messages on the gate edge: diag(S) in both directions                                     (:273-277)
  for (int i = tm.tid(); i < chi * chi; i += tm.nt()) {
    const int r = i % chi, c = i / chi;
    const T v = (r == c && r < k) ? from_real<T>(sig[order[r]] / nrm) : E::zero();
    msgs[gd.msg12 + i] = v;
    msgs[gd.msg21 + i] = v;
  }
  if (sv_out)
    for (int i = tm.tid(); i < chi; i += tm.nt()) sv_out[i] = i < k ? sig[order[i]] / nrm : 0.0;
  tm.sync();
}

// ---- one one-site gate (apply_operators.jl:226-244) -----------------------------------------------------------
// op[o, i] column-major.  normalize: divide by the Frobenius norm of the new tensor with every leg gauged.
// scratch_sum: one double visible to the whole team (shared memory on the device).
template <typename T>
__host__ __device__ void run_one_site(const Team& tm, const GateDesc& gd, T* sites, const T* msgs, const T* ops, T* ws,
                                      int normalize, int* flag, double* scratch_sum) {
  using E = Elem<T>;
  const Side& sd = gd.s[0];
  const T* op = ops + gd.op_off;
  T* a = sites + sd.site_off;
  const int d = sd.d;
  const int64_t nrest = sd.n / d;
  // in place: every thread owns whole physical fibres
  for (int64_t r = tm.tid(); r < nrest; r += tm.nt()) {
    T* f = a + r * d;
    T tmp[16];
    for (int o = 0; o < d; ++o) {
      T acc = E::zero();
      for (int i = 0; i < d; ++i) acc = E::fma(op[o + d * i], f[i], acc);
      tmp[o] = acc;
    }
    for (int o = 0; o < d; ++o) f[o] = tmp[o];
  }
  tm.sync();
  if (!normalize) return;
  const Layout L = layout_of(gd);
  T* w = ws + gd.ws_off;
  T* b0 = w + L.buf[0][0];
  T* b1 = w + L.buf[0][1];
  const int cur = gauge_side<T>(tm, sd, a, msgs, b0, b1, w + L.gauge[0], flag);
  const T* G = cur ? b1 : b0;
  if (tm.tid() == 0) *scratch_sum = 0.0;
  tm.sync();
  double part = 0.0;
  for (int64_t i = tm.tid(); i < sd.n; i += tm.nt()) part += E::abs2(G[i]);
  part = tm.sum(part);
#ifdef __CUDA_ARCH__
  if (tm.lane == 0) atomicAdd(scratch_sum, part);
#else
  if (tm.lane == 0) BPX_HOST_ATOMIC_ADD(scratch_sum, part);
#endif
  tm.sync();
  const double nrm = sqrt(*scratch_sum);
  if (nrm > 0.0)
    for (int64_t i = tm.tid(); i < sd.n; i += tm.nt()) a[i] = scal(a[i], 1.0 / nrm);
  tm.sync();
}

template <typename T>
__host__ __device__ void run_gate(const Team& tm, const GateDesc& gd, T* sites, T* msgs, const T* ops, T* ws,
                                  double* sv_out, int normalize, int* flag, double* scratch_sum) {
  if (gd.nsides == 2)
    run_two_site<T>(tm, gd, sites, msgs, ops, ws, sv_out, normalize, flag);
  else
    run_one_site<T>(tm, gd, sites, msgs, ops, ws, normalize, flag, scratch_sum);
}

#ifdef __CUDACC__
struct ApplyArgs {
  const GateDesc* gates;
  int64_t n_gates;
  void* sites;
  void* msgs;
  const void* ops;
  void* ws;
  double* sv_out;      // [n_gates][sv_stride] or NULL
  int64_t sv_stride;
  int normalize;
};

// one CTA per gate (grid-stride over the batch); gates of a batch are vertex-disjoint, so CTAs never touch the same
// tensor or message
template <typename T>
__global__ void __launch_bounds__(NT) bp_apply_gates(ApplyArgs a) {
  __shared__ int flag;
  __shared__ double ssum;
  Team tm;
  tm.lane = threadIdx.x & 31;
  tm.wid = threadIdx.x >> 5;
  tm.nw = NT / 32;
  for (int64_t g = blockIdx.x; g < a.n_gates; g += gridDim.x) {
    const GateDesc& gd = a.gates[g];
    run_gate<T>(tm, gd, static_cast<T*>(a.sites), static_cast<T*>(a.msgs), static_cast<const T*>(a.ops),
                static_cast<T*>(a.ws), a.sv_out ? a.sv_out + (gd.sv_row_p1 > 0 ? gd.sv_row_p1 - 1 : g) * a.sv_stride : nullptr,
                a.normalize, &flag, &ssum);
    __syncthreads();
  }
}
#endif

}  // namespace applyk
}  // namespace bpx
