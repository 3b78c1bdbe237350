// Two-site gate application, version 2 (OPT-IN: BPX_APPLY_V2=1; bp_apply_gates of bpx_apply.cuh stays the default until
// this one has been MEASURED on a B200).  Developed on the host (tests/native: host-compiled device code, ThreadSanitizer
// schedules) when round 1's GPU budget was all but spent; its first GPU run matched version 1 in all six cases of
// tests/test_zzzz_apply_large_and_v2_gpu.py::test_v2_matches_v1 (profiles/r1o_late_pytest_gpu.txt).
//
// Same algorithm and same building blocks as version 1 (apply_operators.jl:246-283; jacobi_cols, gauge_from_message,
// householder_qr, apply_q, mode_product), reorganised around MEMORY TRAFFIC.  Version 1 keeps the rows x cols matrix
// view of a tensor in global memory and streams it ~2 (z - 1) times for the gauges and ~cols times for the Householder
// sweeps (cfg5's bulk, 4096 x 32: ~80 MB of traffic per tensor for 1 MiB of data).  Here:
//   * gauging / un-gauging run COLUMN BY COLUMN in shared memory: a column (all external legs, fixed physical and bond
//     index; chi^(z-1) elements = 32 KiB at cfg5) is gathered once, multiplied by every leg's X (or X^-1) in a
//     shared-memory ping-pong, and stored once;
//   * the QR is a flat-tree TSQR: row blocks of the matrix view are stacked under the R factor carried from the previous
//     block and factorised IN SHARED MEMORY ((cols + RB) x cols panel); the reflector panels are written to global
//     memory once and read once when Q is applied (blocks in reverse order, the coefficient block carried upwards).
//     Householder throughout: the stability of version 1's QR, no Gram matrix.
// Global traffic per tensor: A read, P written, P read, panels written, panels read, Y written, Y read, A' written
// = 8 passes (16 MiB per gate at cfg5) against ~25 flop per algorithmic byte: the FP64 pipe becomes the bound.
#pragma once
#include "bpx_apply.cuh"

namespace bpx {
namespace applyk2 {

using namespace bpx::applyk;

// block rows of the TSQR for a side, given the shared-memory budget (elements of T): panels S ((cols + RB) x cols) and
// W ((cols + RB) x cols) plus the carried cols x cols block must fit.  0: the side does not fit version 2.
__host__ __device__ inline int64_t block_rows(const Side& s, int64_t smem_elems) {
  const int64_t c = s.cols;
  if (2 * s.rows > smem_elems) return 0;  // the two column buffers of the gauging passes
  int64_t rb = (smem_elems - c * c) / (2 * c) - c;
  if (rb < 1) return 0;
  if (rb >= 32) rb &= ~(int64_t)31;
  return rb < s.rows ? rb : s.rows;
}

struct Layout2 {
  int64_t py[2];      // P (gauged matrix view), later Q Y
  int64_t v[2];       // reflector panels, nb x (cols + RB) x cols
  int64_t tau[2];     // nb x cols
  int64_t gauge[2];
  int64_t r[2];       // nref x cols
  int64_t ytop[2];    // nref x cols: the new R factor (coefficients of Q's columns)
  int64_t theta[2], vs, sig, order;
  int64_t rb[2], nb[2];
  int64_t total;
};

__host__ __device__ inline Layout2 layout2_of(const GateDesc& g, int64_t smem_elems) {
  Layout2 L;
  int64_t o = 0;
  for (int a = 0; a < 2; ++a) {
    const Side& s = g.s[a];
    const int64_t rb = block_rows(s, smem_elems);
    const int64_t nb = rb > 0 ? (s.rows + rb - 1) / rb : 0;
    L.rb[a] = rb;
    L.nb[a] = nb;
    L.py[a] = o; o += s.n;
    L.v[a] = o; o += nb * (s.cols + rb) * s.cols;
    L.tau[a] = o; o += nb * s.cols;
    L.gauge[a] = o; o += gauge_elems(s);
    L.r[a] = o; o += (int64_t)s.nref * s.cols;
    L.ytop[a] = o; o += (int64_t)s.nref * s.cols;
  }
  const int64_t m = (int64_t)g.s[0].nref * g.s[0].d, n = (int64_t)g.s[1].nref * g.s[1].d;
  L.theta[0] = o; o += m * n;
  L.theta[1] = o; o += m * n;
  L.vs = o; o += n * n;
  L.sig = o; o += n;
  L.order = o; o += n;
  L.total = o;
  return L;
}

// canonical element index of A_v[s, l_0..l_{z-1}] from the matrix-view coordinates (inverse of split_index)
__host__ __device__ __forceinline__ int64_t join_index(const Side& sd, int64_t row, int col) {
  const int s = col % sd.d, bond = col / sd.d;
  int64_t idx = 0, stride = 1;
  for (int k = 0; k < sd.z; ++k) {
    int64_t l;
    if (k == sd.bond_slot) {
      l = bond;
    } else {
      l = row % sd.dim[k];
      row /= sd.dim[k];
    }
    idx += l * stride;
    stride *= sd.dim[k];
  }
  return s + sd.d * idx;
}

// jacobi_cols with B (m x n) and V (n x n) staged through shared memory when they fit: every rotation step of the
// round-robin schedule is a dependent round trip to its operands, a few hundred ns in global memory, tens in shared
template <typename T>
__host__ __device__ void jacobi_cols_staged(const Team& tm, T* B, int m, int n, T* V, int* flag, T* smem, int64_t smem_elems) {
  const int64_t nb = (int64_t)m * n, nv = (int64_t)n * n;
  if (nb + nv > smem_elems) {
    jacobi_cols<T, true>(tm, B, m, n, V, flag);
    return;
  }
  T* sb = smem;
  T* sv = smem + nb;
  for (int64_t i = tm.tid(); i < nb; i += tm.nt()) sb[i] = B[i];
  tm.sync();
  jacobi_cols<T, true>(tm, sb, m, n, sv, flag);  // phase-stable rotation (see jacobi_cols)
  for (int64_t i = tm.tid(); i < nb; i += tm.nt()) B[i] = sb[i];
  for (int64_t i = tm.tid(); i < nv; i += tm.nt()) V[i] = sv[i];
  tm.sync();
}

// gauge_from_message of version 1 with the eigen-decomposition staged through shared memory (same arithmetic)
template <typename T>
__host__ __device__ void gauge_from_message_staged(const Team& tm, const T* msg, int chi, T* bx, T* vx, double* ev, int* flag,
                                                   T* smem, int64_t smem_elems) {
  using E = Elem<T>;
  for (int i = tm.tid(); i < chi * chi; i += tm.nt()) {
    const int r = i % chi, c = i / chi;
    bx[i] = scal(E::add(msg[r + c * chi], E::conj(msg[c + r * chi])), 0.5);  // Hermitian part
  }
  tm.sync();
  jacobi_cols_staged<T>(tm, bx, chi, chi, vx, flag, smem, smem_elems);
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
    const int g = i % chi, l = i / chi;
    const double d = ev[g];
    bx[i] = d > cut ? scal(E::conj(vx[l + g * chi]), sqrt(d)) : E::zero();
  }
  tm.sync();
  for (int i = tm.tid(); i < chi * chi; i += tm.nt()) {
    const double d = ev[i / chi];
    vx[i] = d > cut ? scal(vx[i], 1.0 / sqrt(d)) : E::zero();
  }
  tm.sync();
}

// One column (a vector over the external legs) times a chi x chi matrix on the leg with stride `st`: a thread owns whole
// FIBRES (the chi elements along that leg), keeps them in registers and produces the chi outputs from them, so the index
// arithmetic and the shared-memory loads are paid once per fibre instead of once per output (the generic mode_product
// of version 1 decomposes the index and re-loads the inputs for every output element).  x is read as a broadcast.
template <typename T, int CHI>
__host__ __device__ __forceinline__ void mode_fibres_fixed(const Team& tm, const T* in, T* out, int64_t rows, int64_t st,
                                                           const T* x) {
  using E = Elem<T>;
  const int64_t nf = rows / CHI;
  for (int64_t f = tm.tid(); f < nf; f += tm.nt()) {
    const int64_t lo = f % st, hi = f / st, base = hi * st * CHI + lo;
    T v[CHI];
#pragma unroll
    for (int l = 0; l < CHI; ++l) v[l] = in[base + l * st];
#pragma unroll 1  // unrolling g as well hoists CHI^2 loads of x: 10 KB of spills per thread at CHI = 16 complex
    for (int g = 0; g < CHI; ++g) {
      T acc = E::zero();
#pragma unroll
      for (int l = 0; l < CHI; ++l) acc = E::fma(x[g + CHI * l], v[l], acc);
      out[base + g * st] = acc;
    }
  }
  tm.sync();
}

template <typename T>
__host__ __device__ void mode_fibres(const Team& tm, const T* in, T* out, int64_t rows, int64_t st, int chi, const T* x) {
  switch (chi) {
    case 2: mode_fibres_fixed<T, 2>(tm, in, out, rows, st, x); break;
    case 3: mode_fibres_fixed<T, 3>(tm, in, out, rows, st, x); break;
    case 4: mode_fibres_fixed<T, 4>(tm, in, out, rows, st, x); break;
    case 8: mode_fibres_fixed<T, 8>(tm, in, out, rows, st, x); break;
    case 16: mode_fibres_fixed<T, 16>(tm, in, out, rows, st, x); break;
    default: mode_product<T>(tm, in, out, rows, 1, st, chi, x);
  }
}

// Column c of the matrix view <-> the canonical tensor: the index decomposition is done once per run along the FIRST
// external leg (whose elements are `stride0` apart in the canonical layout), not once per element.
struct ColumnWalk {
  int64_t n0, stride0, nruns;  // first external leg: dimension, canonical stride (elements); number of runs = rows / n0
};
__host__ __device__ inline ColumnWalk column_walk(const Side& sd) {
  ColumnWalk w;
  w.n0 = 1;
  w.stride0 = 0;
  int64_t stride = sd.d;
  for (int k = 0; k < sd.z; ++k) {
    if (k != sd.bond_slot) {
      w.n0 = sd.dim[k];
      w.stride0 = stride;
      break;
    }
    stride *= sd.dim[k];
  }
  w.nruns = sd.rows / w.n0;
  return w;
}

template <typename T>
__host__ __device__ void gather_column(const Team& tm, const Side& sd, const T* a, int c, T* col) {
  const ColumnWalk w = column_walk(sd);
  for (int64_t run = tm.tid(); run < w.nruns; run += tm.nt()) {
    const int64_t r0 = run * w.n0;
    const T* src = a + join_index(sd, r0, c);
    for (int64_t t = 0; t < w.n0; ++t) col[r0 + t] = src[t * w.stride0];
  }
  tm.sync();
}

template <typename T>
__host__ __device__ void scatter_column(const Team& tm, const Side& sd, const T* col, int c, T* a) {
  const ColumnWalk w = column_walk(sd);
  for (int64_t run = tm.tid(); run < w.nruns; run += tm.nt()) {
    const int64_t r0 = run * w.n0;
    T* dst = a + join_index(sd, r0, c);
    for (int64_t t = 0; t < w.n0; ++t) dst[t * w.stride0] = col[r0 + t];
  }
  tm.sync();
}

// every external leg of one column multiplied by its gauge matrix, in a shared-memory ping-pong; `which` = 0: X, 1: X^-1.
// Returns the buffer (c0 or c1) that holds the result.
template <typename T>
__host__ __device__ T* gauge_column(const Team& tm, const Side& sd, T* c0, T* c1, const T* gz, int which) {
  T* cur = c0;
  T* oth = c1;
  int64_t st = 1;
  const T* g = gz;
  for (int i = 0; i < sd.z; ++i) {
    if (i == sd.bond_slot) continue;
    const int chi = sd.dim[i];
    mode_fibres<T>(tm, cur, oth, sd.rows, st, chi, g + (which ? (int64_t)chi * chi : 0));
    T* t = cur; cur = oth; oth = t;
    st *= chi;
    g += 2 * (int64_t)chi * chi + chi;
  }
  return cur;
}

// Phase 1: gauges from the messages, then P[:, c] = (X_1 x X_2 x ...) A[:, c] column by column
template <typename T>
__host__ __device__ void gauged_matrix(const Team& tm, const Side& sd, const T* a, const T* msgs, T* P, T* gz, T* smem,
                                       int64_t smem_elems, int* flag) {
  T* g = gz;
  for (int i = 0; i < sd.z; ++i) {
    if (i == sd.bond_slot) continue;
    const int chi = sd.dim[i];
    gauge_from_message_staged<T>(tm, msgs + sd.in_msg[i], chi, g, g + (int64_t)chi * chi,
                                 reinterpret_cast<double*>(g + 2 * (int64_t)chi * chi), flag, smem, smem_elems);
    g += 2 * (int64_t)chi * chi + chi;
  }
  T* c0 = smem;
  T* c1 = smem + sd.rows;
  for (int c = 0; c < sd.cols; ++c) {
    gather_column<T>(tm, sd, a, c, c0);
    const T* res = gauge_column<T>(tm, sd, c0, c1, gz, 0);
    T* pc = P + sd.rows * c;
    for (int64_t r = tm.tid(); r < sd.rows; r += tm.nt()) pc[r] = res[r];
    tm.sync();
  }
}

// Phase 2: flat-tree TSQR of P (rows x cols) with row blocks of RB; panels -> V, scalars -> tau, final R (nref x cols) -> R
template <typename T>
__host__ __device__ void tsqr_factor(const Team& tm, const Side& sd, const T* P, int64_t RB, T* V, T* tau, T* R, T* smem) {
  using E = Elem<T>;
  const int cols = sd.cols;
  const int64_t rows = sd.rows, pstride = (cols + RB) * cols;
  T* Rc = smem;
  T* S = smem + (int64_t)cols * cols;
  int64_t done = 0, blk = 0;
  while (done < rows) {
    const int64_t h = RB < rows - done ? RB : rows - done;
    const int64_t rc = done < cols ? done : cols, ld = rc + h;
    for (int64_t i = tm.tid(); i < ld * cols; i += tm.nt()) {
      const int64_t r = i % ld, c = i / ld;
      S[i] = r < rc ? Rc[r + cols * c] : P[(done + r - rc) + rows * c];
    }
    tm.sync();
    householder_qr<T>(tm, S, ld, cols, tau + blk * cols);
    T* Vb = V + blk * pstride;
    for (int64_t i = tm.tid(); i < ld * cols; i += tm.nt()) Vb[i] = S[i];
    const int64_t rn = ld < cols ? ld : cols;
    for (int64_t i = tm.tid(); i < (int64_t)cols * cols; i += tm.nt()) {
      const int64_t r = i % cols, c = i / cols;
      Rc[i] = (r < rn && r <= c) ? S[r + ld * c] : E::zero();
    }
    tm.sync();
    done += h;
    ++blk;
  }
  for (int i = tm.tid(); i < sd.nref * cols; i += tm.nt()) R[i] = Rc[(i % sd.nref) + cols * (i / sd.nref)];
  tm.sync();
}

// Phase 4: Y = Q [Ytop; 0]: blocks in reverse order, the coefficient block carried upwards.  Ytop is nref x ncols.
template <typename T>
__host__ __device__ void tsqr_apply_q(const Team& tm, const Side& sd, int64_t RB, const T* V, const T* tau, const T* Ytop,
                                      int ncols, T* Y, T* smem) {
  using E = Elem<T>;
  const int cols = sd.cols;
  const int64_t rows = sd.rows, pstride = (cols + RB) * cols;
  T* Cc = smem;                                   // carried coefficients, cols x ncols (leading dimension cols)
  T* S = smem + (int64_t)cols * cols;             // reflector panel
  T* W = S + pstride;                             // right-hand side panel, ld x ncols
  for (int i = tm.tid(); i < sd.nref * ncols; i += tm.nt()) Cc[(i % sd.nref) + cols * (i / sd.nref)] = Ytop[i];
  tm.sync();
  const int64_t nb = (rows + RB - 1) / RB;
  for (int64_t blk = nb - 1; blk >= 0; --blk) {
    const int64_t done = blk * RB;
    const int64_t h = RB < rows - done ? RB : rows - done;
    const int64_t rc = done < cols ? done : cols, ld = rc + h, rn = ld < cols ? ld : cols;
    for (int64_t i = tm.tid(); i < ld * ncols; i += tm.nt()) {
      const int64_t r = i % ld, c = i / ld;
      W[i] = r < rn ? Cc[r + cols * c] : E::zero();
    }
    const T* Vb = V + blk * pstride;
    for (int64_t i = tm.tid(); i < ld * cols; i += tm.nt()) S[i] = Vb[i];
    tm.sync();
    const int nr = (int)((ld - 1) < (int64_t)cols ? (ld - 1) : (int64_t)cols);
    apply_q<T>(tm, S, ld, nr, tau + blk * cols, W, ncols);
    for (int64_t i = tm.tid(); i < h * ncols; i += tm.nt()) {
      const int64_t r = i % h, c = i / h;
      Y[(done + r) + rows * c] = W[(rc + r) + ld * c];
    }
    for (int64_t i = tm.tid(); i < rc * ncols; i += tm.nt()) {
      const int64_t r = i % rc, c = i / rc;
      Cc[r + cols * c] = W[r + ld * c];
    }
    tm.sync();
  }
}

// Phase 5: inverse gauges column by column, scatter into the canonical layout, kept rank zero-padded up to chi_b
template <typename T>
__host__ __device__ void ungauge_and_store(const Team& tm, const Side& sd, const T* Y, int ncols, int k, const T* gz, T* out,
                                           T* smem) {
  using E = Elem<T>;
  T* c0 = smem;
  T* c1 = smem + sd.rows;
  for (int c = 0; c < ncols; ++c) {
    const T* yc = Y + sd.rows * c;
    for (int64_t r = tm.tid(); r < sd.rows; r += tm.nt()) c0[r] = yc[r];
    tm.sync();
    const T* res = gauge_column<T>(tm, sd, c0, c1, gz, 1);
    scatter_column<T>(tm, sd, res, c, out);
  }
  if (k * sd.d < sd.cols) {
    for (int64_t i = tm.tid(); i < sd.n; i += tm.nt()) {
      int64_t row;
      int col;
      split_index(sd, i, row, col);
      if (col / sd.d >= k) out[i] = E::zero();
    }
  }
  tm.sync();
}

// One two-site gate.  smem: `smem_elems` elements of T visible to the whole team (dynamic shared memory on the device).
template <typename T>
__host__ __device__ void run_two_site_v2(const Team& tm, const GateDesc& gd, T* sites, T* msgs, const T* ops, T* ws,
                                         double* sv_out, int normalize, int* flag, T* smem, int64_t smem_elems) {
  using E = Elem<T>;
  const Layout2 L = layout2_of(gd, smem_elems);
  T* w = ws + gd.ws_off;
  for (int a = 0; a < 2; ++a) {
    const Side& sd = gd.s[a];
    gauged_matrix<T>(tm, sd, sites + sd.site_off, msgs, w + L.py[a], w + L.gauge[a], smem, smem_elems, flag);
    tsqr_factor<T>(tm, sd, w + L.py[a], L.rb[a], w + L.v[a], w + L.tau[a], w + L.r[a], smem);
  }
  // ---- the bond problem: identical to version 1 (theta, gate, Jacobi SVD, order, normalisation) ------------------
  const Side& s1 = gd.s[0];
  const Side& s2 = gd.s[1];
  const int d1 = s1.d, d2 = s2.d, n1 = s1.nref, n2 = s2.nref, chi = gd.chi_b;
  const int m = n1 * d1, n = n2 * d2;
  const T* R1 = w + L.r[0];
  const T* R2 = w + L.r[1];
  T* th0 = w + L.theta[0];
  T* th1 = w + L.theta[1];
  for (int i = tm.tid(); i < m * n; i += tm.nt()) {
    const int row = i % m, col = i / m;
    const int q1 = row % n1, x1 = row / n1, q2 = col % n2, x2 = col / n2;
    T acc = E::zero();
    for (int b = 0; b < chi; ++b) acc = E::fma(R1[q1 + n1 * (x1 + d1 * b)], R2[q2 + n2 * (x2 + d2 * b)], acc);
    th0[i] = acc;
  }
  tm.sync();
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
  T* Vs = w + L.vs;
  double* sig = reinterpret_cast<double*>(w + L.sig);
  int32_t* order = reinterpret_cast<int32_t*>(w + L.order);
  jacobi_cols_staged<T>(tm, th1, m, n, Vs, flag, smem, smem_elems);
  for (int j = tm.tid(); j < n; j += tm.nt()) {
    double a = 0.0;
    for (int r = 0; r < m; ++r) a += E::abs2(th1[r + m * j]);
    sig[j] = sqrt(a);
  }
  tm.sync();
  if (tm.tid() == 0) {
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
  // new R factors: Ytop_1[q1, (s1, kk)] = U[(q1, s1), j] sqrt(s_j), Ytop_2[q2, (s2, kk)] = sqrt(s_j) conj(Vs[(q2, s2), j])
  for (int a = 0; a < 2; ++a) {
    const Side& sd = gd.s[a];
    T* yt = w + L.ytop[a];
    const int na = sd.nref, da = sd.d;
    for (int i = tm.tid(); i < na * da * k; i += tm.nt()) {
      const int q = i % na, c = i / na, x = c % da, kk = c / da, j = order[kk];
      const double sj = sig[j], snew = sj / nrm;
      T v = E::zero();
      if (sj > 0.0)
        v = a == 0 ? scal(th1[(q + na * x) + (int64_t)m * j], sqrt(snew) / sj)
                   : scal(E::conj(Vs[(q + na * x) + (int64_t)n * j]), sqrt(snew));
      yt[i] = v;
    }
  }
  tm.sync();
  for (int a = 0; a < 2; ++a) {
    const Side& sd = gd.s[a];
    const int ncols = sd.d * k;
    tsqr_apply_q<T>(tm, sd, L.rb[a], w + L.v[a], w + L.tau[a], w + L.ytop[a], ncols, w + L.py[a], smem);
    ungauge_and_store<T>(tm, sd, w + L.py[a], ncols, k, w + L.gauge[a], sites + sd.site_off, smem);
  }
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

#ifdef __CUDACC__
struct ApplyArgs2 {
  ApplyArgs base;
  int64_t smem_elems;
};

template <typename T>
__global__ void __launch_bounds__(NT) bp_apply_gates_v2(ApplyArgs2 a2) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ int flag;
  const ApplyArgs& a = a2.base;
  Team tm;
  tm.lane = threadIdx.x & 31;
  tm.wid = threadIdx.x >> 5;
  tm.nw = NT / 32;
  for (int64_t g = blockIdx.x; g < a.n_gates; g += gridDim.x) {
    const int64_t sv_row = a.gates[g].sv_row_p1 > 0 ? a.gates[g].sv_row_p1 - 1 : g;
    run_two_site_v2<T>(tm, a.gates[g], static_cast<T*>(a.sites), static_cast<T*>(a.msgs), static_cast<const T*>(a.ops),
                       static_cast<T*>(a.ws), a.sv_out ? a.sv_out + sv_row * a.sv_stride : nullptr, a.normalize, &flag,
                       reinterpret_cast<T*>(dyn_smem), a2.smem_elems);
    __syncthreads();
  }
}
#endif

}  // namespace applyk2
}  // namespace bpx
