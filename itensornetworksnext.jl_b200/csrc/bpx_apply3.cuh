// Two-site gate application, version 3: the GRAM path (default for the shapes it takes; versions 1 / 2 stay as the fallback
// for everything else and for gates this kernel declines at run time).
//
// Same result as apply_operators.jl:246-283, different route.  With full-rank boundary messages M_i = X_i^H X_i the
// reference's chain  gauge -> QR -> gate/SVD -> Q R' -> inverse gauge  collapses algebraically:
//     P = (X_1 x X_2 x ..) A                gauged matrix view, rows = external legs, cols = (s, bond)
//     G = P^H P = A^H (M_1 x M_2 x ..) A    the "vertex environment with the site and bond legs open": messages absorbed
//                                           AS THEY ARE -- no eigen-decomposition of the messages, no X, no X^-1
//     G = V diag(lam) V^H,  R = diag(sqrt(lam)) V^H   (any R with R^H R = G is a valid "R factor": P = Q R with Q = P R^+)
//     theta = R_1 R_2 -> gate -> SVD -> Y_a (the new R factors), exactly as in versions 1 / 2
//     A'_a = X^-1 Q_a Y_a = X^-1 X A R^+ Y_a = A (R^+ Y_a) = A W_a          (X^-1 X = 1 for full-rank messages)
// so per side the heavy work is three passes of plain dense contractions over the tensor -- absorb the messages (column
// by column in shared memory), one (rows x cols)^H (rows x cols) Gram product, one (rows x cols)(cols x cols) product --
// instead of ~80 (version 1) or 8 (version 2) latency-bound Householder passes.  No Householder QR, no TSQR, no Q.
//
// When the shortcut is NOT taken (the gate is left untouched and reported in `status`, the caller re-runs it on version 2 /
// 1, which follow the reference step by step):
//   * a boundary message is not safely positive definite (Cholesky pivot <= PIVOT_MIN * largest diagonal entry): the
//     reference then projects on the message's support (pinv), X^-1 X != 1;
//   * a kept eigenvalue of G is below COND_MIN * the largest: forming G squares the condition number of P, the error of a
//     singular direction with weight sigma is eps * sigma_max / sigma relative to the tensor -- 1e-11 at the threshold;
//   * the Jacobi iteration did not converge.
// Exactly zero eigenvalues of G (zero-padded bonds) are dropped like the reference's pinv does.
//
// Serial parts: one-sided Jacobi WITHOUT accumulating V (for Hermitian G the rotated columns are lam_j v_j themselves; for
// the bond matrix only the k kept right singular vectors are needed: v_j = theta^H u_j / s_j), pairs handled by sub-warp
// lane groups so that one step of the round-robin schedule is one pass of the CTA (versions 1 / 2: one warp per pair).
// Written against the `Team` abstraction like versions 1 / 2: tests/native/apply_host.cu runs the phases on the host
// (run_two_site_v3).  On the device a gate runs as three kernels (bottom of the file: bp_apply3_sides / _bond / _final);
// the Float64 Gram and final passes have device-only tensor-pipe versions (gram_side_mma, final_side_mma: DMMA.8x8x4 fed by
// TMA bulk copies), checked against the step-by-step kernels and the oracle on the GPU (tests/test_zz_gpu_apply.py,
// tests/test_zzzz_apply_large_and_v2_gpu.py).
#pragma once
#include "bpx_apply2.cuh"

namespace bpx {
namespace applyk3 {

using namespace bpx::applyk;

constexpr int PC = 32;               // padded column count of the tiles; the fast path takes sides with cols <= PC
constexpr int TRG = 128;             // rows per tile of the Gram pass
constexpr int PCP = PC + 2;          // row stride of the scratch copies: 16-byte aligned rows
constexpr double COND_MIN = 1e-10;   // smallest kept eigenvalue of G relative to the largest
constexpr double PIVOT_MIN = 1e-12;  // smallest Cholesky pivot of a message relative to its largest diagonal entry
constexpr int MAXDIM = 16;           // largest external link dimension (a fibre lives in registers)

// Layout of the scratch copies `at` / `tt` (the matrix view, rows x cols): COLUMN-GROUP major.  A group is what an absorb
// batch produces -- 2 columns Float64, 1 column ComplexF64: 16 bytes per row -- and the copy is an array of 16-byte vectors
// V[group][row].  A batch's stores are then one contiguous stream (256 bytes per warp instruction), a tile of rows is one
// contiguous piece per group (bulk copies), and in shared memory the groups of a tile are `tile_stride16` vectors apart
// (TR + 4: the DMMA fragment loads of the Gram and final passes are conflict free).
// History (ncu, profiles/r2av_* / r2bg_*): plain row-major rows made every batch store half a 32-byte sector, the L2 did not
// keep the lines until the other half arrived (write hit rate 8 %) and every half went to DRAM as a read-modify-write:
// 11.2 MB of traffic per tensor instead of 5.4, the side kernel bound by it; pairing two rows per sector fixed the traffic
// (side kernel 2 x faster) but left 32-byte pieces 544 bytes apart.
template <typename T>
__host__ __device__ __forceinline__ int64_t gidx(int64_t r, int c, int64_t rows_e) {  // element (r, c) of a scratch copy
  if (Elem<T>::is_complex) return (int64_t)c * rows_e + r;
  return (((int64_t)(c >> 1) * rows_e + r) << 1) + (c & 1);
}
__host__ __device__ __forceinline__ int tile_stride16(bool cplx, int tr) { return cplx ? tr + 1 : tr + 4; }
template <typename T>
__host__ __device__ __forceinline__ int tidx(int r, int c, int tr) {  // element (r, c) of a tile of tr rows in shared memory
  if (Elem<T>::is_complex) return c * (tr + 1) + r;
  return (((c >> 1) * (tr + 4) + r) << 1) + (c & 1);
}
__host__ __device__ __forceinline__ int64_t rows_even(int64_t rows) { return (rows + 1) & ~(int64_t)1; }
__host__ __device__ __forceinline__ int64_t scratch_elems(int64_t rows) { return rows_even(rows) * PCP; }

struct Layout3 {
  int64_t at[2];        // per side: the matrix view of A, column-group major (gidx())
  int64_t tt;           // T = (M_1 x M_2 x ..) A in the same layout; shared by the two sides
  int64_t h[2];         // Hermitian parts of the boundary messages, slot order, chi^2 each
  int64_t g[2];         // G (cols x cols) and its rotated copy
  int64_t gb[2];
  int64_t ev[2];        // eigenvalues (doubles; cols T-slots reserved)
  int64_t r[2];         // R (cols x cols): row q = sqrt(lam_q) v_q^H (zero rows for dropped eigenvalues)
  int64_t rinv[2];      // R^+ (cols x cols): column q = v_q / sqrt(lam_q)
  int64_t y[2];         // Y (cols x d k) then W = R^+ Y (cols x PC, zero padded)
  int64_t w[2];
  int64_t theta[3], sig, order;
  int64_t tab[2];       // int32 address tables per side: rowtab[rows], coltab[PC]
  int64_t total;
};

__host__ __device__ inline int64_t herm_elems(const Side& s) {
  int64_t t = 0;
  for (int i = 0; i < s.z; ++i)
    if (i != s.bond_slot) t += (int64_t)s.dim[i] * s.dim[i];
  return t;
}

__host__ __device__ inline int64_t tt_elems(const GateDesc& g) { return scratch_elems(g.s[0].rows > g.s[1].rows ? g.s[0].rows : g.s[1].rows); }
// with_tt = false: the per-GATE work space of the device kernels (T lives in a per-CTA scratch buffer there)
__host__ __device__ inline Layout3 layout3_of(const GateDesc& g, bool with_tt = true) {
  Layout3 L;
  int64_t o = 0;
  for (int a = 0; a < 2; ++a) { L.at[a] = o; o += scratch_elems(g.s[a].rows); }
  L.tt = o; o += with_tt ? tt_elems(g) : 0;
  for (int a = 0; a < 2; ++a) {
    const Side& s = g.s[a];
    const int64_t cc = (int64_t)s.cols * s.cols;
    L.h[a] = o; o += herm_elems(s);
    L.g[a] = o; o += cc;
    L.gb[a] = o; o += cc;
    L.ev[a] = o; o += s.cols;
    L.r[a] = o; o += cc;
    L.rinv[a] = o; o += cc;
    L.y[a] = o; o += (int64_t)s.cols * PC;
    L.w[a] = o; o += (int64_t)PC * PC;
  }
  const int64_t m = (int64_t)g.s[0].cols * g.s[0].d, n = (int64_t)g.s[1].cols * g.s[1].d;
  for (int i = 0; i < 3; ++i) { L.theta[i] = o; o += m * n; }
  L.sig = o; o += n;
  L.order = o; o += n;
  for (int a = 0; a < 2; ++a) { L.tab[a] = o; o += (g.s[a].rows + PC) / 2 + 2; }  // int32 entries in slots of >= 8 bytes
  L.total = (o + 1) & ~(int64_t)1;  // keep every gate's work space 16-byte aligned for both element types
  return L;
}

// Leading dimension of a Jacobi operand (m rows, n columns) in shared memory: with GS lanes per pair (see jacobi_groups)
// consecutive columns must start GS doubles apart modulo the 16 double-wide banks, so that the groups of a warp hit
// disjoint banks (ld = m puts every column of a 64-row matrix on the same banks: 4-way conflicts on every access).
// Lanes per column pair.  Measured: narrow groups win as long as a lane's rows of the pair fit its register
// batch (rb rows per column: 8 Float64 / 4 ComplexF64); taller operands get wider groups so that the rotation does not
// have to re-load them -- the iteration is bound by shared-memory wavefronts.
__host__ __device__ inline int jacobi_gs(int n, int nprob, int m, int rb) {
  const int npairs = (n + 1) / 2 * nprob;
  int gs = npairs >= 16 ? 4 : (npairs >= 4 ? 8 : 16);
  (void)m;  // (measured, round 2: widening the groups of tall operands so that their rows stay in registers -- 8 lanes for
  (void)rb; // the 64 x 64 bond matrix -- changes nothing for Float64 and costs ComplexF64 15 %)
  return gs;
}
__host__ __device__ inline int jacobi_ld(int m, int n, int nprob = 1, int rb = 0) {
  int ld = m;
  while (ld % 16 != 8) ++ld;  // sized for the widest rule on the host; the kernel only needs callers and callee to agree
#ifdef __CUDA_ARCH__
  const int gs = jacobi_gs(n, nprob, m, rb);
  ld = m;
  if (gs < 16)
    while (ld % 16 != gs) ++ld;
#endif
  return ld;
}
template <typename T>
__host__ __device__ constexpr int jacobi_rb() { return Elem<T>::is_complex ? 4 : 8; }

// Can the fast path take this gate, and how much shared memory (elements of T) do its three kernels want?  0: not supported.
struct SmemNeed3 {
  int64_t sides, bond, fin;  // bp_apply3_sides / bp_apply3_bond / bp_apply3_final
  __host__ __device__ int64_t all() const { return sides > bond ? (sides > fin ? sides : fin) : (bond > fin ? bond : fin); }
};
__host__ __device__ inline SmemNeed3 smem_need3(const GateDesc& g, bool cplx) {
  SmemNeed3 nd = {0, 0, 0};
  if (g.nsides != 2) return nd;
  int64_t sides = 0, fin = 0, bond = 0;
  for (int a = 0; a < 2; ++a) {
    const Side& s = g.s[a];
    if (s.cols > PC || s.cols < 1) return nd;
    int64_t hsum = 0;
    for (int i = 0; i < s.z; ++i)
      if (i != s.bond_slot) {
        if (s.dim[i] > MAXDIM) return nd;
        hsum += (int64_t)s.dim[i] * s.dim[i];
      }
    const int64_t cb = (!cplx && (s.cols % 2 == 0)) ? 2 : 1;
    int64_t dim0 = 1;
    for (int i = 0; i < s.z; ++i)
      if (i != s.bond_slot) { dim0 = s.dim[i]; break; }
    const int64_t absorb = cb * (s.rows + s.rows / dim0) + 2 + hsum;  // padded column batch + the messages
    const int64_t gram = 2 * (int64_t)TRG * PCP;               // A tile, T tile (the reduction re-uses them)
    const int64_t trf = cplx ? 128 : 256;
    const int64_t f = trf * PCP + (int64_t)PC * (PC + 4);      // A tile(s), W (padded rows in the tensor-pipe version)
    sides = sides > absorb ? sides : absorb;
    sides = sides > gram ? sides : gram;
    fin = fin > f ? fin : f;
  }
  // the cross-warp reduction of the Gram pass: 8 partial tiles of PC x (PC or PC/2) elements
  const int64_t red = 8 * (int64_t)PC * (cplx ? PC / 2 : PC);
  sides = sides > red ? sides : red;
  const int64_t m = (int64_t)g.s[0].cols * g.s[0].d, n = (int64_t)g.s[1].cols * g.s[1].d;
  const int64_t ldb = jacobi_ld((int)m, (int)n);
  bond = ldb * n;
  const int64_t gramf = 2 * ((int64_t)(PC + 16) * PC + (int64_t)PC * PC);  // rotated + unrotated Gram matrix, both sides
  bond = bond > gramf ? bond : gramf;
  nd.sides = (sides + 1) & ~(int64_t)1;
  nd.bond = (bond + 1) & ~(int64_t)1;
  nd.fin = (fin + 1) & ~(int64_t)1;
  return nd;
}
__host__ __device__ inline int64_t smem_need(const GateDesc& g, bool cplx) { return smem_need3(g, cplx).all(); }

// ---- canonical addressing of the matrix view: element (row, col) of side `sd` sits at rowaddr(row) + coladdr(col) --------
struct Walk {
  int64_t rstride[MAXZ];  // canonical strides (elements) of the external legs, slot order
  int32_t rdim[MAXZ];
  int32_t next;           // external legs
  int64_t bstride;        // canonical stride of the bond leg
  int32_t d;
};
__host__ __device__ inline Walk walk_of(const Side& sd) {
  Walk w;
  w.next = 0;
  w.bstride = 0;
  w.d = sd.d;
  int64_t stride = sd.d;
  for (int k = 0; k < sd.z; ++k) {
    if (k == sd.bond_slot) {
      w.bstride = stride;
    } else {
      w.rstride[w.next] = stride;
      w.rdim[w.next] = sd.dim[k];
      ++w.next;
    }
    stride *= sd.dim[k];
  }
  return w;
}
__host__ __device__ __forceinline__ int64_t rowaddr(const Walk& w, int64_t row) {
  int64_t a = 0;
  for (int k = 0; k < w.next; ++k) {
    a += (row % w.rdim[k]) * w.rstride[k];
    row /= w.rdim[k];
  }
  return a;
}
__host__ __device__ __forceinline__ int64_t coladdr(const Walk& w, int col) { return (col % w.d) + (col / w.d) * w.bstride; }

// Address tables of one side (global work space, L1 / L2 resident): the hot loops look addresses up instead of
// decomposing indices.  rowtab[row] = canonical offset of the row, coltab[c] = offset of the column (0 for c >= cols).
struct Tabs {
  const int32_t* row;
  const int32_t* col;
  bool col_fast;  // the columns (s, bond) are contiguous in the canonical layout (bond leg first); else (s, first row leg)
};
template <typename T>
__host__ __device__ __forceinline__ Tabs build_tabs(const Team tm, const Side& sd, const Walk& wk, T* slot) {
  int32_t* rt = reinterpret_cast<int32_t*>(slot);
  int32_t* ct = rt + sd.rows;
  for (int64_t r = tm.tid(); r < sd.rows; r += tm.nt()) rt[r] = (int32_t)rowaddr(wk, r);
  for (int c = tm.tid(); c < PC; c += tm.nt()) ct[c] = c < sd.cols ? (int32_t)coladdr(wk, c) : 0;
  tm.sync();
  Tabs t;
  t.row = rt;
  t.col = ct;
  t.col_fast = wk.next == 0 || wk.bstride < wk.rstride[0];
  return t;
}

// the tables build_tabs left in `slot` (another kernel of the same gate)
template <typename T>
__host__ __device__ __forceinline__ Tabs tabs_at(const Side& sd, const Walk& wk, T* slot) {
  Tabs t;
  t.row = reinterpret_cast<const int32_t*>(slot);
  t.col = t.row + sd.rows;
  t.col_fast = wk.next == 0 || wk.bstride < wk.rstride[0];
  return t;
}

// rows [row0, row0 + nr) of a scratch copy -> a tile of tr rows in shared memory, one contiguous piece per column group
template <typename T>
__host__ __device__ __forceinline__ void copy_tile_groups(const Team tm, T* dst, const T* src, int64_t row0, int nr, int64_t rows_e, int cols,
                                                          int tr) {
  constexpr bool CPLX = Elem<T>::is_complex;
  const int ncg = CPLX ? cols : (cols + 1) / 2, ts = tile_stride16(CPLX, tr);
#ifdef __CUDA_ARCH__
  const double2* __restrict__ s2 = reinterpret_cast<const double2*>(src);
  double2* __restrict__ d2 = reinterpret_cast<double2*>(dst);
  const int total = ncg * nr, nt = tm.nt(), tid = tm.tid();
  constexpr int U = 4;
  const int nfull = total / (U * nt) * (U * nt);
  for (int i0 = tid; i0 < nfull; i0 += U * nt) {  // no predicates: the staging registers stay registers
    double2 v[U];
    int o[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * nt, j = i / nr, k = i - j * nr;
      v[u] = s2[(int64_t)j * rows_e + row0 + k];
      o[u] = j * ts + k;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) d2[o[u]] = v[u];
  }
  for (int i = nfull + tid; i < total; i += nt) {
    const int j = i / nr, k = i - j * nr;
    d2[j * ts + k] = s2[(int64_t)j * rows_e + row0 + k];
  }
#else
  constexpr int CG = CPLX ? 1 : 2;
  for (int64_t i = tm.tid(); i < (int64_t)ncg * nr * CG; i += tm.nt()) {
    const int j = (int)(i / (nr * CG)), k = (int)(i % (nr * CG));
    dst[(int64_t)j * ts * CG + k] = src[((int64_t)j * rows_e + row0) * CG + k];
  }
#endif
  tm.sync();
}

// ---- messages: Hermitian part + positive-definiteness check ---------------------------------------------------------------
// One warp per message: right-looking Cholesky on a scratch copy; *bad is set when a pivot is not safely positive.
template <typename T>
__host__ __device__ __forceinline__ void message_check(const Team tm, const Side& sd, const T* msgs, T* H, T* scratch, int* bad) {
  using E = Elem<T>;
  const int L = tm.lanes();
  int64_t off = 0;
  int leg = 0;
  for (int i = 0; i < sd.z; ++i) {
    if (i == sd.bond_slot) continue;
    const int chi = sd.dim[i];
    const T* m = msgs + sd.in_msg[i];
    T* h = H + off;
    for (int e = tm.tid(); e < chi * chi; e += tm.nt()) {
      const int r = e % chi, c = e / chi;
      const T v = scal(E::add(m[r + c * chi], E::conj(m[c + r * chi])), 0.5);
      h[e] = v;
      scratch[off + e] = v;
    }
    off += (int64_t)chi * chi;
    ++leg;
  }
  tm.sync();
  off = 0;
  leg = 0;
  for (int i = 0; i < sd.z; ++i) {
    if (i == sd.bond_slot) continue;
    const int chi = sd.dim[i];
    if (leg % tm.nw == tm.wid) {
      T* a = scratch + off;  // lower triangle, column-major
      double dmax = 0.0;
      for (int j = 0; j < chi; ++j) {
        const double x = real_of(a[j + j * chi]);
        dmax = x > dmax ? x : dmax;
      }
      const double thr = PIVOT_MIN * dmax;
      bool ok = dmax > 0.0;
      for (int j = 0; j < chi && ok; ++j) {
        const double piv = real_of(a[j + j * chi]);
        if (!(piv > thr)) {
          // an EXACTLY zero row / column is a zero-padded link index (the tensor is zero there too, so the reference's
          // projector on the message's support acts as the identity on it): skip it; anything else declines the gate
          const T* h0 = H + off;  // the message itself (the scratch copy has been eliminated up to column j)
          bool zero = true;
          for (int r = 0; r < chi; ++r) zero = zero && E::is_zero(h0[r + j * chi]) && E::is_zero(h0[j + r * chi]);
          if (zero) continue;
          ok = false;
          break;
        }
        const double inv = 1.0 / sqrt(piv);
#ifdef __CUDA_ARCH__
        __syncwarp();
#endif
        for (int r = j + tm.lane; r < chi; r += L) a[r + j * chi] = scal(a[r + j * chi], inv);
#ifdef __CUDA_ARCH__
        __syncwarp();
#endif
        // trailing update: a[r, c] -= l[r] conj(l[c]) for j < c <= r
        const int nt = chi - j - 1;
        for (int e = tm.lane; e < nt * nt; e += L) {
          const int r = j + 1 + e % nt, c = j + 1 + e / nt;
          if (r >= c) a[r + c * chi] = sub(a[r + c * chi], E::mul(a[r + j * chi], E::conj(a[c + j * chi])));
        }
#ifdef __CUDA_ARCH__
        __syncwarp();
#endif
      }
      if (!ok && tm.lane == 0) BPX_FLAG_SET(bad);
    }
    off += (int64_t)chi * chi;
    ++leg;
  }
  tm.sync();
}

// division by a run-time constant that is usually a power of two (link dimensions): a shift when it is
struct FastDiv {
  int d, sh;
  __host__ __device__ explicit FastDiv(int dd) : d(dd), sh(-1) {
    if (dd > 0 && (dd & (dd - 1)) == 0) {
      sh = 0;
      while ((1 << sh) < dd) ++sh;
    }
  }
  __host__ __device__ __forceinline__ int div(int x) const { return sh >= 0 ? (x >> sh) : x / d; }
  __host__ __device__ __forceinline__ int mod(int x) const { return sh >= 0 ? (x & (d - 1)) : x % d; }
};

// ---- absorb: T[:, c] = (M_1 x M_2 x ..) A[:, c], CB columns at a time, in place in shared memory ---------------------------
// A thread owns whole fibres (the chi elements along one leg) of all CB columns: inputs to registers, outputs back to the
// same places -- no ping-pong buffer; every matrix element it loads (a broadcast) feeds CB FMAs.
template <typename T, int CB, int CHI>
__host__ __device__ __forceinline__ void absorb_leg_fixed(const Team tm, T* col, int rows, int prow, int st, int pad0, const T* x) {
  using E = Elem<T>;
  const int nf = rows / CHI;
  const int pst = st >= pad0 ? st + st / pad0 : st;  // stride of the leg in the padded column (see absorb_side)
  const FastDiv dst(st), dpad(pad0);
  for (int f = tm.tid(); f < nf; f += tm.nt()) {
    const int lo = dst.mod(f), hi = dst.div(f), base0 = hi * st * CHI + lo, base = base0 + dpad.div(base0);
    T v[CB][CHI];
#pragma unroll
    for (int b = 0; b < CB; ++b)
#pragma unroll
      for (int l = 0; l < CHI; ++l) v[b][l] = col[b * prow + base + l * pst];
    constexpr int GU = CHI >= 4 ? 4 : CHI;  // outputs per pass: GU * CB independent FMA chains (one chain per output is latency bound)
#pragma unroll 1
    for (int g0 = 0; g0 < CHI; g0 += GU) {
      T acc[GU][CB];
#pragma unroll
      for (int gu = 0; gu < GU; ++gu)
#pragma unroll
        for (int b = 0; b < CB; ++b) acc[gu][b] = E::zero();
      // x is Hermitian: x[g, l] = conj(x[l, g]), and x[l + CHI g] is contiguous in l (vector loads, one broadcast each)
#pragma unroll
      for (int l = 0; l < CHI; ++l) {
#pragma unroll
        for (int gu = 0; gu < GU; ++gu) {
          const T xv = E::conj(x[CHI * (g0 + gu) + l]);
#pragma unroll
          for (int b = 0; b < CB; ++b) acc[gu][b] = E::fma(xv, v[b][l], acc[gu][b]);
        }
      }
#pragma unroll
      for (int gu = 0; gu < GU; ++gu)
#pragma unroll
        for (int b = 0; b < CB; ++b) col[b * prow + base + (g0 + gu) * pst] = acc[gu][b];
    }
  }
  tm.sync();
}
template <typename T, int CB>
__host__ __device__ void absorb_leg_any(const Team tm, T* col, int rows, int prow, int st, int pad0, int chi, const T* x) {
  using E = Elem<T>;
  const int nf = rows / chi;
  const int pst = st >= pad0 ? st + st / pad0 : st;
  const FastDiv dst(st), dpad(pad0);
  for (int f = tm.tid(); f < nf; f += tm.nt()) {
    const int lo = dst.mod(f), hi = dst.div(f), base0 = hi * st * chi + lo, base = base0 + dpad.div(base0);
    for (int b = 0; b < CB; ++b) {
      T v[MAXDIM], o[MAXDIM];
      for (int l = 0; l < chi; ++l) v[l] = col[b * prow + base + l * pst];
      for (int g = 0; g < chi; ++g) {
        T acc = E::zero();
        for (int l = 0; l < chi; ++l) acc = E::fma(x[g + chi * l], v[l], acc);
        o[g] = acc;
      }
      for (int g = 0; g < chi; ++g) col[b * prow + base + g * pst] = o[g];
    }
  }
  tm.sync();
}
template <typename T, int CB>
__host__ __device__ void absorb_leg(const Team tm, T* col, int rows, int prow, int st, int pad0, int chi, const T* x) {
  switch (chi) {
    case 2: absorb_leg_fixed<T, CB, 2>(tm, col, rows, prow, st, pad0, x); break;
    case 4: absorb_leg_fixed<T, CB, 4>(tm, col, rows, prow, st, pad0, x); break;
    case 8: absorb_leg_fixed<T, CB, 8>(tm, col, rows, prow, st, pad0, x); break;
    case 16: absorb_leg_fixed<T, CB, 16>(tm, col, rows, prow, st, pad0, x); break;
    default: absorb_leg_any<T, CB>(tm, col, rows, prow, st, pad0, chi, x);
  }
}

// The gathered column is padded by one element after every rdim[0] rows (index r + r / rdim[0]): the fibres of the first leg
// then start an odd number of elements apart (no bank conflicts; unpadded, a warp's 32 fibres of 16 doubles share one
// bank), and every other leg keeps a uniform stride st + st / rdim[0].
template <typename T, int CB>
__host__ __device__ __forceinline__ void absorb_side(const Team tm, const Side& sd, const Walk& wk, const Tabs tb, const T* a, const T* H, T* aout,
                                     T* tout, T* smem, long long* stamps = nullptr) {
#ifdef __CUDA_ARCH__
#define BPX_ASTAMP(i) do { if (stamps && tm.tid() == 0 && c0 == 0) stamps[i] = clock64(); } while (0)
#else
#define BPX_ASTAMP(i) do { (void)stamps; } while (0)
#endif
  const int rows = (int)sd.rows, ncols = sd.cols;
  const int64_t rows_e = rows_even(sd.rows);
  const int32_t* __restrict__ rowt = tb.row;
  const int32_t* __restrict__ colt = tb.col;
  const int pad0 = wk.next > 0 ? wk.rdim[0] : 1;
  const int prow = rows + rows / pad0;
  const FastDiv dpad(pad0);
  T* col = smem;                             // CB * prow
  T* hs = smem + (((int64_t)CB * prow + 1) & ~(int64_t)1);  // the messages (16-byte aligned)
  const int64_t hn = herm_elems(sd);
  for (int64_t i = tm.tid(); i < hn; i += tm.nt()) hs[i] = H[i];
  tm.sync();
  for (int c0 = 0; c0 < ncols; c0 += CB) {
    constexpr int U = 16;
    const int total = rows * CB, nt = tm.nt(), tid = tm.tid();
    BPX_ASTAMP(0);
    int cofs[CB];
#pragma unroll
    for (int b = 0; b < CB; ++b) cofs[b] = colt[c0 + b];
    const int nfull = total / (U * nt) * (U * nt);
    for (int i0 = tid; i0 < nfull; i0 += U * nt) {  // no predicates: U independent loads in flight, staged in registers
      int ad[U];
      T v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * nt;
        ad[u] = rowt[i / CB] + cofs[i % CB];
      }
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = a[ad[u]];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * nt, b = i % CB, r = i / CB;
        col[b * prow + r + dpad.div(r)] = v[u];
        aout[gidx<T>(r, c0 + b, rows_e)] = v[u];  // the matrix view: the Gram and final passes read plain tiles
      }
    }
    for (int i = nfull + tid; i < total; i += nt) {
      const int b = i % CB, r = i / CB;
      const T v = a[rowt[r] + cofs[b]];
      col[b * prow + r + dpad.div(r)] = v;
      aout[gidx<T>(r, c0 + b, rows_e)] = v;
    }
    tm.sync();
    BPX_ASTAMP(1);
    int st = 1;
    int64_t off = 0;
    for (int k = 0; k < wk.next; ++k) {
      const int chi = wk.rdim[k];
      absorb_leg<T, CB>(tm, col, rows, prow, st, pad0, chi, hs + off);
      st *= chi;
      off += (int64_t)chi * chi;
      BPX_ASTAMP(2 + k);
    }
    for (int i = tm.tid(); i < total; i += nt) {
      const int b = i % CB, r = i / CB;
      tout[gidx<T>(r, c0 + b, rows_e)] = col[b * prow + r + dpad.div(r)];
    }
    tm.sync();
    BPX_ASTAMP(7);
  }
#undef BPX_ASTAMP
}

// ---- Gram pass: G[c', c] = sum_rows conj(A[row, c']) T[row, c] -------------------------------------------------------------
// Tiles of TRG rows, both operands as [row][PC] in shared memory.  A warp takes every nw-th row of the tile; lane (i, j) of
// an 8 x 4 lane grid accumulates the 4 x TJ block G[4 i .., TJ (j + 4 pass) ..]: per row 4 + TJ operand loads (broadcast
// within the lane groups) feed 4 TJ FMAs.  The nw partial tiles are summed through shared memory at the end.
template <typename T, int TJ>
__host__ __device__ __forceinline__ void gram_side(const Team tm, const Side& sd, const T* at, const T* tt, T* G, T* smem) {
  using E = Elem<T>;
  constexpr int NPASS = PC / (4 * TJ);
  T* sA = smem;
  T* sT = smem + (int64_t)TRG * PCP;
  const int L = tm.lanes();
  const int cols = sd.cols;
  const int64_t rows_all = sd.rows;
  for (int pass = 0; pass < NPASS; ++pass) {
    // host lanes (L = 1): one lane plays all 32 roles in turn, so the accumulators live in an array indexed by role
#ifdef __CUDA_ARCH__
    T acc[4][TJ];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < TJ; ++y) acc[x][y] = E::zero();
#else
    std::vector<T> hacc((size_t)32 * 4 * TJ, E::zero());
#endif
    for (int64_t row0 = 0; row0 < rows_all; row0 += TRG) {
      const int nr = (int)((rows_all - row0) < TRG ? (rows_all - row0) : TRG);
      copy_tile_groups<T>(tm, sA, at, row0, nr, rows_even(rows_all), cols, TRG);
      copy_tile_groups<T>(tm, sT, tt, row0, nr, rows_even(rows_all), cols, TRG);
#ifdef __CUDA_ARCH__
      const int li = tm.lane >> 2, lj = tm.lane & 3;
      for (int r = tm.wid; r < nr; r += tm.nw) {
        T av[4], tv[TJ];
        if constexpr (!E::is_complex) {  // column pairs (c even, c + 1) are 16 contiguous, aligned bytes
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            const double2 q = *reinterpret_cast<const double2*>(sA + tidx<T>(r, 4 * li + 2 * x, TRG));
            av[2 * x] = q.x;
            av[2 * x + 1] = q.y;
          }
#pragma unroll
          for (int y = 0; y < TJ / 2; ++y) {
            const double2 q = *reinterpret_cast<const double2*>(sT + tidx<T>(r, TJ * (lj + 4 * pass) + 2 * y, TRG));
            tv[2 * y] = q.x;
            tv[2 * y + 1] = q.y;
          }
        } else {
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const double2 q = *reinterpret_cast<const double2*>(sA + tidx<T>(r, 4 * li + x, TRG));
            av[x] = E::conj(*reinterpret_cast<const T*>(&q));
          }
#pragma unroll
          for (int y = 0; y < TJ; ++y) {
            const double2 q = *reinterpret_cast<const double2*>(sT + tidx<T>(r, TJ * (lj + 4 * pass) + y, TRG));
            tv[y] = *reinterpret_cast<const T*>(&q);
          }
        }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < TJ; ++y) acc[x][y] = E::fma(av[x], tv[y], acc[x][y]);
      }
#else
      for (int role = 0; role < 32; ++role) {
        const int li = role >> 2, lj = role & 3;
        for (int r = tm.wid; r < nr; r += tm.nw)
          for (int x = 0; x < 4; ++x)
            for (int y = 0; y < TJ; ++y) {
              T& o = hacc[(size_t)(role * 4 + x) * TJ + y];
              o = E::fma(E::conj(sA[tidx<T>(r, 4 * li + x, TRG)]), sT[tidx<T>(r, TJ * (lj + 4 * pass) + y, TRG)], o);
            }
      }
      (void)L;
#endif
      tm.sync();
    }
    // cross-warp reduction: red[w][c'][y-block]
    constexpr int PW = 4 * TJ;  // columns of G covered by one pass
    T* red = smem;
#ifdef __CUDA_ARCH__
    {
      const int li = tm.lane >> 2, lj = tm.lane & 3;
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < TJ; ++y) red[((int64_t)tm.wid * PC + 4 * li + x) * PW + TJ * lj + y] = acc[x][y];
    }
#else
    for (int role = 0; role < 32; ++role) {
      const int li = role >> 2, lj = role & 3;
      for (int x = 0; x < 4; ++x)
        for (int y = 0; y < TJ; ++y) red[((int64_t)tm.wid * PC + 4 * li + x) * PW + TJ * lj + y] = hacc[(size_t)(role * 4 + x) * TJ + y];
    }
#endif
    tm.sync();
    for (int e = tm.tid(); e < PC * PW; e += tm.nt()) {
      const int cp = e / PW, cc = pass * PW + e % PW;
      T s = E::zero();
      for (int w = 0; w < tm.nw; ++w) s = E::add(s, red[(int64_t)w * PC * PW + e]);
      if (cp < cols && cc < cols) G[cp + cols * cc] = s;
    }
    tm.sync();
  }
}

#ifdef __CUDACC__
// ---- Gram pass on the FP64 tensor pipe (Float64, device only) --------------------------------------------------------------
// G = A^H T as DMMA.8x8x4: M = c' (4 blocks of 8), N = c (4 blocks of 8), K = rows.  A warp takes every nw-th group of 4
// rows of a tile and keeps all 16 accumulator tiles (32 doubles); per group 4 + 4 fragment loads (one element per lane each:
// A^H[c' = 8 mb + lane / 4][row = lane % 4], T[row = lane % 4][c = 8 nb + lane / 4]) feed 16 DMMAs = 4096 MACs -- against
// 12 operand loads per 32 FMAs = 1024 MACs per warp instruction group of the FMA version.  Partial tiles of the warps are
// summed through shared memory exactly like there.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// The tiles arrive by TMA bulk copies (they are contiguous row ranges of the scratch copies), two stages of TRG / 2 rows, one
// mbarrier per stage: the loads of the next tile overlap the DMMAs of this one (the register-staged synchronous copies of the
// FMA version left the pass waiting on L2 / DRAM round trips most of the time: 26 k clocks per 70 KB tile pair).
__device__ __forceinline__ uint32_t smem_u32_3(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct GramPipe {
  uint64_t* bar;     // two mbarriers in shared memory, initialised once per kernel (gram_pipe_init)
  uint32_t used[2];  // completed phases per stage (the same in every thread)
};
__device__ __forceinline__ void gram_pipe_init(GramPipe& gp, uint64_t* bar) {
  gp.bar = bar;
  gp.used[0] = gp.used[1] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32_3(bar + i)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
}
__device__ __forceinline__ void gram_side_mma(const Team tm, const Side& sd, const double* at, const double* tt, double* G, double* smem,
                                              GramPipe& gp) {
  static_assert(PC == 32, "4 x 4 blocks of 8 x 8");
  constexpr int TR = TRG / 2;
  const int cols = sd.cols;
  const int64_t rows_all = sd.rows;
  const int ntile = (int)((rows_all + TR - 1) / TR);
  const int g = tm.lane >> 2, t = tm.lane & 3;
  const int ncg = (cols + 1) / 2;                 // column groups (16-byte vectors per row)
  const int64_t rows_e = rows_even(rows_all);
  constexpr int TS16 = TR + 4;                    // tile_stride16(false, TR)
  auto issue = [&](int i) {  // warp 0: both tiles of row block i into stage i & 1, one bulk copy per lane (group, A | T)
    const int s = i & 1;
    const int64_t row0 = (int64_t)i * TR;
    const uint32_t nr = (uint32_t)((rows_all - row0) < TR ? (rows_all - row0) : TR);
    const uint32_t bar = smem_u32_3(gp.bar + s), dst = smem_u32_3(smem + (int64_t)s * 2 * TR * PCP);
    if (tm.lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(2 * ncg * nr * 16) : "memory");
    __syncwarp();
    const int j = tm.lane & 15;
    if (j < ncg) {
      const double* src = ((tm.lane < 16) ? at : tt) + (((int64_t)j * rows_e + row0) << 1);
      const uint32_t d = dst + (uint32_t)((tm.lane < 16 ? 0 : TR * PCP) * sizeof(double)) + (uint32_t)(j * TS16 * 16);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d), "l"(src),
                   "r"(nr * 16), "r"(bar)
                   : "memory");
    }
  };
  tm.sync();  // (the scratch copies were written by this CTA's ordinary stores: visible after the barrier; shared memory is free)
  if (tm.wid == 0) {
    asm volatile("fence.proxy.async;\n" ::: "memory");  // generic-proxy writes (global and shared) before the async-proxy copies
    issue(0);
    if (ntile > 1) issue(1);
  }
  double acc[4][4][2];
#pragma unroll
  for (int mb = 0; mb < 4; ++mb)
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
  for (int i = 0; i < ntile; ++i) {
    const int s = i & 1;
    const int nr = (int)((rows_all - (int64_t)i * TR) < TR ? (rows_all - (int64_t)i * TR) : TR);
    {
      const uint32_t bar = smem_u32_3(gp.bar + s), parity = gp.used[s] & 1;
      uint32_t ok;
      do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
      } while (!ok);
      gp.used[s]++;
    }
    const double* sA = smem + (int64_t)s * 2 * TR * PCP;
    const double* sT = sA + TR * PCP;
    // (columns >= cols of the tiles are uninitialised: their products land in rows / columns of G that are never read)
    for (int r0 = 4 * tm.wid; r0 < nr; r0 += 4 * tm.nw) {
      const bool ok = r0 + t < nr;
      const int o = tidx<double>(r0 + t, g, TR);  // (column 8 b + g: 4 b groups further)
      double fa[4], fb[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        fa[b] = ok ? sA[o + 8 * b * TS16] : 0.0;
        fb[b] = ok ? sT[o + 8 * b * TS16] : 0.0;
      }
#pragma unroll
      for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], fa[mb], fb[nb]);
    }
    tm.sync();  // every warp is done with stage s
    if (tm.wid == 0 && i + 2 < ntile) issue(i + 2);
  }
  double* red = smem;  // red[w][c'][c]
#pragma unroll
  for (int mb = 0; mb < 4; ++mb)
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      double* o = red + ((int64_t)tm.wid * PC + 8 * mb + g) * PC + 8 * nb + 2 * t;
      o[0] = acc[mb][nb][0];
      o[1] = acc[mb][nb][1];
    }
  tm.sync();
  for (int e = tm.tid(); e < PC * PC; e += tm.nt()) {
    const int cp = e / PC, cc = e % PC;
    double s = 0.0;
    for (int w = 0; w < tm.nw; ++w) s += red[(int64_t)w * PC * PC + e];
    if (cp < cols && cc < cols) G[cp + cols * cc] = s;
  }
  tm.sync();
}
#endif

// 1 / sqrt(x) and 1 / x to full double accuracy from the hardware's 20-bit approximations + two Newton steps: a short
// dependent chain (the IEEE sqrt / division sequences cost several hundred cycles of latency each, and the Jacobi steps
// below are pure latency).  Outside the safe exponent range the exact functions are used.
// GUARD = false: the caller knows that x is in the safe range (the exact fallbacks are function calls: every register live
// across them is spilled around the call)
template <int STEPS = 2, bool GUARD = true>
__host__ __device__ __forceinline__ double rsqrt_d(double x) {
#ifdef __CUDA_ARCH__
  if (GUARD && !(x > 1e-280 && x < 1e280)) return rsqrt(x);
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
  for (int it = 0; it < STEPS; ++it) {
    const double e = fma(-x * y, y, 1.0);  // 1 - x y^2
    y = fma(0.5 * y, e, fma(0.375 * y, e * e, y));  // second-order step: y (1 + e/2 + 3 e^2 / 8)
  }
  return y;
#else
  return 1.0 / sqrt(x);
#endif
}
template <int STEPS = 2, bool GUARD = true>
__host__ __device__ __forceinline__ double rcp_d(double x) {
#ifdef __CUDA_ARCH__
  if (GUARD && !(fabs(x) > 1e-280 && fabs(x) < 1e280)) return 1.0 / x;
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
  for (int it = 0; it < STEPS; ++it) y = fma(y, fma(-x, y, 1.0), y);
  return y;
#else
  return 1.0 / x;
#endif
}

// ---- one-sided Jacobi without V, pairs on sub-warp lane groups ---------------------------------------------------------------
// B (m x n, leading dimension ld, shared memory) is rotated until its columns are mutually orthogonal; *not_converged is
// set when the last sweep of the budget still rotated.  A group of GS lanes owns one column pair of the current
// round-robin step; a lane walks its rows of the pair in batches of RB rows held in registers: when one batch covers them
// (m <= GS RB) they stay there between the inner products and the rotation, otherwise the rotation re-loads them.
// `nprob` (1 or 2) independent problems of the same shape, `pstride` elements apart, share the steps and the barriers: the
// two Gram matrices of a gate are diagonalised together.  A problem that has converged only fails the rotation test in the
// remaining sweeps, so the result is bit-identical to running the problems one after the other.
// NOT inlined, on purpose: measured on the B200 (round 2, profiles/r2a*_apply*), this form -- a called function with plain
// pointers -- beats every "cleaner" variant tried: inlined into the bond kernel (+35 % kernel time: the kernel's other
// phases share its register allocation), 32-bit shared addresses through inline PTX with a predicate-free path for full
// batches (+60 %), unordered pairs for conflict-free banks (no change), wider groups for tall operands (no change for
// Float64, -15 % for ComplexF64), and -- keeping this very form -- instances without the bounds predicates and the zero fill
// for operands whose rows fill the batches, idle groups reading a zero column (18 % fewer instructions, +12 % kernel time:
// profiles/r2bn_jacobi_full_instances.txt).
template <typename T, int GS, int RB>
__host__ __device__ __noinline__ void jacobi_groups_t(const Team tm, T* B, int m, int n, int ld, int nprob, int pstride, int* flag,
                                                      int* not_converged, long long* sweeps_out) {
  using E = Elem<T>;
  const int L = tm.lanes();
  const int np = (n + 1) & ~1, npairs = np / 2, npt = npairs * nprob;
  const int gpw = L / GS;                // groups per warp
  const int sl = tm.lane % GS, grp = tm.lane / GS;
  const int rpl = (m + GS - 1) / GS;     // rows per lane
  const bool single = rpl <= RB;
  const double tol2 = (double)m * EPS * EPS;
  double zero2[2] = {0.0, 0.0};
  for (int k = 0; k < nprob; ++k) {
    double fro2 = 0.0;
    const T* Bk = B + (int64_t)k * pstride;
    for (int j = 0; j < n; ++j)
      for (int r = tm.lane; r < m; r += L) fro2 += E::abs2(Bk[r + ld * j]);
    fro2 = tm.sum(fro2);
    zero2[k] = (double)n * n * EPS * EPS * fro2;
  }
  const double zero2a = zero2[0], zero2b = zero2[1];
  tm.sync();
  int f = 1, nsweeps = 0;
  for (int sweep = 0; sweep < MAX_JACOBI_SWEEPS && f; ++sweep) {
    ++nsweeps;
    if (tm.tid() == 0) *flag = 0;
    tm.sync();
    for (int step = 0; step < np - 1; ++step) {
      for (int base = tm.wid * gpw; base < npt; base += tm.nw * gpw) {  // warp-uniform bound: shuffles stay converged
        const int gi = base + grp;
        const bool second = gi >= npairs;
        const int idx = second ? gi - npairs : gi;
        int p = 0, q = 1;
        bool active = gi < npt;
        if (active) {
          if (idx == 0) {
            p = np - 1;
            q = step;
          } else {
            p = step + idx;
            if (p >= np - 1) p -= np - 1;
            q = step - idx;
            if (q < 0) q += np - 1;
          }
          if (p > q) { const int t = p; p = q; q = t; }
          if (q >= n) { active = false; p = 0; q = 1; }
        }
        T* bp = B + (second ? pstride : 0) + p * ld + sl;
        T* bq = B + (second ? pstride : 0) + q * ld + sl;
        double a = 0.0, b = 0.0, a1 = 0.0, b1 = 0.0;  // two partial sums per inner product: half the dependent-chain length
        T g = E::zero(), g1 = E::zero();
        T xr[RB], yr[RB];
        // (groups without a pair this step must not touch the matrix: another group owns columns 0, 1)
        for (int j0 = 0; j0 < rpl; j0 += RB) {
#pragma unroll
          for (int j = 0; j < RB; ++j) {
            const bool ok = active && sl + (j0 + j) * GS < m;
            xr[j] = ok ? bp[(j0 + j) * GS] : E::zero();
            yr[j] = ok ? bq[(j0 + j) * GS] : E::zero();
          }
#pragma unroll
          for (int j = 0; j < RB; ++j) {
            if (j & 1) {
              a1 += E::abs2(xr[j]);
              b1 += E::abs2(yr[j]);
              g1 = E::fma(E::conj(xr[j]), yr[j], g1);
            } else {
              a += E::abs2(xr[j]);
              b += E::abs2(yr[j]);
              g = E::fma(E::conj(xr[j]), yr[j], g);
            }
          }
        }
        a += a1;
        b += b1;
        g = E::add(g, g1);
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int o = GS >> 1; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
          g = E::add(g, E::shfl_xor(g, o));
        }
#endif
        const double g2 = E::abs2(g), z2 = second ? zero2b : zero2a;
        if (!active || !(g2 > tol2 * a * b) || !(a > z2) || !(b > z2)) continue;  // group-uniform, no shuffles below
        // rotation parameters on short dependent chains (rsqrt_d / rcp_d)
        // Any angle gives an exactly unitary rotation as long as c = 1 / sqrt(1 + t^2), s = c t and the phase are accurate:
        // the angle itself (zeta, t) is computed with one Newton step (~1e-12, no effect on the quadratic convergence)
        const double inv_ga = rsqrt_d<E::is_complex ? 2 : 1>(g2);
        T ph;
        if constexpr (E::is_complex) ph = scal(g, inv_ga);
        else ph = from_real<T>(real_of(g) >= 0.0 ? 1.0 : -1.0);
        const double zeta = 0.5 * (b - a) * inv_ga, az = fabs(zeta);
        double t;
        if (az < 1e100) {
          const double w1 = fma(zeta, zeta, 1.0);
          t = rcp_d<1>(az + w1 * rsqrt_d<1>(w1));  // 1 / (|zeta| + sqrt(1 + zeta^2))
        } else {
          t = 0.5 * rcp_d<1>(az);
        }
        if (zeta < 0.0) t = -t;
        const double c = rsqrt_d<2>(fma(t, t, 1.0)), s = c * t;
        const T sph = scal(ph, s), scph = scal(E::conj(ph), s);
        for (int j0 = 0; j0 < rpl; j0 += RB) {
          if (!single) {
#pragma unroll
            for (int j = 0; j < RB; ++j) {
              const bool ok = sl + (j0 + j) * GS < m;
              xr[j] = ok ? bp[(j0 + j) * GS] : E::zero();
              yr[j] = ok ? bq[(j0 + j) * GS] : E::zero();
            }
          }
#pragma unroll
          for (int j = 0; j < RB; ++j) {  // all outputs first, then the stores: independent temporaries
            const T x = xr[j], y = yr[j];
            xr[j] = sub(scal(x, c), E::mul(y, scph));
            yr[j] = E::add(E::mul(x, sph), scal(y, c));
          }
#pragma unroll
          for (int j = 0; j < RB; ++j) {
            if (sl + (j0 + j) * GS < m) {
              bp[(j0 + j) * GS] = xr[j];
              bq[(j0 + j) * GS] = yr[j];
            }
          }
        }
        if (sl == 0) BPX_FLAG_SET(flag);
      }
      tm.sync();
    }
    f = *flag;
    tm.sync();
  }
  if (f && tm.tid() == 0) BPX_FLAG_SET(not_converged);
  if (sweeps_out && tm.tid() == 0) *sweeps_out = nsweeps;
  tm.sync();
}

template <typename T>
__host__ __device__ void jacobi_groups(const Team tm, T* B, int m, int n, int ld, int* flag, int* not_converged,
                                       long long* sweeps_out = nullptr, int nprob = 1, int pstride = 0) {
  if (n < 2) return;
#ifdef __CUDA_ARCH__
  constexpr int RB = jacobi_rb<T>();
  switch (jacobi_gs(n, nprob, m, RB)) {
    case 4: jacobi_groups_t<T, 4, RB>(tm, B, m, n, ld, nprob, pstride, flag, not_converged, sweeps_out); break;
    case 8: jacobi_groups_t<T, 8, RB>(tm, B, m, n, ld, nprob, pstride, flag, not_converged, sweeps_out); break;
    default: jacobi_groups_t<T, 16, RB>(tm, B, m, n, ld, nprob, pstride, flag, not_converged, sweeps_out); break;
  }
#else
  jacobi_groups_t<T, 1, 4>(tm, B, m, n, ld, nprob, pstride, flag, not_converged, sweeps_out);  // host lanes: one lane per pair
#endif
}

// Column order for a Jacobi operand: position of every column when they are sorted by decreasing norm (de Rijk's
// ordering: graded matrices converge in fewer sweeps when the large columns come first).  src: m x n, leading dimension m,
// global or shared; nrm (n doubles) and pos (n ints) are scratch / result.  Any column order is a valid input of the
// iteration -- the callers only use order-independent results.
template <typename T>
__host__ __device__ void column_order(const Team tm, const T* src, int m, int n, double* nrm, int32_t* pos) {
  using E = Elem<T>;
  for (int j = tm.tid(); j < n; j += tm.nt()) {
    double a = 0.0;
    for (int r = 0; r < m; ++r) a += E::abs2(src[r + (int64_t)m * j]);
    nrm[j] = a;
  }
  tm.sync();
  for (int j = tm.tid(); j < n; j += tm.nt()) {
    int before = 0;
    const double a = nrm[j];
    for (int i = 0; i < n; ++i) before += (nrm[i] > a || (nrm[i] == a && i < j)) ? 1 : 0;
    pos[j] = before;
  }
  tm.sync();
}

// ---- eigen-decomposition of the Gram matrix -> R, R^+ ------------------------------------------------------------------------
// G Hermitian (cols x cols, global).  Rotating the columns of G: G V = V diag(lam), so column j of the rotated copy is
// lam_j v_j: |lam_j| = |b_j|, sign from Re(b_j^H G b_j) (a slightly indefinite G).  *bad: conditioning / convergence.
// Split in two so that the two sides of a gate share ONE iteration (jacobi_groups, nprob = 2): `gram_factor_load` puts G
// (columns sorted by norm) and an unrotated copy into shared memory, `gram_factor_finish` turns the rotated copy into R, R^+.
template <typename T>
__host__ __device__ __forceinline__ void gram_factor_load(const Team tm, int cols, const T* G, T* Gb, double* ev, T* sb, T* sg, int ld) {
  int32_t* pos = reinterpret_cast<int32_t*>(Gb);  // (Gb is written after the iteration)
  column_order<T>(tm, G, cols, cols, ev, pos);
  for (int i = tm.tid(); i < cols * cols; i += tm.nt()) {
    const T v = G[i];
    sb[(i % cols) + ld * pos[i / cols]] = v;
    sg[i] = v;
  }
  tm.sync();
}
template <typename T>
__host__ __device__ __forceinline__ void gram_factor_finish(const Team tm, int cols, int rank_max, T* Gb, double* ev, T* R, T* Rinv, const T* sb,
                                                         const T* sg, int ld, int* bad) {
  using E = Elem<T>;
  for (int i = tm.tid(); i < cols * cols; i += tm.nt()) Gb[i] = sb[(i % cols) + ld * (i / cols)];
  tm.sync();
  for (int j = tm.tid(); j < cols; j += tm.nt()) {
    const T* b = sb + (int64_t)ld * j;
    double nb = 0.0;
    for (int r = 0; r < cols; ++r) nb += E::abs2(b[r]);
    T acc = E::zero();  // b^H G b,  (G b)[r] = sum_c G[r, c] b[c]
    for (int r = 0; r < cols; ++r) {
      T gb = E::zero();
      for (int c = 0; c < cols; ++c) gb = E::fma(sg[r + cols * c], b[c], gb);
      acc = E::fma(E::conj(b[r]), gb, acc);
    }
    // |lam_j| = |b_j| (b_j = lam_j v_j); the Rayleigh quotient only supplies the sign -- for a column that the iteration
    // left at rounding level its VALUE would be an average of all eigenvalues, not small
    ev[j] = real_of(acc) >= 0.0 ? sqrt(nb) : -sqrt(nb);
  }
  tm.sync();
  double dmax = 0.0, dmin = INFINITY;
  for (int j = 0; j < cols; ++j) dmax = ev[j] > dmax ? ev[j] : dmax;
  const double cut = EPS * cols * dmax;
  // rank bound: G = P^H P with P rows x cols has at most `rank_max` non-zero eigenvalues; whatever the iteration left in
  // the other directions is rounding noise (it may exceed `cut` by a small factor), so only the largest rank_max count
  double evl[PC];  // a private copy: the loop below zeroes entries of ev that other threads still rank against
  for (int i = 0; i < cols; ++i) evl[i] = ev[i];
  tm.sync();
  for (int j = tm.tid(); j < cols; j += tm.nt()) {
    int before = 0;
    for (int i = 0; i < cols; ++i) before += (evl[i] > evl[j] || (evl[i] == evl[j] && i < j)) ? 1 : 0;
    if (before >= rank_max || !(evl[j] > cut)) ev[j] = 0.0;
  }
  tm.sync();
  for (int j = 0; j < cols; ++j)
    if (ev[j] > 0.0 && ev[j] < dmin) dmin = ev[j];
  if (!(dmax > 0.0) || dmin < COND_MIN * dmax) {
    if (tm.tid() == 0) BPX_FLAG_SET(bad);
  }
  for (int i = tm.tid(); i < cols * cols; i += tm.nt()) {
    const int q = i % cols, c = i / cols;  // R[q, c] = sqrt(lam_q) conj(v_q[c]),  v_q = b_q / |b_q|
    const double lam = ev[q];
    T r = E::zero(), ri = E::zero();
    if (lam > 0.0) {
      const T* b = sb + (int64_t)ld * q;
      double nb = 0.0;
      for (int rr = 0; rr < cols; ++rr) nb += E::abs2(b[rr]);
      const double inv = 1.0 / sqrt(nb);
      const T v = scal(b[c], inv);
      r = scal(E::conj(v), sqrt(lam));
      ri = scal(v, 1.0 / sqrt(lam));
    }
    R[q + cols * c] = r;
    Rinv[c + cols * q] = ri;
  }
  tm.sync();
}

// ---- final pass: A'[row, c'] = sum_c A[row, c] W[c, c'], A from its row-major scratch copy, A' into the canonical tensor ----
// Tiles of TRF rows in shared memory (tidx() layout).  A thread owns RT rows x NO outputs: per
// column pair RT 16-byte operand loads (rows 272 bytes apart: conflict free) and NO 16-byte broadcast loads of W feed 2 RT NO
// FMAs.
template <typename T, int TRF, int RT, int NO>
__host__ __device__ __forceinline__ void final_side(const Team tm, const Side& sd, const Tabs tb, const T* at, T* a, const T* W, T* smem) {
  using E = Elem<T>;
  T* sA = smem;                                  // tidx(r, c, TRF)
  T* sW = smem + (int64_t)TRF * PCP;             // [c][PC]
  for (int i = tm.tid(); i < PC * PC; i += tm.nt()) sW[i] = W[i];
  const int cols = sd.cols;
  const int64_t rows_all = sd.rows;
  const int32_t* __restrict__ rowt = tb.row;
  const int32_t* __restrict__ colt = tb.col;
  constexpr int RG = TRF / RT;                   // row groups: thread rows rg, rg + RG, ..
  constexpr int OG = PC / NO;                    // output groups
  for (int64_t row0 = 0; row0 < rows_all; row0 += TRF) {
    const int nr = (int)((rows_all - row0) < TRF ? (rows_all - row0) : TRF);
    copy_tile_groups<T>(tm, sA, at, row0, nr, rows_even(rows_all), cols, TRF);
    for (int item = tm.tid(); item < RG * OG; item += tm.nt()) {
      const int rg = item % RG, og = item / RG;
      if (og * NO >= cols) continue;
      T acc[RT][NO];
#pragma unroll
      for (int x = 0; x < RT; ++x)
#pragma unroll
        for (int y = 0; y < NO; ++y) acc[x][y] = E::zero();
      for (int c = 0; c < cols; c += 2) {
        const bool two = c + 1 < cols;  // (columns >= cols of the scratch copy are not initialised)
        T a0[RT], a1[RT];
#pragma unroll
        for (int x = 0; x < RT; ++x) {
          const T* pa = sA + tidx<T>(rg + x * RG, c, TRF);  // (c is even)
          const T* pb = sA + tidx<T>(rg + x * RG, c + 1, TRF);
#ifdef __CUDA_ARCH__
          if constexpr (!E::is_complex) {
            const double2 q = *reinterpret_cast<const double2*>(pa);  // the pair (c, c + 1): 16 contiguous, aligned bytes
            a0[x] = q.x;
            a1[x] = two ? q.y : 0.0;
          } else {
            const double2 q0 = *reinterpret_cast<const double2*>(pa);
            a0[x] = *reinterpret_cast<const T*>(&q0);
            a1[x] = E::zero();
            if (two) {
              const double2 q1 = *reinterpret_cast<const double2*>(pb);
              a1[x] = *reinterpret_cast<const T*>(&q1);
            }
          }
#else
          a0[x] = pa[0];
          a1[x] = two ? pb[0] : E::zero();
#endif
        }
        const T* pw = sW + c * PC + og * NO;
        T w0[NO], w1[NO];
#ifdef __CUDA_ARCH__
        {  // 16-byte aligned rows of W: vector loads, one broadcast each
          const double2* p0 = reinterpret_cast<const double2*>(pw);
          const double2* p1 = reinterpret_cast<const double2*>(pw + PC);
          double2* v0 = reinterpret_cast<double2*>(w0);
          double2* v1 = reinterpret_cast<double2*>(w1);
#pragma unroll
          for (int y = 0; y < (int)(NO * sizeof(T) / 16); ++y) {
            v0[y] = p0[y];
            v1[y] = p1[y];
          }
        }
#else
        for (int y = 0; y < NO; ++y) {
          w0[y] = pw[y];
          w1[y] = pw[PC + y];
        }
#endif
#pragma unroll
        for (int y = 0; y < NO; ++y)
#pragma unroll
          for (int x = 0; x < RT; ++x) acc[x][y] = E::fma(a1[x], w1[y], E::fma(a0[x], w0[y], acc[x][y]));
      }
#pragma unroll
      for (int x = 0; x < RT; ++x) {
        const int r = rg + x * RG;
        if (r >= nr) continue;
        const int64_t ra = rowt[row0 + r];
#pragma unroll
        for (int y = 0; y < NO; ++y) {
          const int cp = og * NO + y;
          if (cp < cols) a[ra + colt[cp]] = acc[x][y];
        }
      }
    }
    tm.sync();
  }
}

#ifdef __CUDACC__
// ---- final pass on the FP64 tensor pipe (Float64, device only) -------------------------------------------------------------
// A'[rows of the tile, :] = at_tile (128 x 32) W (32 x 32) as DMMA.8x8x4: a warp owns 16 rows (two M blocks), walks K = c in 8
// steps and N = c' in two halves of two blocks; the tiles arrive by TMA bulk copies like in gram_side_mma (two stages of
// 128 rows); W sits in shared memory with a leading dimension of PC + 4 (conflict-free B fragments).  A lane's accumulator
// pair is (c' = 2 t, 2 t + 1) of one row: the two physical components of one bond index, contiguous in the canonical tensor
// when d is even -- one 16-byte store.
constexpr int PCW = PC + 4;
__device__ __forceinline__ void final_side_mma(const Team tm, const Side& sd, const Tabs tb, const double* at, double* a, const double* W,
                                               double* smem, GramPipe& gp) {
  constexpr int TR = 128;
  double* sW = smem + (int64_t)2 * TR * PCP;  // [c][PCW]
  const int cols = sd.cols;
  const int64_t rows_all = sd.rows;
  const int ntile = (int)((rows_all + TR - 1) / TR);
  const int g = tm.lane >> 2, t = tm.lane & 3;
  const int32_t* __restrict__ rowt = tb.row;
  const int32_t* __restrict__ colt = tb.col;
  const int ncg = (cols + 1) / 2;
  const int64_t rows_e = rows_even(rows_all);
  constexpr int TS16 = TR + 4;
  auto issue = [&](int i) {  // warp 0: row block i into stage i & 1, one bulk copy per lane (column group)
    const int s = i & 1;
    const int64_t row0 = (int64_t)i * TR;
    const uint32_t nr = (uint32_t)((rows_all - row0) < TR ? (rows_all - row0) : TR);
    const uint32_t bar = smem_u32_3(gp.bar + s), dst = smem_u32_3(smem + (int64_t)s * TR * PCP);
    if (tm.lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(ncg * nr * 16) : "memory");
    __syncwarp();
    if (tm.lane < ncg)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                       dst + (uint32_t)(tm.lane * TS16 * 16)),
                   "l"(at + (((int64_t)tm.lane * rows_e + row0) << 1)), "r"(nr * 16), "r"(bar)
                   : "memory");
  };
  tm.sync();
  if (tm.wid == 0) {
    asm volatile("fence.proxy.async;\n" ::: "memory");
    issue(0);
    if (ntile > 1) issue(1);
  }
  for (int i = tm.tid(); i < PC * PC; i += tm.nt()) sW[(i / PC) * PCW + (i % PC)] = W[i];
  tm.sync();
  // the pair (c', c' + 1), c' even, is one 16-byte piece of the canonical tensor?
  const bool pair16 = (sd.d % 2 == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0);
  for (int i = 0; i < ntile; ++i) {
    const int s = i & 1;
    const int64_t row0 = (int64_t)i * TR;
    const int nr = (int)((rows_all - row0) < TR ? (rows_all - row0) : TR);
    {
      const uint32_t bar = smem_u32_3(gp.bar + s), parity = gp.used[s] & 1;
      uint32_t ok;
      do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
      } while (!ok);
      gp.used[s]++;
    }
    const double* sA = smem + (int64_t)s * TR * PCP;
    const int rb = 16 * tm.wid;  // this warp's rows of the tile (8 warps x 16 rows)
    if (rb < nr) {
      const bool ok0 = rb + g < nr, ok1 = rb + 8 + g < nr;
      const int o0 = tidx<double>(rb + g, t, TR), o1 = tidx<double>(rb + 8 + g, t, TR);  // (column 4 ks + t: 2 ks groups further)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        double acc[2][2][2];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          // (columns >= cols of the scratch copy are not initialised: W's rows there are zero, so mask them)
          const bool kin = 4 * ks + t < cols;
          const double fa0 = (ok0 && kin) ? sA[o0 + 4 * ks * TS16] : 0.0, fa1 = (ok1 && kin) ? sA[o1 + 4 * ks * TS16] : 0.0;
          const double fb0 = sW[(4 * ks + t) * PCW + 16 * half + g], fb1 = sW[(4 * ks + t) * PCW + 16 * half + 8 + g];
          dmma884(acc[0][0][0], acc[0][0][1], fa0, fb0);
          dmma884(acc[0][1][0], acc[0][1][1], fa0, fb1);
          dmma884(acc[1][0][0], acc[1][0][1], fa1, fb0);
          dmma884(acc[1][1][0], acc[1][1][1], fa1, fb1);
        }
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          const int r = rb + 8 * mb + g;
          if (r >= nr) continue;
          const int64_t ra = rowt[row0 + r];
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
            const int cp = 16 * half + 8 * nb + 2 * t;
            if (cp + 1 < cols && pair16 && colt[cp + 1] == colt[cp] + 1) {
              *reinterpret_cast<double2*>(a + ra + colt[cp]) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
            } else {
              if (cp < cols) a[ra + colt[cp]] = acc[mb][nb][0];
              if (cp + 1 < cols) a[ra + colt[cp + 1]] = acc[mb][nb][1];
            }
          }
        }
      }
    }
    tm.sync();  // every warp is done with stage s
    if (tm.wid == 0 && i + 2 < ntile) issue(i + 2);
  }
}
#endif

// ---- one two-site gate -----------------------------------------------------------------------------------------------------------
// Everything the phases of one gate share; in shared memory on the device.  The gate runs as THREE kernels (bottom of this
// file): bp_apply3_sides -- one CTA per (gate, side): tables, message check, absorb, Gram product; bp_apply3_bond -- one
// small CTA per gate: the two eigen-decompositions, the bond problem and its SVD, W_a; bp_apply3_final -- one CTA per
// (gate, side): A' = A W.  Each kernel inlines its phases (one call site each) and has the occupancy its phase wants: the
// Jacobi iterations are pure latency chains and get 5 resident CTAs of 4 warps with every lane busy, the dense passes 2 CTAs
// of 8 warps.  Why not one kernel walking all phases (round 2's first version): with the phases as non-inlined functions
// ptxas's calling convention left each of them ~75 of the 128 registers (the Gram pass spilled its accumulators, whatever
// the caller kept live came off every callee's budget, and the outcome changed with unrelated edits:
// tools/sass_loop_spills.py); fully inlined, the one big function spilled loop state; and half (SVD) or three quarters
// (eigen-decompositions) of the CTA's warps idled at the step barriers of the Jacobi phases.
// The dynamic shared memory of the kernel.  The phases re-derive it from the symbol instead of loading the pointer from the
// context: the compiler then knows the address space (LDS / STS instead of generic loads) in them and in what they call.
template <typename T>
__host__ __device__ __forceinline__ T* phase_smem(T* from_context) {
#ifdef __CUDA_ARCH__
  extern __shared__ __align__(16) unsigned char dyn_smem3[];
  (void)from_context;
  return reinterpret_cast<T*>(dyn_smem3);
#else
  return from_context;
#endif
}

template <typename T>
struct Gate3 {
  const GateDesc* gd;
  T* sites;
  T* msgs;
  const T* ops;
  T* w;   // this gate's work space (Layout3)
  T* tt;  // T = (M_1 x M_2 x ..) A of the side in flight (scratch of the CTA)
  double* sv_out;
  T* smem;
  int* flag;
  int* bad;
  long long* stamps;
  int normalize;
  Layout3 L;
  Tabs tb[2];
};
#ifdef __CUDACC__
struct GramPipe;
#endif

// debug stamps (BPX_APPLY_TIMING=1), STAMP_SLOTS clock64 values per gate: side a: 8 a + {0 start, 1 message check, 2 absorb,
// 3 Gram}; bond kernel: 16 start, 17 eigen-decompositions, 18 theta + gate, 19 SVD, 20 Y / W; Jacobi sweeps: 21 SVD, 22 eig;
// final kernel, side a: 24 + 2 a start, 25 + 2 a end; 32..39: the first column batch of absorb on side 0
constexpr int STAMP_SLOTS = 48;
#ifdef __CUDA_ARCH__
#define BPX_STAMP(i) do { if (c.stamps && tm.tid() == 0) c.stamps[i] = clock64(); } while (0)
#else
#define BPX_STAMP(i) do { } while (0)
#endif

// tables, message check, absorb, Gram product of side a
template <typename T>
__host__ __device__ __forceinline__ void phase_side(const Team tm, Gate3<T>& c, int a, void* gram_pipe = nullptr) {
  constexpr bool CPLX = Elem<T>::is_complex;
  const Side& sd = c.gd->s[a];
  const Layout3& L = c.L;
  T* w = c.w;
  T* smem = phase_smem<T>(c.smem);
  const Walk wka = walk_of(sd);
  {
    const Tabs t = build_tabs<T>(tm, sd, wka, w + L.tab[a]);
#ifdef __CUDA_ARCH__
    if (tm.tid() == 0) c.tb[a] = t;
    tm.sync();
#else
    c.tb[a] = t;
#endif
  }
  message_check<T>(tm, sd, c.msgs, w + L.h[a], smem, c.bad);
  if (*c.bad) return;
  BPX_STAMP(8 * a + 1);
  const T* A = c.sites + sd.site_off;
  if (!CPLX && sd.cols % 2 == 0)
    absorb_side<T, 2>(tm, sd, wka, c.tb[a], A, w + L.h[a], w + L.at[a], c.tt, smem, (c.stamps && a == 0) ? c.stamps + 32 : nullptr);
  else
    absorb_side<T, 1>(tm, sd, wka, c.tb[a], A, w + L.h[a], w + L.at[a], c.tt, smem);
  BPX_STAMP(8 * a + 2);
#ifdef __CUDA_ARCH__
  if constexpr (!CPLX)
    gram_side_mma(tm, sd, w + L.at[a], c.tt, w + L.g[a], smem, *static_cast<GramPipe*>(gram_pipe));
  else
#endif
    gram_side<T, CPLX ? 4 : 8>(tm, sd, w + L.at[a], c.tt, w + L.g[a], smem);
  BPX_STAMP(8 * a + 3);
}

// G_a = V diag(lam) V^H for both sides: one joint iteration when the shapes agree (they do unless the physical dimensions
// differ), each problem in its own slot of shared memory
template <typename T>
__host__ __device__ __forceinline__ void phase_eig(const Team tm, Gate3<T>& c) {
  const GateDesc& gd = *c.gd;
  const Layout3& L = c.L;
  T* w = c.w;
  T* smem = phase_smem<T>(c.smem);
  const bool joint = gd.s[0].cols == gd.s[1].cols;
  for (int a = 0; a < 2; a += joint ? 2 : 1) {
    const int np = joint ? 2 : 1, cols = gd.s[a].cols;
    const int ld = jacobi_ld(cols, cols, np, jacobi_rb<T>());
    const int slot = ld * cols + cols * cols;
    for (int k = 0; k < np; ++k)
      gram_factor_load<T>(tm, cols, w + L.g[a + k], w + L.gb[a + k], reinterpret_cast<double*>(w + L.ev[a + k]), smem + k * slot,
                          smem + k * slot + ld * cols, ld);
    jacobi_groups<T>(tm, smem, cols, cols, ld, c.flag, c.bad, c.stamps ? c.stamps + 22 : nullptr, np, slot);
    for (int k = 0; k < np; ++k)
      gram_factor_finish<T>(tm, cols, gd.s[a + k].nref, w + L.gb[a + k], reinterpret_cast<double*>(w + L.ev[a + k]), w + L.r[a + k],
                            w + L.rinv[a + k], smem + k * slot, smem + k * slot + ld * cols, ld, c.bad);
  }
}

// the bond problem (apply_operators.jl:260-268), R factors with cols rows each: theta = R_1 R_2, the gate, and the operand
// of the SVD iteration (columns sorted by norm) in shared memory
template <typename T>
__host__ __device__ __forceinline__ void phase_bond_pre(const Team tm, Gate3<T>& c) {
  using E = Elem<T>;
  const GateDesc& gd = *c.gd;
  const Layout3& L = c.L;
  T* w = c.w;
  const Side& s1 = gd.s[0];
  const Side& s2 = gd.s[1];
  const int d1 = s1.d, d2 = s2.d, n1 = s1.cols, n2 = s2.cols, chi = gd.chi_b;
  const int m = n1 * d1, n = n2 * d2;
  const T* R1 = w + L.r[0];
  const T* R2 = w + L.r[1];
  T* th0 = w + L.theta[0];
  T* th1 = w + L.theta[1];  // theta after the gate (kept: V = theta^H U / s)
  for (int i = tm.tid(); i < m * n; i += tm.nt()) {
    const int row = i % m, col = i / m;
    const int q1 = row % n1, x1 = row / n1, q2 = col % n2, x2 = col / n2;
    T acc = E::zero();
    for (int b = 0; b < chi; ++b) acc = E::fma(R1[q1 + n1 * (x1 + d1 * b)], R2[q2 + n2 * (x2 + d2 * b)], acc);
    th0[i] = acc;
  }
  tm.sync();
  const T* op = c.ops + gd.op_off;
  const int dd = d1 * d2;
  const int ldb = jacobi_ld(m, n, 1, jacobi_rb<T>());
  T* sb = phase_smem<T>(c.smem);
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
  double* nrm = reinterpret_cast<double*>(w + L.sig);   // (both are overwritten after the iteration)
  int32_t* pos = reinterpret_cast<int32_t*>(w + L.order);
  column_order<T>(tm, th1, m, n, nrm, pos);
  for (int i = tm.tid(); i < m * n; i += tm.nt()) sb[(i % m) + ldb * pos[i / m]] = th1[i];
  tm.sync();
}

// singular values, their order, Y_a (the new R factors) and W_a = R_a^+ Y_a from the rotated operand in shared memory
template <typename T>
__host__ __device__ __forceinline__ void phase_bond_post(const Team tm, Gate3<T>& c) {
  using E = Elem<T>;
  const GateDesc& gd = *c.gd;
  const Layout3& L = c.L;
  T* w = c.w;
  const int m = gd.s[0].cols * gd.s[0].d, n = gd.s[1].cols * gd.s[1].d;
  const int ldb = jacobi_ld(m, n, 1, jacobi_rb<T>());
  const T* sb = phase_smem<T>(c.smem);
  const T* th1 = w + L.theta[1];
  // (the rotated operand U diag(s) is read where the iteration left it, in shared memory)
  double* sig = reinterpret_cast<double*>(w + L.sig);
  int32_t* order = reinterpret_cast<int32_t*>(w + L.order);
  for (int j = tm.tid(); j < n; j += tm.nt()) {
    double a = 0.0;
    for (int r = 0; r < m; ++r) a += E::abs2(sb[r + ldb * j]);
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
  if (c.normalize) {
    double a = 0.0;
    for (int j = 0; j < k; ++j) a += sig[order[j]] * sig[order[j]];
    nrm = a > 0.0 ? sqrt(a) : 1.0;
  }
  // Y_1[q1, (x, kk)] = U[(q1, x), j] sqrt(s'_j);  Y_2[q2, (x, kk)] = sqrt(s'_j) conj(V[(q2, x), j]),  V_j = theta^H u_j / s_j
  for (int a = 0; a < 2; ++a) {
    const Side& sd = gd.s[a];
    T* y = w + L.y[a];
    const int na = sd.cols, da = sd.d;
    for (int i = tm.tid(); i < na * da * k; i += tm.nt()) {
      const int q = i % na, cc = i / na, x = cc % da, kk = cc / da, j = order[kk];
      const double sj = sig[j], snew = sj / nrm;
      T v = E::zero();
      if (sj > 0.0) {
        if (a == 0) {
          v = scal(sb[(q + na * x) + ldb * j], sqrt(snew) / sj);
        } else {
          // conj(V[c2, j]) = sum_r theta[r, c2] conj(u_j[r]) / s_j,  u_j = sb[:, j] / s_j
          const int c2 = q + na * x;
          T acc = E::zero();
          for (int r = 0; r < m; ++r) acc = E::fma(th1[r + m * c2], E::conj(sb[r + ldb * j]), acc);
          v = scal(acc, sqrt(snew) / (sj * sj));
        }
      }
      y[i] = v;
    }
  }
  tm.sync();
  // W_a = R_a^+ Y_a, zero padded to PC x PC ([c][c'] row-major for the final pass)
  for (int a = 0; a < 2; ++a) {
    const Side& sd = gd.s[a];
    const int na = sd.cols, nc = sd.d * k;
    const T* y = w + L.y[a];
    const T* ri = w + L.rinv[a];
    T* W = w + L.w[a];
    for (int i = tm.tid(); i < PC * PC; i += tm.nt()) {
      const int cp = i % PC, cc = i / PC;
      T acc = E::zero();
      if (cc < na && cp < nc)
        for (int q = 0; q < na; ++q) acc = E::fma(ri[cc + na * q], y[q + na * cp], acc);
      W[cc * PC + cp] = acc;
    }
  }
  tm.sync();
}

// A'_a = A_a W_a in place
template <typename T>
__host__ __device__ __forceinline__ void phase_final_side(const Team tm, Gate3<T>& c, int a, void* pipe = nullptr) {
  constexpr bool CPLX = Elem<T>::is_complex;
  const Side& sd = c.gd->s[a];
  const Layout3& L = c.L;
  T* w = c.w;
  T* smem = phase_smem<T>(c.smem);
  T* A = c.sites + sd.site_off;
#ifdef __CUDA_ARCH__
  if constexpr (!CPLX) {
    if (pipe) {
      final_side_mma(tm, sd, c.tb[a], w + L.at[a], A, w + L.w[a], smem, *static_cast<GramPipe*>(pipe));
      return;
    }
  }
#endif
  (void)pipe;
  if (CPLX)
    final_side<T, 128, 1, 8>(tm, sd, c.tb[a], w + L.at[a], A, w + L.w[a], smem);
  else
    final_side<T, 256, 2, 16>(tm, sd, c.tb[a], w + L.at[a], A, w + L.w[a], smem);
}
// the two messages of the gate edge = diag(S / |S|), the singular values
template <typename T>
__host__ __device__ __forceinline__ void phase_final_bond(const Team tm, Gate3<T>& c) {
  using E = Elem<T>;
  const GateDesc& gd = *c.gd;
  const Layout3& L = c.L;
  T* w = c.w;
  const double* sig = reinterpret_cast<const double*>(w + L.sig);
  const int32_t* order = reinterpret_cast<const int32_t*>(w + L.order);
  const int chi = gd.chi_b, k = gd.k;
  double nrm = 1.0;
  if (c.normalize) {
    double a = 0.0;
    for (int j = 0; j < k; ++j) a += sig[order[j]] * sig[order[j]];
    nrm = a > 0.0 ? sqrt(a) : 1.0;
  }
  T* msgs = c.msgs;
  for (int i = tm.tid(); i < chi * chi; i += tm.nt()) {
    const int r = i % chi, cc = i / chi;
    const T v = (r == cc && r < k) ? from_real<T>(sig[order[r]] / nrm) : E::zero();
    msgs[gd.msg12 + i] = v;
    msgs[gd.msg21 + i] = v;
  }
  if (c.sv_out)
    for (int i = tm.tid(); i < chi; i += tm.nt()) c.sv_out[i] = i < k ? sig[order[i]] / nrm : 0.0;
  tm.sync();
}
// the SVD iteration of the bond matrix (operand in shared memory, left there by phase_bond_pre)
template <typename T>
__host__ __device__ __forceinline__ void phase_bond_svd(const Team tm, Gate3<T>& c) {
  const GateDesc& gd = *c.gd;
  const int m = gd.s[0].cols * gd.s[0].d, n = gd.s[1].cols * gd.s[1].d;
  jacobi_groups<T>(tm, phase_smem<T>(c.smem), m, n, jacobi_ld(m, n, 1, jacobi_rb<T>()), c.flag, c.bad, c.stamps ? c.stamps + 21 : nullptr);
}

// One gate from start to end on ONE team: the host entry (tests/native/apply_host.cu runs exactly the device code of the
// phases).  Returns (to every thread) 0 when the gate was applied, 1 when it was left untouched for the fallback.
template <typename T>
__host__ __device__ int run_two_site_v3(const Team tm, const GateDesc& gd, T* sites, T* msgs, const T* ops, T* w, double* sv_out,
                                        int normalize, int* flag, int* bad, T* smem, long long* stamps = nullptr) {
  Gate3<T> c;
  c.gd = &gd;
  c.sites = sites;
  c.msgs = msgs;
  c.ops = ops;
  c.w = w;
  c.sv_out = sv_out;
  c.smem = smem;
  c.flag = flag;
  c.bad = bad;
  c.stamps = stamps;
  c.normalize = normalize;
  c.L = layout3_of(gd);
  c.tt = w + c.L.tt;
  if (tm.tid() == 0) *bad = 0;
  tm.sync();
  for (int a = 0; a < 2; ++a) {
    phase_side<T>(tm, c, a);
    if (*c.bad) return 1;
  }
  phase_eig<T>(tm, c);
  if (*c.bad) return 1;
  phase_bond_pre<T>(tm, c);
  phase_bond_svd<T>(tm, c);
  if (*c.bad) return 1;
  phase_bond_post<T>(tm, c);
  for (int a = 0; a < 2; ++a) phase_final_side<T>(tm, c, a);
  phase_final_bond<T>(tm, c);
  return 0;
}

#ifdef __CUDACC__
struct ApplyArgs3 {
  ApplyArgs base;       // base.ws: one work space of ws_stride elements PER GATE of the chunk [g0, g1)
  int64_t ws_stride;
  void* tt;             // bp_apply3_sides: one scratch buffer of tt_stride elements per CTA
  int64_t tt_stride;
  int64_t g0, g1;       // the gates of this launch
  int32_t* status;      // per gate of the batch: 0 applied, 1 left for the fallback (zeroed before the first kernel)
  long long* stamps;    // debug (BPX_APPLY_TIMING=1): STAMP_SLOTS clock64 stamps per gate, or NULL
};

constexpr int NT_BOND = 128;  // bp_apply3_bond: 64 columns = 32 pairs x 4 lanes

template <typename T>
__device__ __forceinline__ void gate3_fill(Gate3<T>& c, const ApplyArgs3& a3, int64_t g, int* flag, int* bad) {
  extern __shared__ __align__(16) unsigned char dyn_smem3[];
  const ApplyArgs& a = a3.base;
  const GateDesc& gd = a.gates[g];
  const int64_t sv_row = gd.sv_row_p1 > 0 ? gd.sv_row_p1 - 1 : g;
  c.gd = &gd;
  c.sites = static_cast<T*>(a.sites);
  c.msgs = static_cast<T*>(a.msgs);
  c.ops = static_cast<const T*>(a.ops);
  c.w = static_cast<T*>(a.ws) + (g - a3.g0) * a3.ws_stride;
  c.tt = static_cast<T*>(a3.tt) + (int64_t)blockIdx.x * a3.tt_stride;
  c.sv_out = a.sv_out ? a.sv_out + sv_row * a.sv_stride : nullptr;
  c.smem = reinterpret_cast<T*>(dyn_smem3);
  c.flag = flag;
  c.bad = bad;
  c.stamps = a3.stamps ? a3.stamps + STAMP_SLOTS * g : nullptr;
  c.normalize = a.normalize;
  c.L = layout3_of(gd, false);
  *bad = 0;
}
#undef BPX_STAMP
#define BPX_STAMP(i) do { if (c.stamps && threadIdx.x == 0) c.stamps[i] = clock64(); } while (0)

// (gate, side) items: tables, message check, absorb, Gram product -> at[a], G_a in the gate's work space
template <typename T>
__global__ void __launch_bounds__(NT, 2) bp_apply3_sides(ApplyArgs3 a3) {
  __shared__ int flag, bad;
  __shared__ Gate3<T> c;
  Team tm;
  tm.lane = threadIdx.x & 31;
  tm.wid = threadIdx.x >> 5;
  tm.nw = NT / 32;
  __shared__ uint64_t gram_bar[2];
  GramPipe gp;
  gram_pipe_init(gp, gram_bar);
  const int64_t nitems = 2 * (a3.g1 - a3.g0);
  for (int64_t it = blockIdx.x; it < nitems; it += gridDim.x) {
    const int64_t g = a3.g0 + (it >> 1);
    const int a = (int)(it & 1);
    __syncthreads();
    if (threadIdx.x == 0) gate3_fill<T>(c, a3, g, &flag, &bad);
    __syncthreads();
    BPX_STAMP(8 * a);
    phase_side<T>(tm, c, a, &gp);
    __syncthreads();
    if (threadIdx.x == 0 && bad) a3.status[g] = 1;  // (both sides may store the same 1)
  }
}

// one gate per CTA: G_a -> R_a, R_a^+; the bond problem; W_a
template <typename T, int CTAS>
__global__ void __launch_bounds__(NT_BOND, CTAS) bp_apply3_bond(ApplyArgs3 a3) {
  __shared__ int flag, bad;
  __shared__ Gate3<T> c;
  Team tm;
  tm.lane = threadIdx.x & 31;
  tm.wid = threadIdx.x >> 5;
  tm.nw = NT_BOND / 32;
  for (int64_t g = a3.g0 + blockIdx.x; g < a3.g1; g += gridDim.x) {
    if (a3.status[g] != 0) continue;  // (uniform: written by the previous kernel)
    __syncthreads();
    if (threadIdx.x == 0) gate3_fill<T>(c, a3, g, &flag, &bad);
    __syncthreads();
    BPX_STAMP(16);
    phase_eig<T>(tm, c);
    BPX_STAMP(17);
    if (!bad) {
      phase_bond_pre<T>(tm, c);
      BPX_STAMP(18);
      phase_bond_svd<T>(tm, c);
      BPX_STAMP(19);
    }
    if (!bad) {
      phase_bond_post<T>(tm, c);
      BPX_STAMP(20);
    }
    __syncthreads();
    if (threadIdx.x == 0 && bad) a3.status[g] = 1;
  }
}

// (gate, side) items: A'_a = A_a W_a; the item of side 0 also writes the messages and the singular values
template <typename T>
__global__ void __launch_bounds__(NT, 2) bp_apply3_final(ApplyArgs3 a3) {
  __shared__ int flag, bad;
  __shared__ Gate3<T> c;
  Team tm;
  tm.lane = threadIdx.x & 31;
  tm.wid = threadIdx.x >> 5;
  tm.nw = NT / 32;
  __shared__ uint64_t tile_bar[2];
  GramPipe gp;
  gram_pipe_init(gp, tile_bar);
  const int64_t nitems = 2 * (a3.g1 - a3.g0);
  for (int64_t it = blockIdx.x; it < nitems; it += gridDim.x) {
    const int64_t g = a3.g0 + (it >> 1);
    const int a = (int)(it & 1);
    if (a3.status[g] != 0) continue;
    __syncthreads();
    if (threadIdx.x == 0) {
      gate3_fill<T>(c, a3, g, &flag, &bad);
      const Side& sd = c.gd->s[a];
      const Walk wk = walk_of(sd);
      c.tb[a] = tabs_at<T>(sd, wk, c.w + c.L.tab[a]);
    }
    __syncthreads();
    BPX_STAMP(24 + 2 * a);
    phase_final_side<T>(tm, c, a, &gp);
    if (a == 0) phase_final_bond<T>(tm, c);
    BPX_STAMP(25 + 2 * a);
  }
}
#undef BPX_STAMP
#endif

}  // namespace applyk3
}  // namespace bpx
