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
// Written against the `Team` abstraction like versions 1 / 2: tests/native/apply_host.cu runs it on the host.
#pragma once
#include "bpx_apply2.cuh"

namespace bpx {
namespace applyk3 {

using namespace bpx::applyk;

constexpr int PC = 32;               // padded column count of the tiles; the fast path takes sides with cols <= PC
constexpr int TRG = 128;             // rows per tile of the Gram pass
constexpr int PCP = PC + 2;          // row stride of the Gram tiles: 16-byte aligned rows, stores two wavefronts per warp
constexpr double COND_MIN = 1e-10;   // smallest kept eigenvalue of G relative to the largest
constexpr double PIVOT_MIN = 1e-12;  // smallest Cholesky pivot of a message relative to its largest diagonal entry
constexpr int MAXDIM = 16;           // largest external link dimension (a fibre lives in registers)

struct Layout3 {
  int64_t at[2];        // per side: the matrix view of A as a row-major rows x PCP matrix (tiles of rows are contiguous)
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

__host__ __device__ inline Layout3 layout3_of(const GateDesc& g) {
  Layout3 L;
  int64_t o = 0;
  for (int a = 0; a < 2; ++a) { L.at[a] = o; o += g.s[a].rows * PCP; }
  L.tt = o; o += (g.s[0].rows > g.s[1].rows ? g.s[0].rows : g.s[1].rows) * PCP;
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
#ifdef __CUDACC__
__device__ int g_jacobi_gs_shift = 0;  // debug knob (BPX_APPLY_GS_SHIFT): halve the lane-group width this many times
#endif
__host__ __device__ inline int jacobi_gs(int n, int nw) {  // lanes per column pair (measured: narrow groups win)
  const int npairs = (n + 1) / 2;
  (void)nw;
  int gs = npairs >= 16 ? 4 : (npairs >= 4 ? 8 : 16);
#ifdef __CUDA_ARCH__
  gs >>= g_jacobi_gs_shift;
  if (gs < 1) gs = 1;
#endif
  return gs;
}
__host__ __device__ inline int jacobi_ld(int m, int n, int nw = NT / 32) {
  const int gs = jacobi_gs(n, nw);
  if (gs >= 16) return m;
  int ld = m;
  while (ld % 16 != 8) ++ld;  // sized for the widest rule on the host; the kernel only needs callers and callee to agree
#ifdef __CUDA_ARCH__
  ld = m;
  while (ld % 16 != (gs < 2 ? 1 : gs)) ++ld;
#endif
  return ld;
}

// Can the fast path take this gate, and how much shared memory (elements of T) does it want?  0: not supported.
__host__ __device__ inline int64_t smem_need(const GateDesc& g, bool cplx) {
  if (g.nsides != 2) return 0;
  int64_t need = 0;
  for (int a = 0; a < 2; ++a) {
    const Side& s = g.s[a];
    if (s.cols > PC || s.cols < 1) return 0;
    int64_t hsum = 0;
    for (int i = 0; i < s.z; ++i)
      if (i != s.bond_slot) {
        if (s.dim[i] > MAXDIM) return 0;
        hsum += (int64_t)s.dim[i] * s.dim[i];
      }
    const int64_t cb = (!cplx && (s.cols % 2 == 0)) ? 2 : 1;
    int64_t dim0 = 1;
    for (int i = 0; i < s.z; ++i)
      if (i != s.bond_slot) { dim0 = s.dim[i]; break; }
    const int64_t absorb = cb * (s.rows + s.rows / dim0) + 2 + hsum;  // padded column batch + the messages
    const int64_t gram = 2 * (int64_t)TRG * PCP;               // A tile, T tile (the reduction re-uses them)
    const int64_t trf = cplx ? 128 : 256;
    const int64_t fin = trf * PCP + (int64_t)PC * PC;          // A tile, W
    need = need > absorb ? need : absorb;
    need = need > gram ? need : gram;
    need = need > fin ? need : fin;
  }
  const int64_t m = (int64_t)g.s[0].cols * g.s[0].d, n = (int64_t)g.s[1].cols * g.s[1].d;
  const int64_t ldb = jacobi_ld((int)m, (int)n);
  const int64_t svd = ldb * n;
  need = need > svd ? need : svd;
  const int64_t gramf = (int64_t)(PC + 16) * PC + (int64_t)PC * PC;  // rotated + unrotated Gram matrix
  need = need > gramf ? need : gramf;
  // the cross-warp reduction of the Gram pass: 8 partial tiles of PC x (PC or PC/2) elements
  const int64_t red = 8 * (int64_t)PC * (cplx ? PC / 2 : PC);
  need = need > red ? need : red;
  return (need + 1) & ~(int64_t)1;
}

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
__host__ __device__ __noinline__ Tabs build_tabs(const Team tm, const Side& sd, const Walk& wk, T* slot) {
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

// contiguous copy global -> shared in 16-byte pieces (both 16-byte aligned; n elements, n * sizeof(T) a multiple of 16)
template <typename T>
__host__ __device__ __noinline__ void copy_tile(const Team tm, T* dst, const T* src, int64_t n) {
#ifdef __CUDA_ARCH__
  const int n16 = (int)(n * sizeof(T) / 16), nt = tm.nt(), tid = tm.tid();
  const double2* __restrict__ s2 = reinterpret_cast<const double2*>(src);
  double2* __restrict__ d2 = reinterpret_cast<double2*>(dst);
  constexpr int U = 4;
  const int nfull = n16 / (U * nt) * (U * nt);
  for (int i0 = tid; i0 < nfull; i0 += U * nt) {  // no predicates: the staging registers stay registers
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = s2[i0 + u * nt];
#pragma unroll
    for (int u = 0; u < U; ++u) d2[i0 + u * nt] = v[u];
  }
  for (int i = nfull + tid; i < n16; i += nt) d2[i] = s2[i];
#else
  for (int64_t i = tm.tid(); i < n; i += tm.nt()) dst[i] = src[i];
#endif
  tm.sync();
}

// ---- messages: Hermitian part + positive-definiteness check ---------------------------------------------------------------
// One warp per message: right-looking Cholesky on a scratch copy; *bad is set when a pivot is not safely positive.
template <typename T>
__host__ __device__ __noinline__ void message_check(const Team tm, const Side& sd, const T* msgs, T* H, T* scratch, int* bad) {
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
__host__ __device__ __noinline__ void absorb_side(const Team tm, const Side& sd, const Walk& wk, const Tabs tb, const T* a, const T* H, T* aout,
                                     T* tout, T* smem, long long* stamps = nullptr) {
#ifdef __CUDA_ARCH__
#define BPX_ASTAMP(i) do { if (stamps && tm.tid() == 0 && c0 == 0) stamps[i] = clock64(); } while (0)
#else
#define BPX_ASTAMP(i) do { (void)stamps; } while (0)
#endif
  const int rows = (int)sd.rows, ncols = sd.cols;
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
        aout[(int64_t)r * PCP + c0 + b] = v[u];  // the matrix view, row-major: the Gram and final passes read plain tiles
      }
    }
    for (int i = nfull + tid; i < total; i += nt) {
      const int b = i % CB, r = i / CB;
      const T v = a[rowt[r] + cofs[b]];
      col[b * prow + r + dpad.div(r)] = v;
      aout[(int64_t)r * PCP + c0 + b] = v;
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
      tout[(int64_t)r * PCP + c0 + b] = col[b * prow + r + dpad.div(r)];
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
__host__ __device__ __noinline__ void gram_side(const Team tm, const Side& sd, const T* at, const T* tt, T* G, T* smem) {
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
      copy_tile<T>(tm, sA, at + row0 * PCP, (int64_t)nr * PCP);
      copy_tile<T>(tm, sT, tt + row0 * PCP, (int64_t)nr * PCP);
#ifdef __CUDA_ARCH__
      const int li = tm.lane >> 2, lj = tm.lane & 3;
      const T* pa = sA + 4 * li;
      const T* pt = sT + TJ * (lj + 4 * pass);
      for (int r = tm.wid; r < nr; r += tm.nw) {
        T av[4], tv[TJ];
        if constexpr (!E::is_complex) {  // rows of the tiles are 16-byte aligned (PCP even)
          const double2* a2 = reinterpret_cast<const double2*>(pa + r * PCP);
          const double2* t2 = reinterpret_cast<const double2*>(pt + r * PCP);
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            const double2 q = a2[x];
            av[2 * x] = q.x;
            av[2 * x + 1] = q.y;
          }
#pragma unroll
          for (int y = 0; y < TJ / 2; ++y) {
            const double2 q = t2[y];
            tv[2 * y] = q.x;
            tv[2 * y + 1] = q.y;
          }
        } else {
          const double2* a2 = reinterpret_cast<const double2*>(pa + r * PCP);
          const double2* t2 = reinterpret_cast<const double2*>(pt + r * PCP);
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const double2 q = a2[x];
            av[x] = E::conj(*reinterpret_cast<const T*>(&q));
          }
#pragma unroll
          for (int y = 0; y < TJ; ++y) {
            const double2 q = t2[y];
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
              o = E::fma(E::conj(sA[r * PCP + 4 * li + x]), sT[r * PCP + TJ * (lj + 4 * pass) + y], o);
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

// 1 / sqrt(x) and 1 / x to full double accuracy from the hardware's 20-bit approximations + two Newton steps: a short
// dependent chain (the IEEE sqrt / division sequences cost several hundred cycles of latency each, and the Jacobi steps
// below are pure latency).  Outside the safe exponent range the exact functions are used.
template <int STEPS = 2>
__host__ __device__ __forceinline__ double rsqrt_d(double x) {
#ifdef __CUDA_ARCH__
  if (!(x > 1e-280 && x < 1e280)) return rsqrt(x);
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
template <int STEPS = 2>
__host__ __device__ __forceinline__ double rcp_d(double x) {
#ifdef __CUDA_ARCH__
  if (!(fabs(x) > 1e-280 && fabs(x) < 1e280)) return 1.0 / x;
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
// round-robin step.  The iteration is LATENCY bound (one dependent chain per step: loads, inner products, shuffles, rotation
// parameters, rotation, barrier), so the groups are as wide as the team allows (64 columns on 8 warps: 8 lanes per pair,
// 8 rows per lane): the shortest per-lane chains.  A lane's rows of both columns stay in registers between the inner products
// and the rotation (RC per column).
template <typename T, int GS, int RC>
__host__ __device__ __noinline__ void jacobi_groups_t(const Team tm, T* B, int m, int n, int ld, int* flag, int* not_converged,
                                                      long long* sweeps_out) {
  using E = Elem<T>;
  const int L = tm.lanes();
  const int np = (n + 1) & ~1, npairs = np / 2;
  const int gpw = L / GS;                // groups per warp
  const int sl = tm.lane % GS, grp = tm.lane / GS;
  const bool cached = m <= GS * RC;
  const double tol2 = (double)m * EPS * EPS;
  double fro2 = 0.0;
  for (int64_t i = tm.lane; i < (int64_t)m * n; i += L) fro2 += E::abs2(B[(i % m) + (int64_t)ld * (i / m)]);
  fro2 = tm.sum(fro2);
  const double zero2 = (double)n * n * EPS * EPS * fro2;
  tm.sync();
  int f = 1, nsweeps = 0;
  for (int sweep = 0; sweep < MAX_JACOBI_SWEEPS && f; ++sweep) {
    ++nsweeps;
    if (tm.tid() == 0) *flag = 0;
    tm.sync();
    for (int step = 0; step < np - 1; ++step) {
      for (int base = tm.wid * gpw; base < npairs; base += tm.nw * gpw) {  // warp-uniform bound: shuffles stay converged
        const int idx = base + grp;
        int p = 0, q = 1;
        bool active = idx < npairs;
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
        T* bp = B + p * ld;
        T* bq = B + q * ld;
        double a = 0.0, b = 0.0;
        T g = E::zero();
        T xr[RC], yr[RC];
        if (cached) {
          double a1 = 0.0, b1 = 0.0;  // two partial sums per inner product: half the dependent-chain length
          T g1 = E::zero();
#pragma unroll
          for (int j = 0; j < RC; ++j) {
            const int r = sl + j * GS;
            xr[j] = E::zero();
            yr[j] = E::zero();
            if (active && r < m) {  // (groups without a pair this step must not touch the matrix: another group owns columns 0, 1)
              xr[j] = bp[r];
              yr[j] = bq[r];
            }
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
          a += a1;
          b += b1;
          g = E::add(g, g1);
        } else {
          for (int r = sl; r < m && active; r += GS) {
            const T x = bp[r], y = bq[r];
            a += E::abs2(x);
            b += E::abs2(y);
            g = E::fma(E::conj(x), y, g);
          }
        }
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int o = GS >> 1; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
          g = E::add(g, E::shfl_xor(g, o));
        }
#endif
        const double g2 = E::abs2(g);
        if (!active || !(g2 > tol2 * a * b) || !(a > zero2) || !(b > zero2)) continue;  // group-uniform, no shuffles below
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
        if (cached) {
#pragma unroll
          for (int j = 0; j < RC; ++j) {
            const int r = sl + j * GS;
            if (r < m) {
              bp[r] = sub(scal(xr[j], c), E::mul(yr[j], scph));
              bq[r] = E::add(E::mul(xr[j], sph), scal(yr[j], c));
            }
          }
        } else {
          for (int r = sl; r < m; r += GS) {
            const T x = bp[r], y = bq[r];
            bp[r] = sub(scal(x, c), E::mul(y, scph));
            bq[r] = E::add(E::mul(x, sph), scal(y, c));
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
                                       long long* sweeps_out = nullptr) {
  if (n < 2) return;
#ifdef __CUDA_ARCH__
  constexpr int RC = 8;
  switch (jacobi_gs(n, tm.nw)) {
    case 1: jacobi_groups_t<T, 1, 2 * RC>(tm, B, m, n, ld, flag, not_converged, sweeps_out); break;
    case 2: jacobi_groups_t<T, 2, 2 * RC>(tm, B, m, n, ld, flag, not_converged, sweeps_out); break;
    case 4: jacobi_groups_t<T, 4, 2 * RC>(tm, B, m, n, ld, flag, not_converged, sweeps_out); break;
    case 8: jacobi_groups_t<T, 8, RC>(tm, B, m, n, ld, flag, not_converged, sweeps_out); break;
    default: jacobi_groups_t<T, 16, RC>(tm, B, m, n, ld, flag, not_converged, sweeps_out); break;
  }
#else
  jacobi_groups_t<T, 1, 1>(tm, B, m, n, ld, flag, not_converged, sweeps_out);  // host lanes: one lane per pair
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
template <typename T>
__host__ __device__ __noinline__ void gram_factor(const Team tm, int cols, int rank_max, const T* G, T* Gb, double* ev, T* R, T* Rinv, T* smem, int* flag,
                                     int* bad, long long* sweeps_out = nullptr) {
  using E = Elem<T>;
  const int ld = jacobi_ld(cols, cols);
  T* sb = smem;
  T* sg = smem + (int64_t)ld * cols;  // the unrotated matrix, for the Rayleigh quotients
  int32_t* pos = reinterpret_cast<int32_t*>(Gb);  // (Gb is written after the iteration)
  column_order<T>(tm, G, cols, cols, ev, pos);
  for (int i = tm.tid(); i < cols * cols; i += tm.nt()) {
    const T v = G[i];
    sb[(i % cols) + ld * pos[i / cols]] = v;
    sg[i] = v;
  }
  tm.sync();
  jacobi_groups<T>(tm, sb, cols, cols, ld, flag, bad, sweeps_out);
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
  tm.sync();
  for (int j = tm.tid(); j < cols; j += tm.nt()) {
    int before = 0;
    for (int i = 0; i < cols; ++i) before += (ev[i] > ev[j] || (ev[i] == ev[j] && i < j)) ? 1 : 0;
    if (before >= rank_max || !(ev[j] > cut)) ev[j] = 0.0;
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
// Tiles of TRF rows as [row][PCP] in shared memory (plain contiguous copies).  A thread owns RT rows x NO outputs: per
// column pair RT 16-byte operand loads (rows 272 bytes apart: conflict free) and NO 16-byte broadcast loads of W feed 2 RT NO
// FMAs.
template <typename T, int TRF, int RT, int NO>
__host__ __device__ __noinline__ void final_side(const Team tm, const Side& sd, const Tabs tb, const T* at, T* a, const T* W, T* smem) {
  using E = Elem<T>;
  T* sA = smem;                                  // [r][PCP]
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
    copy_tile<T>(tm, sA, at + row0 * PCP, (int64_t)nr * PCP);
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
          const T* pa = sA + (rg + x * RG) * PCP + c;
#ifdef __CUDA_ARCH__
          if constexpr (!E::is_complex) {
            const double2 q = *reinterpret_cast<const double2*>(pa);
            a0[x] = q.x;
            a1[x] = two ? q.y : 0.0;
          } else {
            const double2 q0 = reinterpret_cast<const double2*>(pa)[0];
            a0[x] = *reinterpret_cast<const T*>(&q0);
            a1[x] = E::zero();
            if (two) {
              const double2 q1 = reinterpret_cast<const double2*>(pa)[1];
              a1[x] = *reinterpret_cast<const T*>(&q1);
            }
          }
#else
          a0[x] = pa[0];
          a1[x] = two ? pa[1] : E::zero();
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

// One two-site gate.  Returns (to every thread) 0 when the gate was applied, 1 when it was left untouched for the fallback.
template <typename T>
__host__ __device__ int run_two_site_v3(const Team tm, const GateDesc& gd, T* sites, T* msgs, const T* ops, T* w /* this CTA's work space */,
                                        double* sv_out, int normalize, int* flag, int* bad, T* smem, long long* stamps = nullptr) {
  using E = Elem<T>;
  constexpr bool CPLX = Elem<T>::is_complex;
  const Layout3 L = layout3_of(gd);
#ifdef __CUDA_ARCH__
#define BPX_STAMP(i) do { if (stamps && tm.tid() == 0) stamps[i] = clock64(); } while (0)
#else
#define BPX_STAMP(i) do { (void)stamps; } while (0)
#endif
  BPX_STAMP(0);
  if (tm.tid() == 0) *bad = 0;
  tm.sync();
  Tabs tb[2];
  for (int a = 0; a < 2; ++a) {
    const Side& sd = gd.s[a];
    const Walk wka = walk_of(sd);
    tb[a] = build_tabs<T>(tm, sd, wka, w + L.tab[a]);
    message_check<T>(tm, sd, msgs, w + L.h[a], smem, bad);
    if (*bad) return 1;
    BPX_STAMP(1 + 4 * a);
    const T* A = sites + sd.site_off;
    if (!CPLX && sd.cols % 2 == 0)
      absorb_side<T, 2>(tm, sd, wka, tb[a], A, w + L.h[a], w + L.at[a], w + L.tt, smem, (stamps && a == 0) ? stamps + 16 : nullptr);
    else
      absorb_side<T, 1>(tm, sd, wka, tb[a], A, w + L.h[a], w + L.at[a], w + L.tt, smem);
    BPX_STAMP(2 + 4 * a);
    gram_side<T, CPLX ? 4 : 8>(tm, sd, w + L.at[a], w + L.tt, w + L.g[a], smem);
    BPX_STAMP(3 + 4 * a);
    gram_factor<T>(tm, sd.cols, sd.nref, w + L.g[a], w + L.gb[a], reinterpret_cast<double*>(w + L.ev[a]), w + L.r[a], w + L.rinv[a], smem,
                   flag, bad, stamps ? stamps + 14 + a : nullptr);
    if (*bad) return 1;
    BPX_STAMP(4 + 4 * a);
  }
  // ---- the bond problem (apply_operators.jl:260-268), R factors with cols rows each -----------------------------------
  const Side& s1 = gd.s[0];
  const Side& s2 = gd.s[1];
  const int d1 = s1.d, d2 = s2.d, n1 = s1.cols, n2 = s2.cols, chi = gd.chi_b;
  const int m = n1 * d1, n = n2 * d2;
  const T* R1 = w + L.r[0];
  const T* R2 = w + L.r[1];
  T* th0 = w + L.theta[0];
  T* th1 = w + L.theta[1];  // theta after the gate (kept: V = theta^H U / s)
  T* th2 = w + L.theta[2];  // rotated copy: U diag(s)
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
  const int ldb = jacobi_ld(m, n);
  T* sb = smem;
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
  {
    double* nrm = reinterpret_cast<double*>(w + L.sig);   // (both are overwritten after the iteration)
    int32_t* pos = reinterpret_cast<int32_t*>(w + L.order);
    column_order<T>(tm, th1, m, n, nrm, pos);
    for (int i = tm.tid(); i < m * n; i += tm.nt()) sb[(i % m) + (int64_t)ldb * pos[i / m]] = th1[i];
  }
  tm.sync();
  BPX_STAMP(9);
  jacobi_groups<T>(tm, sb, m, n, ldb, flag, bad, stamps ? stamps + 13 : nullptr);
  if (*bad) return 1;
  BPX_STAMP(10);
  double* sig = reinterpret_cast<double*>(w + L.sig);
  int32_t* order = reinterpret_cast<int32_t*>(w + L.order);
  for (int i = tm.tid(); i < m * n; i += tm.nt()) th2[i] = sb[(i % m) + (int64_t)ldb * (i / m)];
  for (int j = tm.tid(); j < n; j += tm.nt()) {
    double a = 0.0;
    for (int r = 0; r < m; ++r) a += E::abs2(sb[r + (int64_t)ldb * j]);
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
  // Y_1[q1, (x, kk)] = U[(q1, x), j] sqrt(s'_j);  Y_2[q2, (x, kk)] = sqrt(s'_j) conj(V[(q2, x), j]),  V_j = theta^H u_j / s_j
  for (int a = 0; a < 2; ++a) {
    const Side& sd = gd.s[a];
    T* y = w + L.y[a];
    const int na = sd.cols, da = sd.d;
    for (int i = tm.tid(); i < na * da * k; i += tm.nt()) {
      const int q = i % na, c = i / na, x = c % da, kk = c / da, j = order[kk];
      const double sj = sig[j], snew = sj / nrm;
      T v = E::zero();
      if (sj > 0.0) {
        if (a == 0) {
          v = scal(th2[(q + na * x) + (int64_t)m * j], sqrt(snew) / sj);
        } else {
          // conj(V[c2, j]) = sum_r theta[r, c2] conj(u_j[r]) / s_j,  u_j = th2[:, j] / s_j
          const int c2 = q + na * x;
          T acc = E::zero();
          for (int r = 0; r < m; ++r) acc = E::fma(th1[r + (int64_t)m * c2], E::conj(th2[r + (int64_t)m * j]), acc);
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
      const int cp = i % PC, c = i / PC;
      T acc = E::zero();
      if (c < na && cp < nc)
        for (int q = 0; q < na; ++q) acc = E::fma(ri[c + na * q], y[q + na * cp], acc);
      W[c * PC + cp] = acc;
    }
  }
  tm.sync();
  BPX_STAMP(11);
  for (int a = 0; a < 2; ++a) {
    const Side& sd = gd.s[a];
    T* A = sites + sd.site_off;
    if (CPLX)
      final_side<T, 128, 1, 8>(tm, sd, tb[a], w + L.at[a], A, w + L.w[a], smem);
    else
      final_side<T, 256, 2, 16>(tm, sd, tb[a], w + L.at[a], A, w + L.w[a], smem);
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
  BPX_STAMP(12);
#undef BPX_STAMP
  return 0;
}

#ifdef __CUDACC__
struct ApplyArgs3 {
  ApplyArgs base;       // base.ws: one work space of ws_stride elements PER CTA (not per gate: a layer is one launch)
  int64_t ws_stride;
  int32_t* status;      // per gate: 0 applied, 1 left for the fallback
  long long* stamps;    // debug (BPX_APPLY_TIMING=1): 16 clock64 stamps per gate, or NULL
};

template <typename T>
__global__ void __launch_bounds__(NT, 2) bp_apply_gates_v3(ApplyArgs3 a3) {
  extern __shared__ __align__(16) unsigned char dyn_smem3[];
  __shared__ int flag, bad;
  const ApplyArgs& a = a3.base;
  Team tm;
  tm.lane = threadIdx.x & 31;
  tm.wid = threadIdx.x >> 5;
  tm.nw = NT / 32;
  for (int64_t g = blockIdx.x; g < a.n_gates; g += gridDim.x) {
    const int64_t sv_row = a.gates[g].sv_row_p1 > 0 ? a.gates[g].sv_row_p1 - 1 : g;
    const int st = run_two_site_v3<T>(tm, a.gates[g], static_cast<T*>(a.sites), static_cast<T*>(a.msgs), static_cast<const T*>(a.ops),
                                      static_cast<T*>(a.ws) + (int64_t)blockIdx.x * a3.ws_stride,
                                      a.sv_out ? a.sv_out + sv_row * a.sv_stride : nullptr, a.normalize, &flag, &bad,
                                      reinterpret_cast<T*>(dyn_smem3), a3.stamps ? a3.stamps + 32 * g : nullptr);
    if (threadIdx.x == 0) a3.status[g] = st;
    __syncthreads();
  }
}
#endif

}  // namespace applyk3
}  // namespace bpx
