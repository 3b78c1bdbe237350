// BPX_KERNEL_ONCHIP, ComplexF64 variant for link dimensions <= 8, degree 2..4, any physical dimension d (complex PEPS on
// a square lattice: the complex twin of BASELINE config 2).  Link dimensions below 8 (also different ones per leg) are
// ZERO-PADDED to 8 in the kernel's private tensor image; message fragments, old values and stores are masked to the
// true dimensions -- exact, because every padded tensor entry is zero.
//
// The kernel is a template over the element type.  CPLX = false is the general REAL variant -- any physical dimension (the
// 16-byte chunk holds the physical pair (s = 2q, s = 2q + 1), one slice per pair, odd d padded with a zero), link
// dimensions <= 8, also per-leg different -- for the Float64 buckets the hand-tuned (chi = 8, d = 2) kernel of
// bpx_onchip.cuh does not cover (BASELINE config 1: chi = 2).  Two DMMA chains per operand pair instead of four.
//
// Same idea as bpx_onchip16c.cuh: a complex tensor is processed one PHYSICAL SLICE at a time.  The slice
// A_s[(re, im), a0, a1, a2, a3] has exactly the shape and the XOR-swizzled layout (leg_pos) of the real kernel's
// [s, a0..a3] tile (8192 doubles = 64 KiB; degree 3: 1024, degree 2: 128), one LDS.128 feeds the real and the imaginary
// DMMA operand, and a complex MAC is four real DMMA issues (sign flips on the integer pipe).  Per slice the
// leave-one-out tree of bpx_onchip.cuh runs unchanged -- branch P = A_s·M0·M1 -> out3, out2; branch Q = A_s·M2·M3 ->
// out1, out0, register-chained absorb-absorb and absorb-close groups -- and the closure accumulators are summed over
// the slices in registers.  8 warps (2 per scheduler: up to 255 registers per thread), a 2-slot TMA ring without a
// producer warp (the last warp to release a slot refills it), host-side longest-processing-time schedule laid out as
// rounds, block-wide epilogue (two warps per 8x8 output tile).
#pragma once
#include "bpx_onchip.cuh"
#include "bpx_onchip16c.cuh"  // neg(), Cursor

namespace bpx {
namespace onchip8c {

using onchip::CHI;
using onchip::MSG;
using onchip::NELEM;
using onchip::bar_sync;
using onchip::col_pos;
using onchip::dmma;
using onchip::fence_proxy_async;
using onchip::leg_pos;
using onchip::mbar_expect_tx;
using onchip::mbar_init;
using onchip::mbar_wait;
using onchip::n_cols;
using onchip::tma_bulk_g2s;
using onchip16c::Cursor;
using onchip16c::Tr;
using onchip16c::neg;

constexpr int NW = 8;
constexpr int NT = NW * 32;
constexpr int CMSG8 = 2 * MSG;  // doubles per complex 8x8 message
constexpr int MAXT = 3;         // output tiles per item (degree 3: three, else two)

struct ItemDesc {
  int64_t site_off;     // DOUBLES, into the private image (slice s at + s * slice doubles)
  int64_t canon_off;    // complex elements, into the canonical site buffer
  int64_t in_off[4];    // complex elements: message arriving on leg i
  int64_t out_off[MAXT];  // per output tile, in the order the kernel produces them
  int32_t out_edge[MAXT];
  int32_t peer[MAXT];
  int32_t kind;         // -1 null | 0 degree 4 branch P (out3, out2) | 1 degree 4 branch Q (out1, out0) | 2 degree 3 | 3 degree 2
  int32_t d;            // number of slices: complex: physical dimension; real: physical PAIRS, (phys + 1) / 2
  int32_t first;        // this item swizzles the vertex's tensor
  int32_t phys;         // physical dimension
  int64_t need;         // streamed host I/O: prefix of the upload that holds every message this item reads
  int32_t dim[4];       // true link dimension per leg (<= 8; absent legs: 1)
  int32_t out_dim[MAXT];  // link dimension of each output tile's message
  int32_t pad2;
};

struct Args {
  const ItemDesc* items;
  int n_slots;
  const double* sites;  // private image
  const double* msg_in;
  double* msg_out;
  unsigned long long* resmax;
  int normalize;
  PeerArgs peer;
  HostIO io;
  unsigned long long stop_key;  // device-side convergence test (sweep_already_converged), 0: none
};

__device__ __forceinline__ int slice_doubles(int kind) { return kind <= 1 ? NELEM : (kind == 2 ? NELEM / 8 : NELEM / 64); }
__device__ __forceinline__ int n_tiles(int kind) { return kind == 2 ? 3 : 2; }

struct CMsgFrag {
  double mar[2], mai[2];  // M[g, t + 4j]  : A operand of "absorb first leg", B operand of the T-GEMM
  double mbr[2], mbi[2];  // M[g, 2t + i]  : B operand of "absorb second leg" (register-chained)
};
// M is chi x chi (chi <= 8), column-major; entries beyond chi read as zero
template <bool CPLX>
__device__ __forceinline__ CMsgFrag load_cfrag8(const typename Tr<CPLX>::T* __restrict__ M, int g, int t, int chi) {
  CMsgFrag f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int ca = t + 4 * j, cb = 2 * t + j;
    double2 a = make_double2(0.0, 0.0), b = a;
    if (g < chi && ca < chi) a = Tr<CPLX>::pack(M[g + chi * ca]);
    if (g < chi && cb < chi) b = Tr<CPLX>::pack(M[g + chi * cb]);
    f.mar[j] = a.x;
    f.mai[j] = a.y;
    f.mbr[j] = b.x;
    f.mbi[j] = b.y;
  }
  return f;
}

// D[x' = g, y = 2t + i] = sum_x M[x', x] src[x, y]   (accumulators start at zero).  Complex: (xr, xi) = real / imaginary
// part.  Real: xr / xi are the two physical values of the chunk, each its own chain.
template <bool CPLX>
__device__ __forceinline__ void cabsorb(const CMsgFrag& m, const double2& b0, const double2& b1, double (&xr)[2], double (&xi)[2]) {
  xr[0] = xr[1] = xi[0] = xi[1] = 0.0;
  if (CPLX) {
    dmma(xr[0], xr[1], m.mar[0], b0.x);
    dmma(xi[0], xi[1], m.mai[0], b0.x);
    dmma(xr[0], xr[1], m.mai[0], neg(b0.y));
    dmma(xi[0], xi[1], m.mar[0], b0.y);
    dmma(xr[0], xr[1], m.mar[1], b1.x);
    dmma(xi[0], xi[1], m.mai[1], b1.x);
    dmma(xr[0], xr[1], m.mai[1], neg(b1.y));
    dmma(xi[0], xi[1], m.mar[1], b1.y);
  } else {
    dmma(xr[0], xr[1], m.mar[0], b0.x);
    dmma(xi[0], xi[1], m.mar[0], b0.y);
    dmma(xr[0], xr[1], m.mar[1], b1.x);
    dmma(xi[0], xi[1], m.mar[1], b1.y);
  }
}

// dst[.., x', y', ..] = sum_{x,y} MX[x', x] MY[y', y] src[.., x, y, ..]   (legs X then Y absorbed; dst != src)
template <bool CPLX, int X, int Y, int C0, int C1>
__device__ __forceinline__ void absorb_pair_c(const double* src, double* dst, const CMsgFrag& mx, const CMsgFrag& my, int warp, int g, int t) {
  const uint32_t ld0 = leg_pos(X, t) ^ leg_pos(Y, g), ld1 = leg_pos(X, t + 4) ^ leg_pos(Y, g);
  const uint32_t st0 = leg_pos(X, g) ^ leg_pos(Y, 2 * t), st1 = leg_pos(X, g) ^ leg_pos(Y, 2 * t + 1);
#pragma unroll 2
  for (int col = warp; col < n_cols<C0, C1>(); col += NW) {
    const uint32_t base = onchip::col_pos<C0, C1>(col);
    const double2 b0 = *reinterpret_cast<const double2*>(src + (base ^ ld0));
    const double2 b1 = *reinterpret_cast<const double2*>(src + (base ^ ld1));
    double xr[2], xi[2];
    cabsorb<CPLX>(mx, b0, b1, xr, xi);
    // absorb Y from registers: D2[x' = g, y' = 2t + i'] = sum_{y = 2t + i} D1[x', y] MY[y', y]
    double pr0 = 0, pr1 = 0, pi0 = 0, pi1 = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (CPLX) {
        dmma(pr0, pr1, xr[i], my.mbr[i]);
        dmma(pi0, pi1, xr[i], my.mbi[i]);
        dmma(pr0, pr1, neg(xi[i]), my.mbi[i]);
        dmma(pi0, pi1, xi[i], my.mbr[i]);
      } else {
        dmma(pr0, pr1, xr[i], my.mbr[i]);
        dmma(pi0, pi1, xi[i], my.mbr[i]);
      }
    }
    *reinterpret_cast<double2*>(dst + (base ^ st0)) = make_double2(pr0, pi0);
    *reinterpret_cast<double2*>(dst + (base ^ st1)) = make_double2(pr1, pi1);
  }
}

// dst[.., x', y, ..] = sum_x MX[x', x] src[.., x, y, ..]   (single absorption; Y is a passive tile leg)
template <bool CPLX, int X, int Y, int C0, int C1>
__device__ __forceinline__ void absorb_one_c(const double* src, double* dst, const CMsgFrag& mx, int warp, int g, int t) {
  const uint32_t ld0 = leg_pos(X, t) ^ leg_pos(Y, g), ld1 = leg_pos(X, t + 4) ^ leg_pos(Y, g);
  const uint32_t st0 = leg_pos(X, g) ^ leg_pos(Y, 2 * t), st1 = leg_pos(X, g) ^ leg_pos(Y, 2 * t + 1);
  for (int col = warp; col < n_cols<C0, C1>(); col += NW) {
    const uint32_t base = onchip::col_pos<C0, C1>(col);
    const double2 b0 = *reinterpret_cast<const double2*>(src + (base ^ ld0));
    const double2 b1 = *reinterpret_cast<const double2*>(src + (base ^ ld1));
    double xr[2], xi[2];
    cabsorb<CPLX>(mx, b0, b1, xr, xi);
    *reinterpret_cast<double2*>(dst + (base ^ st0)) = make_double2(xr[0], xi[0]);
    *reinterpret_cast<double2*>(dst + (base ^ st1)) = make_double2(xr[1], xi[1]);
  }
}

// partial closure tile of one warp: acc[i][c] accumulates out[v' = g, v = 2t + c] over the k-split i (summed at the end)
struct CAcc {
  double r[2][2], i[2][2];
};
__device__ __forceinline__ void cacc_zero(CAcc& a) {
#pragma unroll
  for (int i = 0; i < 2; ++i) a.r[i][0] = a.r[i][1] = a.i[i][0] = a.i[i][1] = 0.0;
}

// acc[v', v] += sum_{cols, u'} conj(A[.., u', v']) * ( sum_u MU[u', u] P[.., u, v] )   (leg U absorbed on the fly, V open)
template <bool CPLX, int U, int V, int C0, int C1>
__device__ __forceinline__ void absorb_close_c(const double* P, const double* A, const CMsgFrag& mu, int warp, int g, int t, CAcc& acc) {
  const uint32_t lp0 = leg_pos(U, t) ^ leg_pos(V, g), lp1 = leg_pos(U, t + 4) ^ leg_pos(V, g);
  const uint32_t la0 = leg_pos(U, 2 * t) ^ leg_pos(V, g), la1 = leg_pos(U, 2 * t + 1) ^ leg_pos(V, g);
#pragma unroll 2
  for (int col = warp; col < n_cols<C0, C1>(); col += NW) {
    const uint32_t base = onchip::col_pos<C0, C1>(col);
    const double2 p0 = *reinterpret_cast<const double2*>(P + (base ^ lp0));
    const double2 p1 = *reinterpret_cast<const double2*>(P + (base ^ lp1));
    const double2 a0 = *reinterpret_cast<const double2*>(A + (base ^ la0));
    const double2 a1 = *reinterpret_cast<const double2*>(A + (base ^ la1));
    // T[v = g, u' = 2t + i] = sum_u P[u, v] MU[u', u]
    double tr0 = 0, tr1 = 0, ti0 = 0, ti1 = 0;
    if (CPLX) {
      dmma(tr0, tr1, p0.x, mu.mar[0]);
      dmma(ti0, ti1, p0.x, mu.mai[0]);
      dmma(tr0, tr1, neg(p0.y), mu.mai[0]);
      dmma(ti0, ti1, p0.y, mu.mar[0]);
      dmma(tr0, tr1, p1.x, mu.mar[1]);
      dmma(ti0, ti1, p1.x, mu.mai[1]);
      dmma(tr0, tr1, neg(p1.y), mu.mai[1]);
      dmma(ti0, ti1, p1.y, mu.mar[1]);
      // out[v' = g, v] += sum_{u' = 2t + i} conj(A[u', v']) T[v, u']
      dmma(acc.r[0][0], acc.r[0][1], a0.x, tr0);
      dmma(acc.i[0][0], acc.i[0][1], a0.x, ti0);
      dmma(acc.r[1][0], acc.r[1][1], a1.x, tr1);
      dmma(acc.i[1][0], acc.i[1][1], a1.x, ti1);
      dmma(acc.r[0][0], acc.r[0][1], a0.y, ti0);
      dmma(acc.i[0][0], acc.i[0][1], neg(a0.y), tr0);
      dmma(acc.r[1][0], acc.r[1][1], a1.y, ti1);
      dmma(acc.i[1][0], acc.i[1][1], neg(a1.y), tr1);
    } else {  // (tr, ti) and (acc.r, acc.i): the two physical values of the chunk; summed when the tile is stored
      dmma(tr0, tr1, p0.x, mu.mar[0]);
      dmma(ti0, ti1, p0.y, mu.mar[0]);
      dmma(tr0, tr1, p1.x, mu.mar[1]);
      dmma(ti0, ti1, p1.y, mu.mar[1]);
      dmma(acc.r[0][0], acc.r[0][1], a0.x, tr0);
      dmma(acc.i[0][0], acc.i[0][1], a0.y, ti0);
      dmma(acc.r[1][0], acc.r[1][1], a1.x, tr1);
      dmma(acc.i[1][0], acc.i[1][1], a1.y, ti1);
    }
  }
}

// shared memory (doubles): slot[2][NELEM] | P[NELEM] | red[NW][MAXT][CMSG8] | part[128] | 2 mbarriers | 2 counters
constexpr size_t SMEM_DOUBLES8C = (size_t)3 * NELEM + NW * MAXT * CMSG8 + 128 + 4;
constexpr size_t SMEM_BYTES8C = SMEM_DOUBLES8C * sizeof(double);
enum { BAR_C8 = 1 };

// canonical A_v[s, a0..] (column-major, true dims) -> private image: slices [chunk, a0..] in leg_pos order, every leg
// zero-padded to 8; chunk = (re, im) of physical value s (complex) or the physical pair (2s, 2s + 1) (real)
template <bool CPLX>
__global__ void swizzle_sites_c8(const ItemDesc* items, int n_slots, const double* __restrict__ src, double* __restrict__ dst) {
  for (int it = blockIdx.x; it < n_slots; it += gridDim.x) {
    const ItemDesc d = items[it];
    if (d.kind < 0 || !d.first) continue;
    const int nsl = slice_doubles(d.kind), nb = nsl / 2;
    const double* s0 = src + (CPLX ? 2 : 1) * d.canon_off;
    double* d0 = dst + d.site_off;
    for (int c = threadIdx.x; c < nb * d.d; c += blockDim.x) {
      const int s = c % d.d, b = c / d.d;
      const int a0 = b & 7, a1 = (b >> 3) & 7, a2 = (b >> 6) & 7, a3 = (b >> 9) & 7;
      const uint32_t p = leg_pos(0, a0) ^ leg_pos(1, a1) ^ leg_pos(2, a2) ^ leg_pos(3, a3);
      double2 v = make_double2(0.0, 0.0);
      if (a0 < d.dim[0] && a1 < d.dim[1] && a2 < d.dim[2] && a3 < d.dim[3]) {
        const size_t lin = (size_t)d.phys * (a0 + (size_t)d.dim[0] * (a1 + (size_t)d.dim[1] * (a2 + (size_t)d.dim[2] * a3)));
        if (CPLX) {
          v = *reinterpret_cast<const double2*>(s0 + 2 * (lin + s));
        } else {
          v.x = s0[lin + 2 * s];
          if (2 * s + 1 < d.phys) v.y = s0[lin + 2 * s + 1];
        }
      }
      *reinterpret_cast<double2*>(d0 + (size_t)s * nsl + p) = v;
    }
  }
}

template <bool CPLX>
__device__ __forceinline__ void store_tile(double* mine, const CAcc& a, int g, int t) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int el = g + CHI * (2 * t + c);  // out[v', v] at v' + 8 v
    const double re = a.r[0][c] + a.r[1][c], im = a.i[0][c] + a.i[1][c];
    *reinterpret_cast<double2*>(mine + 2 * el) = CPLX ? make_double2(re, im) : make_double2(re + im, 0.0);
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(NT, 1) bp_update_onchip_c8x(Args k) {
  using T = typename Tr<CPLX>::T;
  using E = Elem<T>;
  extern __shared__ __align__(128) double smem[];
  double* Pbuf = smem + 2 * NELEM;
  double* red = smem + 3 * NELEM;
  double* part = red + NW * MAXT * CMSG8;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(part + 128);  // full[2]
  unsigned int* cnt = reinterpret_cast<unsigned int*>(mbar + 2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  if (sweep_already_converged(k.resmax, k.stop_key)) return;
  const int G = gridDim.x;

  auto valid = [&](const Cursor& c) { return c.idx < k.n_slots && k.items[c.idx].kind >= 0; };
  auto issue = [&](const Cursor& c, int sl) {  // one thread
    const ItemDesc* d = k.items + c.idx;
    const int nsl = slice_doubles(d->kind);
    mbar_expect_tx(&mbar[sl], nsl * 8);
    const double* src = k.sites + d->site_off + (size_t)c.s * nsl;
    if (nsl == NELEM) {
#pragma unroll
      for (int q = 0; q < 4; ++q) tma_bulk_g2s(smem + sl * NELEM + q * 2048, src + q * 2048, 16384, &mbar[sl]);
    } else {
      tma_bulk_g2s(smem + sl * NELEM, src, nsl * 8, &mbar[sl]);
    }
  };
  auto advance = [&](Cursor& c) {
    if (!valid(c)) return;
    if (++c.s >= k.items[c.idx].d) {
      c.s = 0;
      c.idx += G;
    }
  };
  Cursor cur{(int)blockIdx.x, 0};
  if (threadIdx.x == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    cnt[0] = cnt[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    Cursor c = cur;
    for (int sl = 0; sl < 2; ++sl) {
      if (valid(c)) issue(c, sl);
      advance(c);
    }
  }
  advance(cur);
  advance(cur);  // -> unit 2
  if (warp == 0) peer_gate(k.peer, lane);  // the message fragments below may have been written by peers
  __syncthreads();

  auto release = [&](int sl) {  // warp-uniform; the LAST warp to release a slot refills it with the slice two units ahead
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      if (atomicAdd(&cnt[sl], 1u) == NW - 1) {
        atomicExch(&cnt[sl], 0u);
        __threadfence_block();
        if (valid(cur)) {
          fence_proxy_async();
          issue(cur, sl);
        }
      }
    }
    advance(cur);
  };

  uint32_t u = 0;
  for (int idx = blockIdx.x; idx < k.n_slots; idx += G) {
    const ItemDesc* d = k.items + idx;
    const int kind = d->kind;
    if (kind < 0) break;
    const int nd = d->d, ntile = n_tiles(kind);
    if (k.io.progress) {  // streamed upload: the item's messages (fragments, old values) have arrived -- ONE warp polls
      if (warp == 0) hostio_wait(k.io, d->need);
      bar_sync(BAR_C8, NT);
    }
    // old value of the output element this thread finalises: tile = thread / 64, element = thread % 64 (issued early)
    const int my_tile = threadIdx.x >> 6, my_el = threadIdx.x & 63;
    const int chi_o = my_tile < ntile ? d->out_dim[my_tile] : 0;
    const bool my_valid = (my_el & 7) < chi_o && (my_el >> 3) < chi_o;  // inside the true chi_o x chi_o message
    const int64_t my_off = my_valid ? d->out_off[my_tile] + (my_el & 7) + chi_o * (my_el >> 3) : 0;
    T old = E::zero();
    if (my_valid) old = reinterpret_cast<const T*>(k.msg_in)[my_off];
    CAcc acc[MAXT];
#pragma unroll
    for (int i = 0; i < MAXT; ++i) cacc_zero(acc[i]);
    const T* min_c = reinterpret_cast<const T*>(k.msg_in);
    if (kind <= 1) {
      const CMsgFrag m0 = load_cfrag8<CPLX>(min_c + d->in_off[0], g, t, d->dim[0]);
      const CMsgFrag m1 = load_cfrag8<CPLX>(min_c + d->in_off[1], g, t, d->dim[1]);
      const CMsgFrag m2 = load_cfrag8<CPLX>(min_c + d->in_off[2], g, t, d->dim[2]);
      const CMsgFrag m3 = load_cfrag8<CPLX>(min_c + d->in_off[3], g, t, d->dim[3]);
      for (int s = 0; s < nd; ++s, ++u) {
        const int sl = u & 1;
        mbar_wait(&mbar[sl], (u >> 1) & 1);
        const double* A = smem + sl * NELEM;
        if (kind == 0) {  // P = A·M0·M1 -> out3 (absorb 2, close 3), out2 (absorb 3, close 2)
          absorb_pair_c<CPLX, 0, 1, 2, 3>(A, Pbuf, m0, m1, warp, g, t);
          bar_sync(BAR_C8, NT);
          absorb_close_c<CPLX, 2, 3, 0, 1>(Pbuf, A, m2, warp, g, t, acc[0]);
          absorb_close_c<CPLX, 3, 2, 0, 1>(Pbuf, A, m3, warp, g, t, acc[1]);
        } else {          // Q = A·M2·M3 -> out1 (absorb 0, close 1), out0 (absorb 1, close 0)
          absorb_pair_c<CPLX, 2, 3, 0, 1>(A, Pbuf, m2, m3, warp, g, t);
          bar_sync(BAR_C8, NT);
          absorb_close_c<CPLX, 0, 1, 2, 3>(Pbuf, A, m0, warp, g, t, acc[0]);
          absorb_close_c<CPLX, 1, 0, 2, 3>(Pbuf, A, m1, warp, g, t, acc[1]);
        }
        release(sl);
        bar_sync(BAR_C8, NT);  // P is rewritten by the next slice
      }
    } else if (kind == 2) {
      const CMsgFrag m0 = load_cfrag8<CPLX>(min_c + d->in_off[0], g, t, d->dim[0]);
      const CMsgFrag m1 = load_cfrag8<CPLX>(min_c + d->in_off[1], g, t, d->dim[1]);
      const CMsgFrag m2 = load_cfrag8<CPLX>(min_c + d->in_off[2], g, t, d->dim[2]);
      for (int s = 0; s < nd; ++s, ++u) {
        const int sl = u & 1;
        mbar_wait(&mbar[sl], (u >> 1) & 1);
        const double* A = smem + sl * NELEM;
        // X = A·M0 -> out2 (absorb 1, close 2), out1 (absorb 2, close 1);  X' = A·M2 -> out0 (absorb 1, close 0)
        absorb_one_c<CPLX, 0, 1, 2, -1>(A, Pbuf, m0, warp, g, t);
        bar_sync(BAR_C8, NT);
        absorb_close_c<CPLX, 1, 2, 0, -1>(Pbuf, A, m1, warp, g, t, acc[0]);
        absorb_close_c<CPLX, 2, 1, 0, -1>(Pbuf, A, m2, warp, g, t, acc[1]);
        bar_sync(BAR_C8, NT);
        absorb_one_c<CPLX, 2, 1, 0, -1>(A, Pbuf, m2, warp, g, t);
        bar_sync(BAR_C8, NT);
        absorb_close_c<CPLX, 1, 0, 2, -1>(Pbuf, A, m1, warp, g, t, acc[2]);
        release(sl);
        bar_sync(BAR_C8, NT);
      }
    } else {  // degree 2: out1 (absorb 0, close 1), out0 (absorb 1, close 0) straight from A
      const CMsgFrag m0 = load_cfrag8<CPLX>(min_c + d->in_off[0], g, t, d->dim[0]);
      const CMsgFrag m1 = load_cfrag8<CPLX>(min_c + d->in_off[1], g, t, d->dim[1]);
      for (int s = 0; s < nd; ++s, ++u) {
        const int sl = u & 1;
        mbar_wait(&mbar[sl], (u >> 1) & 1);
        const double* A = smem + sl * NELEM;
        absorb_close_c<CPLX, 0, 1, -1, -1>(A, A, m0, warp, g, t, acc[0]);
        absorb_close_c<CPLX, 1, 0, -1, -1>(A, A, m1, warp, g, t, acc[1]);
        release(sl);
      }
    }
    // ---- cross-warp reduction + block-wide epilogue: two warps per output tile, one element per thread ----
#pragma unroll
    for (int i = 0; i < MAXT; ++i)
      if (i < ntile) store_tile<CPLX>(red + (warp * MAXT + i) * CMSG8, acc[i], g, t);
    bar_sync(BAR_C8, NT);
    {
      double2 vsum = make_double2(0.0, 0.0);
      if (my_valid) {
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          const double2 q = *reinterpret_cast<const double2*>(red + (w * MAXT + my_tile) * CMSG8 + 2 * my_el);
          vsum.x += q.x;
          vsum.y += q.y;
        }
      }
      const T v = Tr<CPLX>::unpack(vsum);
      double* part1 = part;       // [NW] (complex) sums
      double* part2 = part + 32;  // [NW][4]
      const T sw = warp_sum<T>(v);
      if (lane == 0) *reinterpret_cast<double2*>(part1 + 2 * warp) = Tr<CPLX>::pack(sw);
      bar_sync(BAR_C8, NT);
      const double2 q0 = *reinterpret_cast<const double2*>(part1 + 2 * (warp & ~1)), q1 = *reinterpret_cast<const double2*>(part1 + 2 * (warp | 1));
      const T s = Tr<CPLX>::unpack(make_double2(q0.x + q1.x, q0.y + q1.y));
      T x = v;
      if (my_valid) {
        if (k.normalize && !E::is_zero(s)) x = E::div(v, s);
        const int64_t off = my_off;
        reinterpret_cast<T*>(k.msg_out)[off] = x;
        if (k.io.host_out) reinterpret_cast<T*>(k.io.host_out)[off] = x;
        if (k.peer.nranks > 1 && d->peer[my_tile] >= 0) {
          reinterpret_cast<T*>(k.peer.peer_out[d->peer[my_tile]])[off] = x;
          __threadfence_system();  // released here instead of at the kernel's tail
        }
      }
      const double2 dot = Tr<CPLX>::pack(warp_sum<T>(E::fma(E::conj(old), x, E::zero())));
      const double n_old = warp_sum_d(E::abs2(old)), n_new = warp_sum_d(my_valid ? E::abs2(x) : 0.0);
      if (lane == 0) {
        double* q = part2 + 4 * warp;
        q[0] = dot.x;
        q[1] = dot.y;
        q[2] = n_old;
        q[3] = n_new;
      }
      bar_sync(BAR_C8, NT);
      if (lane == 0 && (warp & 1) == 0 && my_tile < ntile) {
        const double* a = part2 + 4 * warp;
        const double* b = a + 4;
        const double dr = a[0] + b[0], di = a[1] + b[1];
        residual_record(k.resmax, 1.0 - (dr * dr + di * di) / ((a[2] + b[2]) * (a[3] + b[3])));
      }
    }
  }
  peer_post_when_last(k.peer, false);  // peer stores were released where they were issued
  hostio_finish(k.io);
}

}  // namespace onchip8c
}  // namespace bpx
