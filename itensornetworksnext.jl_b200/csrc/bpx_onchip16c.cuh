// BPX_KERNEL_ONCHIP, ComplexF64 variant: link dimensions <= 16, degree 1..3, any physical dimension d -- BASELINE
// config 3 (heavy-hex lattice, chi = 16, mixed degree 2 / 3 buckets).  Smaller (also per-leg different) link dimensions
// are ZERO-PADDED to 16 in the private tensor image; message fragments, old values and stores are masked to the true
// dimensions -- exact, because every padded tensor entry is zero.
//
// A complex tensor A[s, b0, b1, b2] is processed one PHYSICAL SLICE at a time: the slice A_s[(re, im), b0, b1, b2]
// has exactly the shape of the real kernels' [s, b0, b1, b2] tile (8192 doubles = 64 KiB, layout L_A3 of
// bpx_sliced.cuh, the 16-byte chunk now holds (re, im) of one element instead of (s = 0, s = 1)), and the update is
// a sum over slices:   out[b', b] = sum_s  sum_{..} conj(A_s[b', a'..]) (A_s absorbed with the incoming messages).
// So one LDS.128 still feeds two DMMA operands; a complex MAC is four real DMMA issues
//     Dr += Mr·Br + Mi·(-Bi),   Di += Mi·Br + Mr·Bi        (absorption)
//     acc_r += Ar·Tr + Ai·Ti,   acc_i += Ar·Ti + (-Ai)·Tr   (closure with conj(A))
// with the sign flips done on the integer pipe (DMMA has no negate modifier), and the register chaining of
// bpx_onchip.cuh (accumulator fragment = operand fragment of the next GEMM) holds per real / imaginary component.
//
// The kernel is a template over the element type; CPLX = false is the general REAL variant (the chunk holds the physical
// pair (2q, 2q + 1), one slice per pair, odd d padded with a zero; two DMMA chains per operand pair instead of four) for
// Float64 buckets with degree 1..3 and link dims <= 16 outside the hand-tuned shapes.
//
// Work items (cfg3 has 127 vertices for 148 SMs, so the sweep is ONE wave and its length is the longest item):
//   kind 0  degree 3, ONE output leg:  X = A_s·M_first (64 KiB, shared memory) -> absorb M_second, close the out leg.
//           Per-output items (3 GEMM units each) instead of the per-vertex leave-one-out tree (8 units): the
//           critical path is what counts here, and 108 + 89 + 2 items fill the machine.
//   kind 1  degree 2 vertex, both outputs: slices are 4 KiB (layout L_Z2), four warps = (output, half tile).
//   kind 2  degree 1 vertex: out[b', b] = sum_s A[s, b] conj(A[s, b']), one element per thread.
// The host assigns items to CTAs with a longest-processing-time schedule and lays them out as rounds
// (slot r * grid + cta, padded with null items), so the kernel needs no atomics.
// Slices stream through a 2-slot TMA ring (full mbarriers; the last warp to release a slot refills it).
#pragma once
#include "bpx_sliced.cuh"

namespace bpx {
namespace onchip16c {

using namespace sliced;  // pos<>, dmma, mbarrier / TMA helpers, CHI, MSG

constexpr int NCWC = 8;
constexpr int NCTC = NCWC * 32;
constexpr int NTHREADSC = NCTC;  // no producer warp: see the slice ring in the kernel
constexpr int NSL3 = 8192, NSL2 = 512, NSL1 = 32;  // doubles per physical slice, by degree
constexpr int CMSG = 2 * MSG;                      // doubles per complex 16x16 message

struct ItemDesc {
  int64_t site_off;   // DOUBLES, into the private image (slice s at + s * NSL)
  int64_t in_off[2];  // complex elements.  kind 0: (first absorbed, second absorbed); kind 1: (M0, M1)
  int64_t out_off[2];
  int32_t out_edge[2];
  int32_t peer[2];
  int32_t kind;       // -1: null (padding of the round layout)
  int32_t leg;        // kind 0: output leg
  int32_t d;          // number of slices: complex: physical dimension; real: physical PAIRS, (phys + 1) / 2
  int32_t first;      // this item swizzles the vertex's tensor (one item per vertex)
  int64_t canon_off;  // complex elements, into the canonical site buffer
  int64_t need;       // streamed host I/O: message-set prefix (elements) that holds every message this item reads
  int32_t dim[3];     // true link dimension per leg (<= 16; absent legs: 1)
  int32_t in_dim[2];  // dimension of the messages at in_off[0..1]
  int32_t out_dim[2]; // dimension of the messages at out_off[0..1]
  int32_t phys;       // physical dimension
};

struct Args {
  const ItemDesc* items;
  int n_slots;          // rounds * grid
  const double* sites;  // private image
  const double* msg_in;
  double* msg_out;
  unsigned long long* resmax;
  int normalize;
  PeerArgs peer;
  HostIO io;          // streamed host I/O (bpx_sweep_host), all NULL otherwise
  long long* timing;  // debug (BPX_ONCHIP_TIMING builds): clock64 stamps of CTA 0's first item
  unsigned long long stop_key;  // device-side convergence test (sweep_already_converged), 0: none
};

#ifdef BPX_ONCHIP_TIMING
#define TSTAMPC(i) do { if (lane == 0 && blockIdx.x == 0 && k.timing && u < 2) k.timing[warp * 16 + (i) + 5 * (int)u] = clock64(); } while (0)
#define TSTAMPC0(i) do { if (lane == 0 && blockIdx.x == 0 && k.timing && idx == (int)blockIdx.x) k.timing[warp * 16 + (i)] = clock64(); } while (0)
#else
#define TSTAMPC(i) do { } while (0)
#define TSTAMPC0(i) do { } while (0)
#endif

__device__ __forceinline__ double neg(double x) { return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x)); }

template <bool CPLX>
struct Tr;
template <>
struct Tr<true> {
  using T = c64;
  __device__ static __forceinline__ double2 pack(c64 a) { return make_double2(a.re, a.im); }
  __device__ static __forceinline__ c64 unpack(double2 q) { return make_c64(q.x, q.y); }
};
template <>
struct Tr<false> {
  using T = double;
  __device__ static __forceinline__ double2 pack(double a) { return make_double2(a, 0.0); }
  __device__ static __forceinline__ double unpack(double2 q) { return q.x; }
};

struct CFrag {  // M[g + 8 mt, t + 4 j], real and imaginary parts
  double r[2][4], i[2][4];
};
// M is chi x chi (chi <= 16), column-major; entries beyond chi read as zero
template <bool CPLX>
__device__ __forceinline__ CFrag load_cfrag(const typename Tr<CPLX>::T* __restrict__ M, int g, int t, int chi) {
  CFrag f;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = g + 8 * mt, col = t + 4 * j;
      double2 v = make_double2(0.0, 0.0);
      if (row < chi && col < chi) v = Tr<CPLX>::pack(M[row + chi * col]);
      f.r[mt][j] = v.x;
      f.i[mt][j] = v.y;
    }
  return f;
}

// dst[x', y, c] = sum_x M[x', x] src[x, y, c]  (complex; one column c of the spectator leg; dst != src)
template <bool CPLX, int X, int Y>
__device__ __forceinline__ void absorb_one16c(const double* src, double* dst, uint32_t base, const CFrag& m, int g, int t) {
  double2 b[4][2];
  double nbi[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      b[j][h] = *reinterpret_cast<const double2*>(src + (base ^ pos<L_A3>(X, t + 4 * j) ^ pos<L_A3>(Y, g + 8 * h)));
      nbi[j][h] = neg(b[j][h].y);
    }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double p0 = 0, p1 = 0, q0 = 0, q1 = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (CPLX) {
          dmma(p0, p1, m.r[mt][j], b[j][h].x);
          dmma(q0, q1, m.i[mt][j], b[j][h].x);
          dmma(p0, p1, m.i[mt][j], nbi[j][h]);
          dmma(q0, q1, m.r[mt][j], b[j][h].y);
        } else {  // (p, q): the two physical values of the chunk
          dmma(p0, p1, m.r[mt][j], b[j][h].x);
          dmma(q0, q1, m.r[mt][j], b[j][h].y);
        }
      }
      const uint32_t a = base ^ pos<L_A3>(X, g + 8 * mt);
      *reinterpret_cast<double2*>(dst + (a ^ pos<L_A3>(Y, 2 * t + 8 * h))) = make_double2(p0, q0);
      *reinterpret_cast<double2*>(dst + (a ^ pos<L_A3>(Y, 2 * t + 1 + 8 * h))) = make_double2(p1, q1);
    }
}

// acc[v', v] += sum_{u'} conj(A[u', v']) * ( sum_u M[u', u] P[u, v] )   (complex; HSEL >= 0: only the v-tile HSEL)
template <bool CPLX, int LAY, int U, int V, int HSEL>
__device__ __forceinline__ void absorb_close16c(const double* P, const double* A, uint32_t base, const CFrag& m, int g, int t,
                                                double (&accr)[2][2][2], double (&acci)[2][2][2]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (HSEL >= 0 && h != HSEL) continue;
    double2 p[4];
    double npi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      p[j] = *reinterpret_cast<const double2*>(P + (base ^ pos<LAY>(U, t + 4 * j) ^ pos<LAY>(V, g + 8 * h)));
      npi[j] = neg(p[j].y);
    }
    double tr[2][2], ti[2][2];  // T[v = g + 8h, u' = 2t + i + 8 nt]
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      tr[nt][0] = tr[nt][1] = ti[nt][0] = ti[nt][1] = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (CPLX) {
          dmma(tr[nt][0], tr[nt][1], p[j].x, m.r[nt][j]);
          dmma(ti[nt][0], ti[nt][1], p[j].x, m.i[nt][j]);
          dmma(tr[nt][0], tr[nt][1], npi[j], m.i[nt][j]);
          dmma(ti[nt][0], ti[nt][1], p[j].y, m.r[nt][j]);
        } else {
          dmma(tr[nt][0], tr[nt][1], p[j].x, m.r[nt][j]);
          dmma(ti[nt][0], ti[nt][1], p[j].y, m.r[nt][j]);
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const double2 a = *reinterpret_cast<const double2*>(A + (base ^ pos<LAY>(U, 2 * t + i + 8 * nt) ^ pos<LAY>(V, g + 8 * mt)));
          if (CPLX) {
            const double nai = neg(a.y);
            dmma(accr[mt][h][0], accr[mt][h][1], a.x, tr[nt][i]);
            dmma(acci[mt][h][0], acci[mt][h][1], a.x, ti[nt][i]);
            dmma(accr[mt][h][0], accr[mt][h][1], a.y, ti[nt][i]);
            dmma(acci[mt][h][0], acci[mt][h][1], nai, tr[nt][i]);
          } else {  // (accr, acci): the two physical values; summed when the partial tile is stored
            dmma(accr[mt][h][0], accr[mt][h][1], a.x, tr[nt][i]);
            dmma(acci[mt][h][0], acci[mt][h][1], a.y, ti[nt][i]);
          }
        }
  }
}

// shared memory (doubles): slot[2][NSL3] | X[NSL3] (aliased by red[NCWC][CMSG]) | raw[2][CMSG] | 2 mbarriers | 2 counters
constexpr size_t SMEM_DOUBLES16C = (size_t)3 * NSL3 + 2 * CMSG + 4;
constexpr size_t SMEM_BYTES16C = SMEM_DOUBLES16C * sizeof(double);
enum { BAR_CC = 1 };

// canonical A_v[s, b0..] (column-major, true dims) -> private image: slices [chunk, b..] in L_A3 / L_Z2 / plain order, every
// leg zero-padded to 16; chunk = (re, im) of physical value s (complex) or the physical pair (2s, 2s + 1) (real)
template <bool CPLX>
__global__ void swizzle_sites_c16(const ItemDesc* items, int n_slots, const double* __restrict__ src, double* __restrict__ dst) {
  for (int it = blockIdx.x; it < n_slots; it += gridDim.x) {
    const ItemDesc d = items[it];
    if (d.kind < 0 || !d.first) continue;
    const int z = d.kind == 0 ? 3 : (d.kind == 1 ? 2 : 1);
    const int nsl = d.kind == 0 ? NSL3 : (d.kind == 1 ? NSL2 : NSL1);
    const int nb = nsl / 2;  // padded chunks per slice
    const double* s0 = src + (CPLX ? 2 : 1) * d.canon_off;
    double* d0 = dst + d.site_off;
    for (int c = threadIdx.x; c < nb * d.d; c += blockDim.x) {
      const int s = c % d.d, b = c / d.d;
      const int b0 = b & 15, b1 = (b >> 4) & 15, b2 = (b >> 8) & 15;
      uint32_t p;
      if (z == 3)
        p = pos<L_A3>(0, b0) ^ pos<L_A3>(1, b1) ^ pos<L_A3>(2, b2);
      else if (z == 2)
        p = pos<L_Z2>(0, b0) ^ pos<L_Z2>(1, b1);
      else
        p = 2 * b;
      double2 v = make_double2(0.0, 0.0);
      if (b0 < d.dim[0] && b1 < d.dim[1] && b2 < d.dim[2]) {
        const size_t lin = (size_t)d.phys * (b0 + (size_t)d.dim[0] * (b1 + (size_t)d.dim[1] * b2));
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
__device__ __forceinline__ void store_partial(double* mine, const double (&accr)[2][2][2], const double (&acci)[2][2][2], int g, int t) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int el = (g + 8 * mt) + CHI * (2 * t + i + 8 * h);  // out[v', v] at v' + 16 v
        *reinterpret_cast<double2*>(mine + 2 * el) =
            CPLX ? make_double2(accr[mt][h][i], acci[mt][h][i]) : make_double2(accr[mt][h][i] + acci[mt][h][i], 0.0);
      }
}

// thread el = b' + 16 b of a 16x16 tile -> offset inside the true chi x chi message, or -1 outside
__device__ __forceinline__ int msg_elem_off(int chi) {
  const int bp = threadIdx.x & 15, b = threadIdx.x >> 4;
  return (bp < chi && b < chi) ? bp + chi * b : -1;
}
template <bool CPLX>
__device__ __forceinline__ typename Tr<CPLX>::T load_old(const Args& k, const ItemDesc* d, int o) {
  using T = typename Tr<CPLX>::T;
  const int eo = msg_elem_off(d->out_dim[o]);
  return eo >= 0 ? reinterpret_cast<const T*>(k.msg_in)[d->out_off[o] + eo] : Elem<T>::zero();
}

// Block-wide epilogue (one element per thread and output): sum-normalise (beliefpropagation.jl:248-253), residual term
// 1 - |<old^, new^>|^2 (beliefpropagation.jl:261-267), store (+ peer store on cut edges).  `part` = 96 doubles of shared
// scratch.  A single warp would spend microseconds here on the eight complex divisions per lane -- on the critical path
// of a one-wave sweep -- so all 256 threads take one element each.
template <bool CPLX, int NOUT>
__device__ __forceinline__ void block_epilogue(const typename Tr<CPLX>::T (&v)[NOUT], const typename Tr<CPLX>::T (&old)[NOUT], const ItemDesc* d,
                                               const Args& k, double* part, int warp, int lane) {
  using T = typename Tr<CPLX>::T;
  using E = Elem<T>;
  double* part1 = part;        // [NCWC][2] (complex) sums
  double* part2 = part + 32;   // [NCWC][2][4] dot.re, dot.im, |old|^2, |new|^2
  int eo[NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; ++o) {
    eo[o] = msg_elem_off(d->out_dim[o]);
    const T s = warp_sum<T>(eo[o] >= 0 ? v[o] : E::zero());
    if (lane == 0) *reinterpret_cast<double2*>(part1 + (warp * 2 + o) * 2) = Tr<CPLX>::pack(s);
  }
  onchip::bar_sync(BAR_CC, NCTC);
#pragma unroll
  for (int o = 0; o < NOUT; ++o) {
    double2 ss = make_double2(0.0, 0.0);
#pragma unroll
    for (int w = 0; w < NCWC; ++w) {
      const double2 q = *reinterpret_cast<const double2*>(part1 + (w * 2 + o) * 2);
      ss.x += q.x;
      ss.y += q.y;
    }
    const T s = Tr<CPLX>::unpack(ss);
    const bool scale = k.normalize && !E::is_zero(s);
    T x = E::zero();
    if (eo[o] >= 0) {
      x = scale ? E::div(v[o], s) : v[o];
      const int64_t off = d->out_off[o] + eo[o];
      reinterpret_cast<T*>(k.msg_out)[off] = x;
      if (k.io.host_out) reinterpret_cast<T*>(k.io.host_out)[off] = x;
      if (k.peer.nranks > 1 && d->peer[o] >= 0) {
        reinterpret_cast<T*>(k.peer.peer_out[d->peer[o]])[off] = x;
        __threadfence_system();  // released here instead of at the kernel's tail
      }
    }
    const double2 dot = Tr<CPLX>::pack(warp_sum<T>(E::fma(E::conj(old[o]), x, E::zero())));
    const double n_old = warp_sum_d(E::abs2(old[o])), n_new = warp_sum_d(E::abs2(x));
    if (lane == 0) {
      double* q = part2 + (warp * 2 + o) * 4;
      q[0] = dot.x;
      q[1] = dot.y;
      q[2] = n_old;
      q[3] = n_new;
    }
  }
  onchip::bar_sync(BAR_CC, NCTC);
  if (threadIdx.x < NOUT) {
    const int o = threadIdx.x;
    double dr = 0, di = 0, n_old = 0, n_new = 0;
#pragma unroll
    for (int w = 0; w < NCWC; ++w) {
      const double* q = part2 + (w * 2 + o) * 4;
      dr += q[0];
      di += q[1];
      n_old += q[2];
      n_new += q[3];
    }
    residual_record(k.resmax, 1.0 - (dr * dr + di * di) / (n_old * n_new));
  }
}

// position of the next physical slice to load, tracked identically by every warp
struct Cursor {
  int idx, s;
};

template <bool CPLX>
__global__ void __launch_bounds__(NTHREADSC, 1) bp_update_onchip_c16x(Args k) {
  using T = typename Tr<CPLX>::T;
  const T* const msg_in_t = reinterpret_cast<const T*>(k.msg_in);
  extern __shared__ __align__(128) double smem[];
  double* Xbuf = smem + 2 * NSL3;
  double* red = Xbuf;  // alias: X is dead when the partial tiles are published
  double* raw = smem + 3 * NSL3;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(raw + 2 * CMSG);  // full[2]
  unsigned int* cnt = reinterpret_cast<unsigned int*>(mbar + 2);  // warps that released slot 0 / 1
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  if (sweep_already_converged(k.resmax, k.stop_key)) return;
  const int G = gridDim.x;

  // ---- slice ring without a producer warp (8 warps = 2 per scheduler, so every thread may hold 255 registers):
  // the LAST warp to release a slot refills it with the slice two units ahead ----
  auto valid = [&](const Cursor& c) { return c.idx < k.n_slots && k.items[c.idx].kind >= 0; };
  auto issue = [&](const Cursor& c, int sl) {  // one thread
    const ItemDesc* d = k.items + c.idx;
    const int nsl = d->kind == 0 ? NSL3 : (d->kind == 1 ? NSL2 : NSL1);
    mbar_expect_tx(&mbar[sl], nsl * 8);
    const double* src = k.sites + d->site_off + (size_t)c.s * nsl;
    if (d->kind == 0) {  // four 16 KiB copies in flight instead of one 64 KiB copy
#pragma unroll
      for (int q = 0; q < 4; ++q) tma_bulk_g2s(smem + sl * NSL3 + q * 2048, src + q * 2048, 16384, &mbar[sl]);
    } else {
      tma_bulk_g2s(smem + sl * NSL3, src, nsl * 8, &mbar[sl]);
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
  // the message fragments below may have been written by peers: wait for their posts of the previous sweep
  if (warp == 0) peer_gate(k.peer, lane);
  __syncthreads();

  auto release = [&](int sl) {  // warp-uniform; called when the warp is done with the slice in slot sl
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      if (atomicAdd(&cnt[sl], 1u) == NCWC - 1) {
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
#ifdef BPX_ONCHIP_TIMING
  if (lane == 0 && blockIdx.x == 0 && k.timing) k.timing[warp * 16] = clock64();
  if (threadIdx.x == 0 && k.timing) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    k.timing[2048 + 4 * blockIdx.x] = (long long)gt;
    k.timing[2048 + 4 * blockIdx.x + 2] = clock64();
  }
#endif
  for (int idx = blockIdx.x; idx < k.n_slots; idx += G) {
    const ItemDesc* d = k.items + idx;
    const int kind = d->kind;
    if (kind < 0) break;
    const int nd = d->d;
    if (k.io.progress) {  // streamed upload: the item's messages (fragments, old values) have arrived -- ONE warp polls
      if (warp == 0) hostio_wait(k.io, d->need);
      onchip::bar_sync(BAR_CC, NCTC);
    }
    if (kind == 0) {
      const int leg = d->leg;
      const CFrag m1 = load_cfrag<CPLX>(msg_in_t + d->in_off[0], g, t, d->in_dim[0]);
      const CFrag m2 = load_cfrag<CPLX>(msg_in_t + d->in_off[1], g, t, d->in_dim[1]);
      const T old[1] = {load_old<CPLX>(k, d, 0)};  // early: hides the miss
      double accr[2][2][2], acci[2][2][2];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) accr[a][b][0] = accr[a][b][1] = acci[a][b][0] = acci[a][b][1] = 0.0;
      TSTAMPC0(1);
      for (int s = 0; s < nd; ++s, ++u) {
        const int sl = u & 1;
        mbar_wait(&mbar[sl], (u >> 1) & 1);
        TSTAMPC(2);
        const double* A = smem + sl * NSL3;
        // X = A_s · M_first (columns: a spectator leg), then absorb M_second and close the output leg
        if (leg == 0) {
#pragma unroll 1
          for (int c = warp; c < 16; c += NCWC) absorb_one16c<CPLX, 2, 1>(A, Xbuf, pos<L_A3>(0, c), m1, g, t);
        } else {
#pragma unroll 1
          for (int c = warp; c < 16; c += NCWC) absorb_one16c<CPLX, 0, 1>(A, Xbuf, pos<L_A3>(2, c), m1, g, t);
        }
        TSTAMPC(3);
        onchip::bar_sync(BAR_CC, NCTC);
        TSTAMPC(4);
        if (leg == 0) {
#pragma unroll 1
          for (int c = warp; c < 16; c += NCWC) absorb_close16c<CPLX, L_A3, 1, 0, -1>(Xbuf, A, pos<L_A3>(2, c), m2, g, t, accr, acci);
        } else if (leg == 1) {
#pragma unroll 1
          for (int c = warp; c < 16; c += NCWC) absorb_close16c<CPLX, L_A3, 2, 1, -1>(Xbuf, A, pos<L_A3>(0, c), m2, g, t, accr, acci);
        } else {
#pragma unroll 1
          for (int c = warp; c < 16; c += NCWC) absorb_close16c<CPLX, L_A3, 1, 2, -1>(Xbuf, A, pos<L_A3>(0, c), m2, g, t, accr, acci);
        }
        TSTAMPC(5);
        release(sl);
        onchip::bar_sync(BAR_CC, NCTC);  // X is rewritten by the next slice / aliased by red
        TSTAMPC(6);
      }
      store_partial<CPLX>(red + warp * CMSG, accr, acci, g, t);
      onchip::bar_sync(BAR_CC, NCTC);
      TSTAMPC0(12);
      {
        const int el = threadIdx.x;
        double2 vs = make_double2(0.0, 0.0);
#pragma unroll
        for (int w = 0; w < NCWC; ++w) {
          const double2 q = *reinterpret_cast<const double2*>(red + w * CMSG + 2 * el);
          vs.x += q.x;
          vs.y += q.y;
        }
        const T v[1] = {Tr<CPLX>::unpack(vs)};
        block_epilogue<CPLX, 1>(v, old, d, k, raw, warp, lane);
        TSTAMPC0(13);
      }
    } else if (kind == 1) {
      // warps 0..3 = (output o, half tile hs): out0 absorbs leg 1 / closes leg 0 (M1), out1 absorbs leg 0 / closes leg 1 (M0)
      const int o = warp & 1, hs = (warp >> 1) & 1;
      double accr[2][2][2], acci[2][2][2];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) accr[a][b][0] = accr[a][b][1] = acci[a][b][0] = acci[a][b][1] = 0.0;
      CFrag m;
      if (warp < 4) m = load_cfrag<CPLX>(msg_in_t + d->in_off[1 - o], g, t, d->in_dim[1 - o]);
      const T old[2] = {load_old<CPLX>(k, d, 0), load_old<CPLX>(k, d, 1)};
      for (int s = 0; s < nd; ++s, ++u) {
        const int sl = u & 1;
        mbar_wait(&mbar[sl], (u >> 1) & 1);
        const double* A = smem + sl * NSL3;
        if (warp < 4) {
          if (o == 0) {
            if (hs == 0) absorb_close16c<CPLX, L_Z2, 1, 0, 0>(A, A, 0, m, g, t, accr, acci);
            else absorb_close16c<CPLX, L_Z2, 1, 0, 1>(A, A, 0, m, g, t, accr, acci);
          } else {
            if (hs == 0) absorb_close16c<CPLX, L_Z2, 0, 1, 0>(A, A, 0, m, g, t, accr, acci);
            else absorb_close16c<CPLX, L_Z2, 0, 1, 1>(A, A, 0, m, g, t, accr, acci);
          }
        }
        release(sl);
      }
      if (warp < 4) store_partial<CPLX>(red + warp * CMSG, accr, acci, g, t);
      onchip::bar_sync(BAR_CC, NCTC);
      {
        const int el = threadIdx.x;
        T v[2];
#pragma unroll
        for (int oo = 0; oo < 2; ++oo) {
          const double2 v0 = *reinterpret_cast<const double2*>(red + oo * CMSG + 2 * el);
          const double2 v1 = *reinterpret_cast<const double2*>(red + (oo + 2) * CMSG + 2 * el);
          v[oo] = Tr<CPLX>::unpack(make_double2(v0.x + v1.x, v0.y + v1.y));
        }
        block_epilogue<CPLX, 2>(v, old, d, k, raw, warp, lane);
      }
    } else {
      // degree 1: out[b', b] = sum_s A[s, b] conj(A[s, b']); thread el = b' + 16 b
      const int bp = threadIdx.x & 15, b = threadIdx.x >> 4;
      const T old[1] = {load_old<CPLX>(k, d, 0)};
      double sr = 0, si = 0;
      for (int s = 0; s < nd; ++s, ++u) {
        const int sl = u & 1;
        mbar_wait(&mbar[sl], (u >> 1) & 1);
        const double* A = smem + sl * NSL3;
        const double2 x = *reinterpret_cast<const double2*>(A + 2 * b), y = *reinterpret_cast<const double2*>(A + 2 * bp);
        sr += x.x * y.x + x.y * y.y;              // complex: Re(x conj(y)); real: both physical values of the chunk
        if (CPLX) si += x.y * y.x - x.x * y.y;
        release(sl);
      }
      const T v[1] = {Tr<CPLX>::unpack(make_double2(sr, si))};
      block_epilogue<CPLX, 1>(v, old, d, k, raw, warp, lane);
    }
  }
#ifdef BPX_ONCHIP_TIMING
  if (threadIdx.x == 0 && k.timing) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    k.timing[2048 + 4 * blockIdx.x + 1] = (long long)gt;
    k.timing[2048 + 4 * blockIdx.x + 3] = clock64();
  }
#endif
  peer_post_when_last(k.peer, false);  // peer stores were released where they were issued
  hostio_finish(k.io);
}

}  // namespace onchip16c
}  // namespace bpx
