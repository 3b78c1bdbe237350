// BPX_KERNEL_ONCHIP: vertex-centric update kernel for buckets whose site tensor fits in shared memory
// (chi = 8, d = 2, Float64, degree 2..4: at most 8192 doubles = 64 KiB) -- BASELINE config 2.
//
// One persistent CTA per SM loops over a cost-sorted list of vertices (degree 4 first, then 3, then 2; the
// cheap boundary vertices fill the last, partially empty wave).  For a vertex u it produces ALL outgoing
// messages from ONE read of A_u with a leave-one-out tree, e.g. for degree 4
//     P = A·M0·M1 ;  out3 = close_3(P·M2) ; out2 = close_2(P·M3)
//     Q = A·M2·M3 ;  out1 = close_1(Q·M0) ; out0 = close_0(Q·M1)
// (8 absorptions + 4 closures = 12 GEMM units of d·chi^5 MACs instead of 16 for four independent
// updates of beliefpropagation.jl:242-257).  Every unit runs on the FP64 tensor pipe
// (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4); tcgen05 has no FP64 kind.
//
// Register chaining.  The accumulator fragment of one DMMA (row g = lane/4, cols 2t, 2t+1, t = lane%4) is
// a valid A or B operand fragment of the next one when the contraction index is enumerated as
// k <-> 2t+i.  Hence  "absorb leg X, then absorb leg Y"  and  "absorb leg U, then close over V with
// conj(A)"  each run as two back-to-back DMMA groups without staging the intermediate: only P / Q are
// materialised (in shared memory), T = P·M is never stored.
//
// Shared-memory layout.  Both physical values s = 0,1 of one (a0..a3) sit in one 16-byte chunk, so one
// LDS.128 / STS.128 feeds the two DMMA chains of s = 0 and s = 1.  The chunk position is an XOR swizzle
// of the canonical index, linear over GF(2), chosen such that every fragment access pattern used below
// (g <-> one leg, t <-> two bits of another leg; leg pairs (0,1), (2,3), (1,2), (0,3)) touches 8 distinct
// 16-byte bank groups per quarter warp: conflict free.  The library keeps a private, PRE-SWIZZLED HBM image of
// every on-chip site tensor (built once per upload by swizzle_sites), so the whole 64 KiB tensor and its
// incoming messages arrive by TMA bulk copies (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP) issued by
// one thread while the previous vertex is being computed -- no LSU / issue-slot cost for the copy.
//
// Warp roles.  16 compute warps issue the DMMA chains and reduce the per-warp partial 8x8 tiles; two more
// warps own the epilogue (sum-normalisation, residual, store), handed over through named barriers so the
// compute warps never wait for the division / global-memory latency of the epilogue.
#pragma once
#include "bpx_common.cuh"
#include "bpx_peer.cuh"

namespace bpx {
namespace onchip {

constexpr int CHI = 8;
constexpr int NELEM = 2 * CHI * CHI * CHI * CHI;  // 8192 doubles (degree 4)
constexpr int NCW = 16;                            // compute warps
constexpr int NCT = NCW * 32;                      // compute threads
constexpr int NEW = 2;                             // epilogue warps (one per tile of a branch)
constexpr int NTHREADS = NCT + 32 * NEW + 32;      // + TMA producer warp
constexpr int MSG = CHI * CHI;

enum { BAR_PHASE = 1, BAR_RED = 2, BAR_RAW_FULL = 3, BAR_RAW_FREE = 4, BAR_SLOT_FREE = 5 /* and 6 */ };
constexpr int NRAW = NCT + 32 * NEW;  // participants of the raw-tile hand-over

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }

// per work item, built on the host (fast_prepare): everything the kernel needs without pointer chasing
struct ItemDesc {
  int64_t site_off;    // elements
  int64_t in_off[4];   // message offsets (elements) of the message arriving on leg i
  int64_t out_off[4];  // message offsets of the outgoing message on leg i
  int32_t out_edge[4];
  int32_t z;
  int32_t branch;  // degree 4 only: 0 = branch P (out3, out2), 1 = branch Q (out1, out0); each is its own work item
  int32_t pad[2];
  int32_t peer[4];  // rank that owns the head of out-edge i when it lives on another rank (cut edge), else -1
  int64_t need;     // streamed host I/O: message-set prefix (elements) that holds every message this item reads
};

// ---- swizzled position (in doubles, s = 0) of element (a0,a1,a2,a3); XOR-linear in every index bit ----
// bit 0: s | bits 1-3: bank group | bits 4-12: a0[2], a1[1], a1[2], a2[0..2], a3[0..2]
__device__ __forceinline__ uint32_t leg_pos(int leg, uint32_t a) {
  const uint32_t x02 = (a ^ (a >> 2)) & 1u, b1 = (a >> 1) & 1u, b2 = (a >> 2) & 1u;
  switch (leg) {
    case 0: return (x02 << 1) | (b1 << 2) | (b2 << 4);
    case 1: return (b1 << 2) | (x02 << 3) | (b1 << 5) | (b2 << 6);
    case 2: return (x02 << 1) | (b1 << 2) | (a << 7);
    case 3: return (b1 << 2) | (x02 << 3) | (a << 10);
    default: return 0;  // leg -1: absent
  }
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- mbarrier + TMA bulk copy (global -> shared::cta) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// message fragments (M[bra, ket] column-major: (a', a) at a' + CHI*a), read from the staged copy
struct MsgFrag {
  double ma[2];  // M[g, t + 4j]   : A operand of "absorb first leg", B operand of the T-GEMM
  double mb[2];  // M[g, 2t + i]   : B operand of "absorb second leg" (register-chained)
};
__device__ __forceinline__ MsgFrag load_frag(const double* M, int g, int t) {
  MsgFrag f;
  f.ma[0] = M[g + CHI * t];
  f.ma[1] = M[g + CHI * (t + 4)];
  f.mb[0] = M[g + CHI * (2 * t)];
  f.mb[1] = M[g + CHI * (2 * t + 1)];
  return f;
}

// columns: the index values of up to two "spectator" legs C0, C1 (leg -1: absent)
template <int C0, int C1>
__device__ __forceinline__ uint32_t col_pos(int col) {
  return leg_pos(C0, col & 7) ^ leg_pos(C1, col >> 3);
}
template <int C0, int C1>
__device__ __forceinline__ constexpr int n_cols() {
  return (C0 >= 0 ? CHI : 1) * (C1 >= 0 ? CHI : 1);
}

// dst[.., x', y', ..] = sum_{x,y} MX[x', x] MY[y', y] src[.., x, y, ..]      (legs X then Y absorbed)
template <int X, int Y, int C0, int C1>
__device__ __forceinline__ void absorb_pair(const double* src, double* dst, const MsgFrag& mx, const MsgFrag& my, int warp, int g,
                                            int t) {
  const uint32_t ld0 = leg_pos(X, t) ^ leg_pos(Y, g), ld1 = leg_pos(X, t + 4) ^ leg_pos(Y, g);
  const uint32_t st0 = leg_pos(X, g) ^ leg_pos(Y, 2 * t), st1 = leg_pos(X, g) ^ leg_pos(Y, 2 * t + 1);
#pragma unroll 2
  for (int col = warp; col < n_cols<C0, C1>(); col += NCW) {
    const uint32_t base = col_pos<C0, C1>(col);
    const double2 b0 = *reinterpret_cast<const double2*>(src + (base ^ ld0));
    const double2 b1 = *reinterpret_cast<const double2*>(src + (base ^ ld1));
    // absorb X: D[x' = g, y = 2t+i] = sum_x MX[x', x] src[x, y]      (one chain per physical value s)
    double xa0 = 0, xa1 = 0, xb0 = 0, xb1 = 0;
    dmma(xa0, xa1, mx.ma[0], b0.x);
    dmma(xb0, xb1, mx.ma[0], b0.y);
    dmma(xa0, xa1, mx.ma[1], b1.x);
    dmma(xb0, xb1, mx.ma[1], b1.y);
    // absorb Y from registers: D[x' = g, y' = 2t+i] = sum_{y = 2t+i} D1[x', y] MY[y', y]
    double pa0 = 0, pa1 = 0, pb0 = 0, pb1 = 0;
    dmma(pa0, pa1, xa0, my.mb[0]);
    dmma(pb0, pb1, xb0, my.mb[0]);
    dmma(pa0, pa1, xa1, my.mb[1]);
    dmma(pb0, pb1, xb1, my.mb[1]);
    *reinterpret_cast<double2*>(dst + (base ^ st0)) = make_double2(pa0, pb0);
    *reinterpret_cast<double2*>(dst + (base ^ st1)) = make_double2(pa1, pb1);
  }
}

// dst[.., x', y, ..] = sum_x MX[x', x] src[.., x, y, ..]      (single absorption; Y is a passive tile leg)
template <int X, int Y, int C0, int C1>
__device__ __forceinline__ void absorb_one(const double* src, double* dst, const MsgFrag& mx, int warp, int g, int t) {
  const uint32_t ld0 = leg_pos(X, t) ^ leg_pos(Y, g), ld1 = leg_pos(X, t + 4) ^ leg_pos(Y, g);
  const uint32_t st0 = leg_pos(X, g) ^ leg_pos(Y, 2 * t), st1 = leg_pos(X, g) ^ leg_pos(Y, 2 * t + 1);
  for (int col = warp; col < n_cols<C0, C1>(); col += NCW) {
    const uint32_t base = col_pos<C0, C1>(col);
    const double2 b0 = *reinterpret_cast<const double2*>(src + (base ^ ld0));
    const double2 b1 = *reinterpret_cast<const double2*>(src + (base ^ ld1));
    double xa0 = 0, xa1 = 0, xb0 = 0, xb1 = 0;
    dmma(xa0, xa1, mx.ma[0], b0.x);
    dmma(xb0, xb1, mx.ma[0], b0.y);
    dmma(xa0, xa1, mx.ma[1], b1.x);
    dmma(xb0, xb1, mx.ma[1], b1.y);
    *reinterpret_cast<double2*>(dst + (base ^ st0)) = make_double2(xa0, xb0);
    *reinterpret_cast<double2*>(dst + (base ^ st1)) = make_double2(xa1, xb1);
  }
}

// out[v', v] = sum_{s, cols, u'} conj(A[.., u', v']) * ( sum_u MU[u', u] P[.., u, v] )
// (leg U absorbed on the fly, leg V left open).  Returns this warp's partial 8x8 tile: thread (g, t) holds
// out[v' = g, v = 2t], out[v' = g, v = 2t + 1].
template <int U, int V, int C0, int C1>
__device__ __forceinline__ void absorb_close(const double* P, const double* A, const MsgFrag& mu, int warp, int g, int t, double& o0,
                                             double& o1) {
  const uint32_t lp0 = leg_pos(U, t) ^ leg_pos(V, g), lp1 = leg_pos(U, t + 4) ^ leg_pos(V, g);
  const uint32_t la0 = leg_pos(U, 2 * t) ^ leg_pos(V, g), la1 = leg_pos(U, 2 * t + 1) ^ leg_pos(V, g);
  double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
#pragma unroll 2
  for (int col = warp; col < n_cols<C0, C1>(); col += NCW) {
    const uint32_t base = col_pos<C0, C1>(col);
    const double2 p0 = *reinterpret_cast<const double2*>(P + (base ^ lp0));
    const double2 p1 = *reinterpret_cast<const double2*>(P + (base ^ lp1));
    const double2 a0 = *reinterpret_cast<const double2*>(A + (base ^ la0));
    const double2 a1 = *reinterpret_cast<const double2*>(A + (base ^ la1));
    // T[v = g, u' = 2t+i] = sum_u P[u, v] MU[u', u]
    double ta0 = 0, ta1 = 0, tb0 = 0, tb1 = 0;
    dmma(ta0, ta1, p0.x, mu.ma[0]);
    dmma(tb0, tb1, p0.y, mu.ma[0]);
    dmma(ta0, ta1, p1.x, mu.ma[1]);
    dmma(tb0, tb1, p1.y, mu.ma[1]);
    // out[v' = g, v] += sum_{u' = 2t+i} A[u', v'] T[v, u']      (T re-used as B operand from registers)
    dmma(acc[0][0], acc[0][1], a0.x, ta0);
    dmma(acc[1][0], acc[1][1], a0.y, tb0);
    dmma(acc[2][0], acc[2][1], a1.x, ta1);
    dmma(acc[3][0], acc[3][1], a1.y, tb1);
  }
  o0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
  o1 = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
}

#ifdef BPX_ONCHIP_TIMING
__device__ __forceinline__ long long gtimer() {
  unsigned long long gt;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
  return (long long)gt;
}
#define GSTAMP(i) do { if (lane == 0 && k.timing) k.timing[2048 + 8 * blockIdx.x + (i)] = gtimer(); } while (0)
#else
#define GSTAMP(i) do { } while (0)
#endif
#ifdef BPX_ONCHIP_TIMING
#define TSTAMP(i) do { if (lane == 0 && blockIdx.x == 0 && k.timing && n_iter < 8) k.timing[(n_iter * 32 + warp) * 16 + (i)] = clock64(); } while (0)
#else
#define TSTAMP(i) do { } while (0)
#endif

struct Args {
  long long* timing;
  const ItemDesc* items;
  int n_items;
  const double* sites;  // PRE-SWIZZLED image (same offsets as the canonical buffer)
  const double* msg_in;
  double* msg_out;
  double* residual;              // per-edge residual terms (may be NULL)
  unsigned long long* resmax;    // this sweep's residual key (atomicMax)
  int normalize;
  PeerArgs peer;                 // multi-GPU: gate / direct peer stores / post (nranks <= 1: unused)
  HostIO io;                     // streamed host I/O (bpx_sweep_host), all NULL otherwise
  unsigned long long stop_key;   // device-side convergence test (sweep_already_converged), 0: none
};

// shared memory (doubles): A[2][NELEM] | P[NELEM] | red[2][NCW][MSG] | raw[2][MSG] | msgs[2][4][MSG] | 2 mbarriers
constexpr size_t SMEM_DOUBLES = (size_t)3 * NELEM + 2 * NCW * MSG + 2 * MSG + 2 * 4 * MSG + 2;
constexpr size_t SMEM_BYTES = SMEM_DOUBLES * sizeof(double);

// conflict-free position of tile element el = v' + 8 v inside a 64-element partial tile
__device__ __forceinline__ int red_pos(int el) { return el ^ (((el >> 4) & 3) << 2); }

// one thread: TMA the pre-swizzled tensor and the incoming messages of an item into buffer slot `slot`
__device__ __forceinline__ void tma_item(double* Adst, double* Mdst, uint64_t* bar, const Args& k, const ItemDesc* d) {
  const int z = d->z;
  const uint32_t abytes = (16u << (3 * z));  // 2 * 8^z doubles
  fence_proxy_async();  // generic-proxy reads of this slot (previous tenant) happen-before the async-proxy writes
  mbar_expect_tx(bar, abytes + z * MSG * 8);
  const char* src = reinterpret_cast<const char*>(k.sites + d->site_off);
  for (uint32_t off = 0; off < abytes; off += 16384) {
    const uint32_t n = abytes - off < 16384 ? abytes - off : 16384;
    tma_bulk_g2s(reinterpret_cast<char*>(Adst) + off, src + off, n, bar);
  }
  for (int i = 0; i < z; ++i) tma_bulk_g2s(Mdst + i * MSG, k.msg_in + d->in_off[i], MSG * 8, bar);
}

// compute warps: reduce the 16 per-warp partial tiles of a branch (two 8x8 outputs) and hand them over
__device__ __forceinline__ void publish(double* red, double* raw, int warp, int lane, int g, int t, double a0, double a1, double b0,
                                        double b1) {
  double* r0 = red + warp * MSG;
  double* r1 = red + (NCW + warp) * MSG;
  r0[red_pos(g + CHI * (2 * t))] = a0;
  r0[red_pos(g + CHI * (2 * t + 1))] = a1;
  r1[red_pos(g + CHI * (2 * t))] = b0;
  r1[red_pos(g + CHI * (2 * t + 1))] = b1;
  bar_sync(BAR_RED, NCT);  // partial tiles visible; every compute warp is done reading P and A of this branch
  // warp w reduces elements 8w'..8w'+7 of tile (w / 8): lane = 8 * q + e sums partial warps 4q..4q+3 of element e
  const int tile = warp >> 3, el = (warp & 7) * 8 + (lane & 7), q = lane >> 3;
  const double* r = red + tile * NCW * MSG + red_pos(el);
  double s = (r[(4 * q) * MSG] + r[(4 * q + 1) * MSG]) + (r[(4 * q + 2) * MSG] + r[(4 * q + 3) * MSG]);
  s += __shfl_xor_sync(0xffffffffu, s, 8);
  s += __shfl_xor_sync(0xffffffffu, s, 16);
  bar_sync(BAR_RAW_FREE, NRAW);  // epilogue warps are done with the previous branch's raw tiles
  if (lane < 8) raw[tile * MSG + el] = s;
  bar_arrive(BAR_RAW_FULL, NRAW);
}

// epilogue warp: sum-normalise (beliefpropagation.jl:248-253), residual term (beliefpropagation.jl:261-267), store.
// Lane holds elements lane and lane + 32.  The residual 1 - |<old^, new^>|^2 is invariant under the scaling,
// so all four reductions run interleaved on the raw tile.
__device__ __forceinline__ void epilogue_tile(double v0, double v1, double o0, double o1, int lane, double* new_m, int normalize,
                                              double* residual_slot, unsigned long long* resmax, double* peer_m, double* host_m = nullptr) {
  double s = v0 + v1, dot = o0 * v0 + o1 * v1, n_old = o0 * o0 + o1 * o1, n_new = v0 * v0 + v1 * v1;
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, m);
    dot += __shfl_xor_sync(0xffffffffu, dot, m);
    n_old += __shfl_xor_sync(0xffffffffu, n_old, m);
    n_new += __shfl_xor_sync(0xffffffffu, n_new, m);
  }
  if (normalize && s != 0.0) {
    v0 /= s;
    v1 /= s;
  }
  new_m[lane] = v0;
  new_m[lane + 32] = v1;
  if (peer_m) {  // cut edge: the owner of the head reads this message next sweep -- store it there too (NVLink)
    peer_m[lane] = v0;
    peer_m[lane + 32] = v1;
    __threadfence_system();  // released here, by the (otherwise idle) epilogue warp, instead of at the kernel's tail
  }
  if (host_m) {  // streamed host I/O: the caller's host buffer (mapped), posted writes over PCIe
    host_m[lane] = v0;
    host_m[lane + 32] = v1;
  }
  if (lane == 0) {
    const double r = 1.0 - dot * dot / (n_old * n_new);
    if (residual_slot) *residual_slot = r;
    residual_record(resmax, r);
  }
}

// Build the pre-swizzled image: dst[site_off + pos(c)] = src[site_off + 2c .. 2c+1] for every 16-byte chunk c.
__global__ void swizzle_sites(const ItemDesc* items, int n_items, const double* __restrict__ src, double* __restrict__ dst) {
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int64_t off = items[item].site_off;
    const int nchunks = 1 << (3 * items[item].z);
    for (int c = threadIdx.x; c < nchunks; c += blockDim.x) {
      const uint32_t pos = leg_pos(0, c & 7) ^ leg_pos(1, (c >> 3) & 7) ^ leg_pos(2, (c >> 6) & 7) ^ leg_pos(3, c >> 9);
      const double2 v = *reinterpret_cast<const double2*>(src + off + 2 * c);
      *reinterpret_cast<double2*>(dst + off + pos) = v;
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) bp_update_onchip_c8(Args k) {
  extern __shared__ __align__(128) double smem[];
  double* Pbuf = smem + 2 * NELEM;
  double* red = smem + 3 * NELEM;
  double* raw = red + 2 * NCW * MSG;
  double* msgs = raw + 2 * MSG;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(msgs + 2 * 4 * MSG);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int G = gridDim.x;
  if (sweep_already_converged(k.resmax, k.stop_key)) return;
  if ((int)blockIdx.x >= k.n_items) return;

  if (threadIdx.x == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp == NCW + NEW) {
    // ================= producer warp: TMA of item n into slot n & 1, two items ahead of the compute warps ==========
    GSTAMP(0);
    peer_gate(k.peer, lane);  // multi-GPU: the peers' cut-edge messages of the previous sweep have landed
    GSTAMP(1);
    int n = 0;
    for (int item = blockIdx.x; item < k.n_items; item += G, ++n) {
      if (n >= 2) bar_sync(BAR_SLOT_FREE + (n & 1), NCT + 32);  // compute warps are done with the slot's previous tenant
      hostio_wait(k.io, k.items[item].need);  // streamed upload: the item's messages have arrived
      if (lane == 0) tma_item(smem + (n & 1) * NELEM, msgs + (n & 1) * 4 * MSG, &mbar[n & 1], k, k.items + item);
    }
  } else if (warp >= NCW) {

    // ================= epilogue warps: warp NCW + i owns tile i of every branch =================
    const int which = warp - NCW;
    bar_arrive(BAR_RAW_FREE, NRAW);  // raw starts free
    for (int item = blockIdx.x; item < k.n_items; item += G) {
      const ItemDesc* d = k.items + item;
      const int z = d->z, br = d->branch;
      const int nb = z == 3 ? 2 : 1;
      for (int b = 0; b < nb; ++b) {
        // branch list mirrors the compute warps: (leg of tile 0, leg of tile 1 or -1)
        int l0, l1;
        if (z == 4) { l0 = br == 0 ? 3 : 1; l1 = br == 0 ? 2 : 0; }
        else if (z == 3) { l0 = b == 0 ? 2 : 0; l1 = b == 0 ? 1 : -1; }
        else { l0 = 1; l1 = 0; }
        const int l = which == 0 ? l0 : l1;
        double o0 = 0, o1 = 0;
        int64_t off = 0;
        int e = 0;
        double* peer_m = nullptr;
        if (l >= 0) {  // old message: issued before the hand-over so its latency overlaps the compute
          hostio_wait(k.io, d->need);
          off = d->out_off[l];
          e = d->out_edge[l];
          if (k.peer.nranks > 1 && d->peer[l] >= 0) peer_m = k.peer.peer_out[d->peer[l]] + off;
          o0 = k.msg_in[off + lane];
          o1 = k.msg_in[off + lane + 32];
        }
        bar_sync(BAR_RAW_FULL, NRAW);
        const double v0 = raw[which * MSG + lane], v1 = raw[which * MSG + lane + 32];
        bar_arrive(BAR_RAW_FREE, NRAW);  // values are in registers: raw may be overwritten
        if (l >= 0)
          epilogue_tile(v0, v1, o0, o1, lane, k.msg_out + off, k.normalize, k.residual ? k.residual + e : nullptr, k.resmax, peer_m,
                        k.io.host_out ? k.io.host_out + off : nullptr);
      }
    }
  } else {
  // ================= compute warps =================
  int n_iter = 0;
  for (int item = blockIdx.x; item < k.n_items; item += G, ++n_iter) {
    const int cur = n_iter & 1;
    const ItemDesc* d = k.items + item;
    const int z = d->z, br = d->branch;  // L1/L2 hits; consumed only after the mbarrier wait
    TSTAMP(0);
    mbar_wait(&mbar[cur], (n_iter >> 1) & 1);  // A_u and its messages landed
    TSTAMP(2);
    const double* A = smem + cur * NELEM;
    const double* M = msgs + cur * 4 * MSG;
    double a0, a1, b0 = 0, b1 = 0;
    if (z == 4) {
      if (br == 0) {
        // branch P = A·M0·M1  ->  out3 (absorb 2, close 3), out2 (absorb 3, close 2)
        const MsgFrag m0 = load_frag(M, g, t), m1 = load_frag(M + MSG, g, t), m2 = load_frag(M + 2 * MSG, g, t),
                      m3 = load_frag(M + 3 * MSG, g, t);
        TSTAMP(3);
        absorb_pair<0, 1, 2, 3>(A, Pbuf, m0, m1, warp, g, t);
        TSTAMP(4);
        bar_sync(BAR_PHASE, NCT);
        TSTAMP(5);
        absorb_close<2, 3, 0, 1>(Pbuf, A, m2, warp, g, t, a0, a1);
        TSTAMP(6);
        absorb_close<3, 2, 0, 1>(Pbuf, A, m3, warp, g, t, b0, b1);
        TSTAMP(7);
      } else {
        // branch Q = A·M2·M3  ->  out1 (absorb 0, close 1), out0 (absorb 1, close 0)
        const MsgFrag m0 = load_frag(M, g, t), m1 = load_frag(M + MSG, g, t), m2 = load_frag(M + 2 * MSG, g, t),
                      m3 = load_frag(M + 3 * MSG, g, t);
        TSTAMP(3);
        absorb_pair<2, 3, 0, 1>(A, Pbuf, m2, m3, warp, g, t);
        TSTAMP(4);
        bar_sync(BAR_PHASE, NCT);
        TSTAMP(5);
        absorb_close<0, 1, 2, 3>(Pbuf, A, m0, warp, g, t, a0, a1);
        TSTAMP(6);
        absorb_close<1, 0, 2, 3>(Pbuf, A, m1, warp, g, t, b0, b1);
        TSTAMP(7);
      }
      publish(red, raw, warp, lane, g, t, a0, a1, b0, b1);
      TSTAMP(8);
    } else if (z == 3) {
      const MsgFrag m0 = load_frag(M, g, t), m1 = load_frag(M + MSG, g, t), m2 = load_frag(M + 2 * MSG, g, t);
      // X = A·M0  ->  out2 (absorb 1, close 2), out1 (absorb 2, close 1)
      absorb_one<0, 1, 2, -1>(A, Pbuf, m0, warp, g, t);
      bar_sync(BAR_PHASE, NCT);
      absorb_close<1, 2, 0, -1>(Pbuf, A, m1, warp, g, t, a0, a1);
      absorb_close<2, 1, 0, -1>(Pbuf, A, m2, warp, g, t, b0, b1);
      publish(red, raw, warp, lane, g, t, a0, a1, b0, b1);
      // X = A·M2  ->  out0 (absorb 1, close 0)
      absorb_one<2, 1, 0, -1>(A, Pbuf, m2, warp, g, t);
      bar_sync(BAR_PHASE, NCT);
      absorb_close<1, 0, 2, -1>(Pbuf, A, m1, warp, g, t, a0, a1);
      publish(red, raw, warp, lane, g, t, a0, a1, 0.0, 0.0);
    } else {  // z == 2: out1 (absorb 0, close 1), out0 (absorb 1, close 0) straight from A
      const MsgFrag m0 = load_frag(M, g, t), m1 = load_frag(M + MSG, g, t);
      absorb_close<0, 1, -1, -1>(A, A, m0, warp, g, t, a0, a1);
      absorb_close<1, 0, -1, -1>(A, A, m1, warp, g, t, b0, b1);
      publish(red, raw, warp, lane, g, t, a0, a1, b0, b1);
    }
    // every compute warp passed the item's last BAR_RED: nobody reads slot `cur` any more
    if (item + 2 * G < k.n_items) bar_arrive(BAR_SLOT_FREE + cur, NCT + 32);
  }
  // let the epilogue warps' last arrive complete
  bar_sync(BAR_RAW_FREE, NRAW);
  if (warp == 0) GSTAMP(2);
  }
  // multi-GPU: every role of this CTA is done; the last CTA posts (sweep id, local residual) to all ranks
  if (warp == NCW) GSTAMP(3);  // epilogue warp 0 done
  peer_post_when_last(k.peer, false);  // peer stores were released where they were issued
  if (warp == 0) GSTAMP(4);
  hostio_finish(k.io);
}

}  // namespace onchip
}  // namespace bpx
