// BPX_KERNEL_SLICED: vertex-centric update kernel for (degree 4, chi = 16, d = 2, Float64) -- BASELINE config 5.
//
// A_u[s, a0..a3] is 131072 doubles = 1 MiB: it cannot live in one SM's shared memory, so a work item
// (vertex, branch) streams it in 64 KiB LEG SLICES through a TMA ring while one CTA keeps the tensor pipe busy:
//
//   branch P (out3, out2)                                   branch Q (out1, out0)
//   phase 1: for r: A[.., a3 = r]  --absorb M0, M1 in place--> P[.., a3 = r]   -> scratch      (slices of leg 0)
//   phase 2: for r, half: P[a0' = r, ..], A[a0 = r, ..] --absorb 2 / close 3, absorb 3 / close 2--> accumulators
//
// The absorbed pair must be whole inside a slice in phase 1 (slice on the other pair), the closed pair must be
// whole in phase 2 (slice on the absorbed pair), so P is re-sliced in between: it round-trips through a per-CTA
// 1 MiB scratch that stays L2 resident.  Only legs 0 and 3 are ever sliced; the private HBM image of A and P
// puts (a3, a0) in the top address bits and XOR-swizzles the 16-byte chunks exactly like the on-chip kernel, so
//   * an a3-slice is one contiguous 64 KiB run, an a0-slice is sixteen 4 KiB runs (TMA bulk copies, no LSU),
//   * every DMMA fragment access in shared memory is an LDS.128/STS.128 without bank conflicts,
//   * phase 1 runs IN PLACE (each warp owns whole columns), so a 3-buffer ring overlaps load / compute / store.
// Same register-chained DMMA groups as bpx_onchip.cuh (absorb-absorb, absorb-close), with 16-wide legs:
// 2x2 tiles of m8n8, 4 k-steps.  Closure accumulators (two 16x16 outputs) stay in registers for the whole item.
#pragma once
#include "bpx_common.cuh"
#include "bpx_onchip.cuh"
#include "bpx_peer.cuh"

namespace bpx {
namespace sliced {

using onchip::dmma;
using onchip::fence_proxy_async;
using onchip::mbar_expect_tx;
using onchip::mbar_init;
using onchip::mbar_wait;
using onchip::smem_u32;
using onchip::tma_bulk_g2s;

constexpr int CHI = 16;
constexpr int MSG = CHI * CHI;             // 256
constexpr int NTENSOR = 2 * 16 * 16 * 16 * 16;  // 131072 doubles
constexpr int SLICE = 8192;                // doubles per 64 KiB slice
constexpr int HALF = 4096;
constexpr int NCW = 8;                     // compute warps
constexpr int NCT = NCW * 32;
constexpr int NTHREADS = NCT + 32;         // + producer warp

// ---- private image: bit 0 s | 1-3 bank group | 4 a1[2] | 5-7 a2[1..3] | 8 a1[3] | 9-12 a0 | 13-16 a3 ----
__device__ __host__ __forceinline__ uint32_t grp_bits(int leg, uint32_t a) {
  const uint32_t x02 = (a ^ (a >> 2)) & 1u, b1 = (a >> 1) & 1u;
  return (leg == 0 || leg == 2) ? ((x02 << 1) | (b1 << 2)) : ((b1 << 2) | (x02 << 3));
}
__device__ __host__ __forceinline__ uint32_t global_pos(int leg, uint32_t a) {
  const uint32_t g = grp_bits(leg, a);
  switch (leg) {
    case 0: return g | (a << 9);
    case 1: return g | (((a >> 2) & 1u) << 4) | (((a >> 3) & 1u) << 8);
    case 2: return g | ((a >> 1) << 5);
    default: return g | (a << 13);
  }
}
// shared-memory layouts of the streamed pieces
enum { L_A3 = 0,   // full a3-slice: global bits 0..12
       L_A0 = 1,   // full a0-slice: (a3 << 9) | global bits 0..8
       L_A0H = 2,  // a0-slice, a1[3] fixed: (a3 << 8) | global bits 0..7
       L_A3H = 3,  // a3-slice, a2[3] fixed: (a0 << 8) | (a1[3] << 7) | global bits 0..6
       L_Z2 = 4    // two legs only (512 doubles): bit 0 | 1-3 bank group (^ a1[3] on bit 1) | 4 a1[2] | 5-8 a0
};
template <int LAY>
__device__ __forceinline__ uint32_t pos(int leg, uint32_t a) {
  const uint32_t g = grp_bits(leg, a);
  if (LAY == L_A3) {
    switch (leg) {
      case 0: return g | (a << 9);
      case 1: return g | (((a >> 2) & 1u) << 4) | (((a >> 3) & 1u) << 8);
      case 2: return g | ((a >> 1) << 5);
      default: return g;
    }
  } else if (LAY == L_A0) {
    switch (leg) {
      case 3: return g | (a << 9);
      case 1: return g | (((a >> 2) & 1u) << 4) | (((a >> 3) & 1u) << 8);
      case 2: return g | ((a >> 1) << 5);
      default: return g;
    }
  } else if (LAY == L_A0H) {
    switch (leg) {
      case 3: return g | (a << 8);
      case 1: return g | (((a >> 2) & 1u) << 4);
      case 2: return g | ((a >> 1) << 5);
      default: return g;
    }
  } else if (LAY == L_A3H) {
    switch (leg) {
      case 0: return g | (a << 8);
      case 1: return g | (((a >> 2) & 1u) << 4) | (((a >> 3) & 1u) << 7);
      case 2: return g | (((a >> 1) & 3u) << 5);
      default: return g;
    }
  } else {
    switch (leg) {
      case 0: return g | (a << 5);
      case 1: return g | (((a >> 2) & 1u) << 4) | (((a >> 3) & 1u) << 1);
      default: return 0;
    }
  }
}

struct ItemDesc {
  int64_t site_off;    // elements, into the private image buffer
  int64_t in_off[4];
  int64_t out_off[4];
  int32_t out_edge[4];
  int32_t branch;      // 0: P (out3, out2), 1: Q (out1, out0)
  int32_t pad[3];
  int32_t peer[4];     // rank owning the head of out-edge i if it lives elsewhere (cut edge), else -1
  int64_t need;        // streamed host I/O: prefix of the upload that holds every message this item reads
};

struct Args {
  const ItemDesc* items;
  int n_items;
  const double* sites;  // private swizzled image
  double* scratch;      // NTENSOR doubles per CTA
  const double* msg_in;
  double* msg_out;
  double* residual;
  unsigned long long* resmax;  // this sweep's residual key (atomicMax)
  int normalize;
  PeerArgs peer;               // multi-GPU: gate / direct peer stores / post (nranks <= 1: unused)
  HostIO io;                   // streamed host I/O (bpx_sweep_host), all NULL otherwise
  unsigned long long stop_key; // device-side convergence test (sweep_already_converged), 0: none
};

// message fragments for 16-wide legs
struct FragA {  // M[g + 8 mt, t + 4 j]: A operand of the first absorption / B operand of the T-GEMM
  double v[2][4];
};
struct FragB {  // M[g + 8 nt, 2 t + i + 8 h]: B operand of the register-chained second absorption
  double v[2][2][2];
};
__device__ __forceinline__ FragA load_fragA(const double* __restrict__ M, int g, int t) {
  FragA f;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int j = 0; j < 4; ++j) f.v[mt][j] = M[(g + 8 * mt) + CHI * (t + 4 * j)];
  return f;
}
__device__ __forceinline__ FragB load_fragB(const double* __restrict__ M, int g, int t) {
  FragB f;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < 2; ++i) f.v[nt][h][i] = M[(g + 8 * nt) + CHI * (2 * t + i + 8 * h)];
  return f;
}

// Phase 1, one column, IN PLACE:  buf[x', y'] = sum_{x,y} MX[x', x] MY[y', y] buf[x, y]   (legs X then Y)
template <int LAY, int X, int Y>
__device__ __forceinline__ void absorb_pair16(double* buf, uint32_t base, const FragA& mx, const FragB& my, int g, int t) {
  double2 b[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int h = 0; h < 2; ++h) b[j][h] = *reinterpret_cast<const double2*>(buf + (base ^ pos<LAY>(X, t + 4 * j) ^ pos<LAY>(Y, g + 8 * h)));
  __syncwarp();  // the column is overwritten below: every lane has read it first (mma.sync converges the warp anyway)
  // absorb X: D1[x' = g + 8 mt, y = 2t + i + 8h]
  double d1[2][2][2][2];  // [mt][h][s][i]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      d1[mt][h][0][0] = d1[mt][h][0][1] = d1[mt][h][1][0] = d1[mt][h][1][1] = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dmma(d1[mt][h][0][0], d1[mt][h][0][1], mx.v[mt][j], b[j][h].x);
        dmma(d1[mt][h][1][0], d1[mt][h][1][1], mx.v[mt][j], b[j][h].y);
      }
    }
  // absorb Y from registers: D2[x' = g + 8 mt, y' = 2t + i' + 8 nt]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      double p0 = 0, p1 = 0, q0 = 0, q1 = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          dmma(p0, p1, d1[mt][h][0][i], my.v[nt][h][i]);
          dmma(q0, q1, d1[mt][h][1][i], my.v[nt][h][i]);
        }
      const uint32_t a = base ^ pos<LAY>(X, g + 8 * mt);
      *reinterpret_cast<double2*>(buf + (a ^ pos<LAY>(Y, 2 * t + 8 * nt))) = make_double2(p0, q0);
      *reinterpret_cast<double2*>(buf + (a ^ pos<LAY>(Y, 2 * t + 1 + 8 * nt))) = make_double2(p1, q1);
    }
}

// Phase 2, one column: acc[v' tile mt][v tile h] += sum_{s, u'} A[u', v'] * ( sum_u MU[u', u] P[u, v] )
template <int LAY, int U, int V>
__device__ __forceinline__ void absorb_close16(const double* P, const double* A, uint32_t base, const FragA& mu, int g, int t,
                                               double (&acc)[2][2][2]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    double2 p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = *reinterpret_cast<const double2*>(P + (base ^ pos<LAY>(U, t + 4 * j) ^ pos<LAY>(V, g + 8 * h)));
    // T[v = g + 8h, u' = 2t + i + 8 nt]
    double tt[2][2][2];  // [nt][s][i]
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      tt[nt][0][0] = tt[nt][0][1] = tt[nt][1][0] = tt[nt][1][1] = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dmma(tt[nt][0][0], tt[nt][0][1], p[j].x, mu.v[nt][j]);
        dmma(tt[nt][1][0], tt[nt][1][1], p[j].y, mu.v[nt][j]);
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const double2 a = *reinterpret_cast<const double2*>(A + (base ^ pos<LAY>(U, 2 * t + i + 8 * nt) ^ pos<LAY>(V, g + 8 * mt)));
          dmma(acc[mt][h][0], acc[mt][h][1], a.x, tt[nt][0][i]);
          dmma(acc[mt][h][0], acc[mt][h][1], a.y, tt[nt][1][i]);
        }
  }
}

// shared memory: ring (3 x 64 KiB) | red 16 KiB (one output at a time) | raw 2 x 256 doubles | 16 mbarriers
constexpr size_t SMEM_DOUBLES = (size_t)3 * SLICE + NCW * MSG + 2 * MSG + 16;
constexpr size_t SMEM_BYTES = SMEM_DOUBLES * sizeof(double);
enum { MB_FULL = 0 /*3*/, MB_DONE = 3 /*3*/, MB_FULL2 = 6 /*2*/, MB_EMPTY2 = 8 /*2*/ };
enum { BAR_COMPUTE = 1 };

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(N) : "memory");
}

// canonical A_v[s, a0..a3] -> private image (run once per upload)
__global__ void swizzle_sites16(const ItemDesc* items, int n_items, const double* __restrict__ src, double* __restrict__ dst) {
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    if (items[item].branch != 0) continue;  // one pass per vertex
    const int64_t off = items[item].site_off;
    for (int c = threadIdx.x; c < NTENSOR / 2; c += blockDim.x) {
      const uint32_t p = global_pos(0, c & 15) ^ global_pos(1, (c >> 4) & 15) ^ global_pos(2, (c >> 8) & 15) ^ global_pos(3, c >> 12);
      *reinterpret_cast<double2*>(dst + off + p) = *reinterpret_cast<const double2*>(src + off + 2 * c);
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) bp_update_sliced_c16(Args k) {
  extern __shared__ __align__(128) double smem[];
  double* ring = smem;
  double* red = smem + 3 * SLICE;
  double* raw = red + NCW * MSG;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(raw + 2 * MSG);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int G = gridDim.x;
  if (sweep_already_converged(k.resmax, k.stop_key)) return;
  if ((int)blockIdx.x >= k.n_items) return;
  double* scratch = k.scratch + (size_t)blockIdx.x * NTENSOR;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) {
      mbar_init(&mbar[MB_FULL + i], 1);
      mbar_init(&mbar[MB_DONE + i], NCW);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&mbar[MB_FULL2 + i], 1);
      mbar_init(&mbar[MB_EMPTY2 + i], NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  // mbarrier phase parities follow from use counts: ring-1 slot b is used by slices b, b+3, ... (6 / 5 / 5 uses per
  // item), ring-2 stage st by steps st, st+2, ... (16 uses per item: parity (q >> 1) & 1)
#define RING1_PARITY(it, r) ((((r) % 3 == 0 ? 6u : 5u) * (uint32_t)(it) + (uint32_t)(r) / 3u) & 1u)

  if (warp == NCW) {
    // ================================ producer warp ================================
    peer_gate(k.peer, lane);  // multi-GPU: the peers' cut-edge messages of the previous sweep have landed
    int it = 0;
    for (int item = blockIdx.x; item < k.n_items; item += G, ++it) {
      const ItemDesc* d = k.items + item;
      const double* Aimg = k.sites + d->site_off;
      const int br = d->branch;
      // ---- phase 1: slices of leg 3 (branch P, contiguous) or leg 0 (branch Q, 16 runs of 4 KiB) ----
      auto load1 = [&](int r) {
        const int b = r % 3;
        double* dst = ring + b * SLICE;
        fence_proxy_async();  // generic-proxy accesses of the previous tenant happen-before the async-proxy writes
        if (lane == 0) mbar_expect_tx(&mbar[MB_FULL + b], SLICE * 8);
        __syncwarp();
        if (br == 0) {
          if (lane < 4) tma_bulk_g2s(dst + lane * 2048, Aimg + ((size_t)r << 13) + lane * 2048, 16384, &mbar[MB_FULL + b]);
        } else {
          if (lane < 16) tma_bulk_g2s(dst + lane * 512, Aimg + ((size_t)lane << 13) + ((size_t)r << 9), 4096, &mbar[MB_FULL + b]);
        }
      };
      load1(0);
      load1(1);
      for (int r = 0; r < 16; ++r) {
        const int b = r % 3;
        mbar_wait(&mbar[MB_DONE + b], RING1_PARITY(it, r));  // compute warps finished slice r in place
        const double* src = ring + b * SLICE;
        // every lane that stores owns its bulk group: lanes issue, commit and (below) wait symmetrically
        if (br == 0) {
          if (lane < 4) tma_bulk_s2g(scratch + ((size_t)r << 13) + lane * 2048, src + lane * 2048, 16384);
        } else {
          if (lane < 16) tma_bulk_s2g(scratch + ((size_t)lane << 13) + ((size_t)r << 9), src + lane * 512, 4096);
        }
        bulk_commit();
        if (r + 2 < 16) {
          bulk_wait_read<1>();  // the store of slice r-1 has drained its buffer, which slice r+2 re-uses
          __syncwarp();
          load1(r + 2);
        }
      }
      bulk_wait<0>();  // P / Q image complete in L2 before phase 2 reads it back
      __syncwarp();
      // ---- phase 2: 32 half slices of P (scratch) and A, ring of 2 x (32 KiB + 32 KiB) ----
      for (int q = 0; q < 32; ++q) {
        const int st = q & 1, r = q >> 1, hh = q & 1;
        if (q >= 2) mbar_wait(&mbar[MB_EMPTY2 + st], ((q - 2) >> 1) & 1);
        double* dp = ring + st * SLICE;
        double* da = dp + HALF;
        fence_proxy_async();
        if (lane == 0) mbar_expect_tx(&mbar[MB_FULL2 + st], SLICE * 8);
        __syncwarp();
        if (br == 0) {
          // a0-half slice (a0 = r, a1[3] = hh): 16 runs (a3) of 2 KiB
          const size_t go = ((size_t)r << 9) + ((size_t)hh << 8);
          if (lane < 16) tma_bulk_g2s(dp + lane * 256, scratch + ((size_t)lane << 13) + go, 2048, &mbar[MB_FULL2 + st]);
          else tma_bulk_g2s(da + (lane - 16) * 256, Aimg + ((size_t)(lane - 16) << 13) + go, 2048, &mbar[MB_FULL2 + st]);
        } else {
          // a3-half slice (a3 = r, a2[3] = hh): 32 runs (a0, a1[3]) of 1 KiB
          const size_t go = ((size_t)r << 13) + ((size_t)hh << 7) + ((size_t)lane << 8);
          tma_bulk_g2s(dp + lane * 128, scratch + go, 1024, &mbar[MB_FULL2 + st]);
          tma_bulk_g2s(da + lane * 128, Aimg + go, 1024, &mbar[MB_FULL2 + st]);
        }
      }
      // drain the two outstanding "empty" signals so that the counters stay aligned with the next item
      for (int st = 0; st < 2; ++st) mbar_wait(&mbar[MB_EMPTY2 + st], 1);  // 16th use of the item
    }
  } else {
  // ================================ compute warps ================================
  // (they read peer-written messages too -- the fragments below -- so they wait for the gate as well)
  if (k.peer.nranks > 1 && warp == 0) peer_gate(PeerArgs{k.peer.nranks, k.peer.rank, k.peer.my_mailbox, k.peer.peer_mailbox, k.peer.wait_id, k.peer.wait_mask, 0,
                                                        nullptr, nullptr, nullptr, nullptr, k.peer.error_flag}, lane);
  if (k.peer.nranks > 1) onchip::bar_sync(BAR_COMPUTE, NCT);
  int it = 0;
  for (int item = blockIdx.x; item < k.n_items; item += G, ++it) {
    const ItemDesc* d = k.items + item;
    const int br = d->branch;
    if (k.io.progress) {  // streamed upload: the item's messages (fragments, old values) have arrived -- ONE warp polls
      if (warp == 0) hostio_wait(k.io, d->need);
      onchip::bar_sync(BAR_COMPUTE, NCT);
    }
    const int lx = br == 0 ? 0 : 2, ly = br == 0 ? 1 : 3;  // pair absorbed in phase 1
    const int lu = br == 0 ? 2 : 0, lv = br == 0 ? 3 : 1;  // pair closed in phase 2
    {
      const FragA mx = load_fragA(k.msg_in + d->in_off[lx], g, t);
      const FragB my = load_fragB(k.msg_in + d->in_off[ly], g, t);
      for (int r = 0; r < 16; ++r) {
        const int b = r % 3;
        mbar_wait(&mbar[MB_FULL + b], RING1_PARITY(it, r));
        double* buf = ring + b * SLICE;
#pragma unroll 1
        for (int c = warp; c < 16; c += NCW) {
          if (br == 0)  // slice a3 = r, columns a2 = c, pair (0, 1)
            absorb_pair16<L_A3, 0, 1>(buf, pos<L_A3>(2, c) ^ pos<L_A3>(3, r), mx, my, g, t);
          else          // slice a0 = r, columns a1 = c, pair (2, 3)
            absorb_pair16<L_A0, 2, 3>(buf, pos<L_A0>(1, c) ^ pos<L_A0>(0, r), mx, my, g, t);
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) mbar_arrive(&mbar[MB_DONE + b]);
      }
    }
    double accA[2][2][2], accB[2][2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) accA[a][b][0] = accA[a][b][1] = accB[a][b][0] = accB[a][b][1] = 0.0;
    {
      const FragA mu = load_fragA(k.msg_in + d->in_off[lu], g, t);  // absorb U, close V  -> out[V]
      const FragA mv = load_fragA(k.msg_in + d->in_off[lv], g, t);  // absorb V, close U  -> out[U]
      for (int q = 0; q < 32; ++q) {
        const int st = q & 1, r = q >> 1, hh = q & 1;
        mbar_wait(&mbar[MB_FULL2 + st], (q >> 1) & 1);
        const double* Pb = ring + st * SLICE;
        const double* Ab = Pb + HALF;
        const int c = warp + 8 * hh;  // column index of the spectator leg
        if (br == 0) {  // half slice a0' = r, column a1' = c; pair (2, 3)
          const uint32_t base = pos<L_A0H>(1, c) ^ pos<L_A0H>(0, r);
          absorb_close16<L_A0H, 2, 3>(Pb, Ab, base, mu, g, t, accA);
          absorb_close16<L_A0H, 3, 2>(Pb, Ab, base, mv, g, t, accB);
        } else {        // half slice a3' = r, column a2' = c; pair (0, 1)
          const uint32_t base = pos<L_A3H>(2, c) ^ pos<L_A3H>(3, r);
          absorb_close16<L_A3H, 0, 1>(Pb, Ab, base, mu, g, t, accA);
          absorb_close16<L_A3H, 1, 0>(Pb, Ab, base, mv, g, t, accB);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&mbar[MB_EMPTY2 + st]);
      }
    }
    // ---- cross-warp reduction + fused epilogue, one output at a time (accA -> out[lv], accB -> out[lu]) ----
#pragma unroll 1
    for (int o = 0; o < 2; ++o) {
      double* mine = red + warp * MSG;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int el = (g + 8 * mt) + CHI * (2 * t + i + 8 * h);  // out[v', v] at v' + 16 v
            mine[el] = o == 0 ? accA[mt][h][i] : accB[mt][h][i];
          }
      onchip::bar_sync(BAR_COMPUTE, NCT);
      {
        const int el = threadIdx.x;  // 256 compute threads, 256 elements
        double s = 0;
#pragma unroll
        for (int w = 0; w < NCW; ++w) s += red[w * MSG + el];
        raw[o * MSG + el] = s;
      }
      onchip::bar_sync(BAR_COMPUTE, NCT);
    }
    if (warp < 2) {
      const int leg = warp == 0 ? lv : lu;
      const int64_t off = d->out_off[leg];
      double* peer_m = (k.peer.nranks > 1 && d->peer[leg] >= 0) ? k.peer.peer_out[d->peer[leg]] + off : nullptr;
      warp_epilogue<double>(raw + warp * MSG, k.msg_in + off, k.msg_out + off, MSG, k.normalize,
                            k.residual ? k.residual + d->out_edge[leg] : nullptr, lane, k.resmax, peer_m,
                            k.io.host_out ? k.io.host_out + off : nullptr);
    }
    onchip::bar_sync(BAR_COMPUTE, NCT);  // raw / red are re-used by the next item
  }
  }
  peer_post_when_last(k.peer, false);  // peer stores were released where they were issued (warp_epilogue)
  hostio_finish(k.io);
}

}  // namespace sliced
}  // namespace bpx
