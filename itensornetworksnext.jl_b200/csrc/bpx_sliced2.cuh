// BPX_KERNEL_SLICED, version 2: (degree 4, chi = 16, d = 2, Float64) -- BASELINE config 5 -- with the working set of the
// vertices in flight kept INSIDE the 126 MB L2.
//
// Version 1 (bpx_sliced.cuh) gives every CTA its own (vertex, branch) item: 148 x (1 MiB tensor + 1 MiB scratch) does not
// fit the L2, the tensor is streamed four times and the intermediate round-trips through DRAM: 6.85 MB of DRAM traffic
// per vertex against 1.07 MB algorithmic, 3.7 TB/s at full speed -- the memory system, not the tensor pipe, was the limit.
//
// Here a GROUP of G CTAs (8, or 4) shares ONE vertex, so that only 148/G vertices are in flight (G = 8: 19 groups x
// (2 tensors + 2 scratch images) = 76 MiB), and the three passes over the tensor are software-pipelined ACROSS vertices so
// that no CTA ever waits at a group barrier:
//
//   S1(v): a3-slices            A[.., a3 = r]  --absorb M0, M1 in place-->  P[.., a3 = r]              -> scratch[v & 1]
//   S2(v): a0-half-slices       P, A           --absorb 2 / close 3, absorb 3 / close 2-->  out3, out2 partial sums
//                               A (same slice) --absorb M2, M3 in place-->  Q[a0 = r, ..]  OVER P in scratch[v & 1]
//   S3(v): a3-half-slices       Q, A           --absorb 0 / close 1, absorb 1 / close 0-->  out1, out0 partial sums
//
// (branch Q's first pass rides on the tensor slices that branch P's second pass has in shared memory anyway: the tensor is
// read three times instead of four, and Q overwrites P slice by slice).  Every CTA runs the stage sequence
//   S1(0) S2(0) | S1(1) S3(0) S2(1) | S1(2) S3(1) S2(2) | ... | S3(n-1)
// on its share of the slices (member j of a group owns a3 / a0 values j, j + G, ...): between the end of a stage and the
// first load of the stage that depends on ALL members' stores there is always a whole stage of independent work, so the
// group barriers (monotone counters in global memory, release / acquire at gpu scope) cost nothing.  The compute warps
// never synchronise with each other or with anything global: every warp stores its partial 16x16 output tiles straight
// into an L2-resident buffer and bumps a shared-memory event counter; a COMMUNICATION warp turns those events (and the
// producer's "stage stored" events) into gpu-scope fences + arrivals on the group counters, and an EPILOGUE warp of the
// member that owns an output sums the 8 x G partial tiles in a fixed order and finishes the message (sum-normalisation,
// residual, stores, cut-edge peer stores) -- all off the tensor pipe's critical path.
//
// Warp roles (352 threads, 1 CTA / SM): 8 compute warps (DMMA), 1 producer warp (TMA loads / stores, L2 prefetch),
// 1 communication warp, 1 epilogue warp.  Same private tensor image, shared-memory layouts and DMMA micro-kernels as
// version 1.
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libbpx links cudart only)

#include "bpx_sliced.cuh"

namespace bpx {
namespace sliced2 {

using namespace sliced;

// CW compute warps + producer + communication + post-processing (reduce partial tiles, finish messages)
template <int CW>
constexpr int nthreads2() { return (CW + 3) * 32; }
constexpr int PART_GEN = 3;  // generations of partial tiles in flight (vertex v uses v % 3): the epilogue of v may lag two vertices
constexpr size_t PART_PER_GROUP = (size_t)PART_GEN * 8 * 4 * MSG;  // doubles: [v % 3][member][output leg][256], fragment order
constexpr size_t PART1_PER_CTA = (size_t)2 * 2 * 8 * MSG;          // doubles: [dump parity][output][contributor][256], fragment order

struct VItem {  // one (degree 4, chi 16) vertex
  int64_t site_off;   // elements, into the private image buffer
  int64_t in_off[4];
  int64_t out_off[4];
  int32_t out_edge[4];
  int32_t peer[4];    // rank owning the head of out-edge i if it lives elsewhere (cut edge), else -1
  int64_t need;       // streamed host I/O: prefix of the upload that holds every message this vertex reads
};

// per-group counters.  B1 / B2: monotone over the vertices (a member can only arrive for vertex v + 1 after it has passed
// the barrier of vertex v, so "count >= G (v + 1)" is exact).  B3: one counter per GENERATION of partial tiles (vertex
// v uses v % PART_GEN), both dumps of a vertex counted: a member's post-processing warp may lag its compute warps, so a
// single monotone counter could reach its target with one member counted twice and another missing; per generation the
// arrivals for vertex v + PART_GEN cannot start before every member has finished the epilogue of vertex v.
enum { GS_B1 = 0, GS_B2 = 1, GS_B3 = 2 /* .. 4 */, GS_STRIDE = 8 };

struct Args {
  const VItem* items;        // grouped: group g owns items[group_ptr[g] .. group_ptr[g + 1])
  const int32_t* group_ptr;  // [n_groups + 1]
  int n_groups;
  int G;                     // CTAs per group (8 or 4); the last group may have fewer (>= 4, a divisor of 16)
  const double* sites;       // private swizzled image
  double* scratch;           // per group: 2 x NTENSOR doubles
  double* partials;          // per group: PART_PER_GROUP doubles (one partial tile per member and output leg)
  double* part1;             // per CTA: PART1_PER_CTA doubles (per-warp partial tiles of the last two dumps, L2 resident)
  unsigned int* gsync;       // per group: GS_STRIDE counters, zeroed before the launch
  const double* msg_in;
  double* msg_out;
  double* residual;
  unsigned long long* resmax;  // this sweep's residual key (atomicMax)
  int normalize;
  PeerArgs peer;               // multi-GPU: gate / direct peer stores / post (nranks <= 1: unused)
  HostIO io;                   // streamed host I/O (bpx_sweep_host), all NULL otherwise
  unsigned long long stop_key; // device-side convergence test (sweep_already_converged), 0: none
  long long* timing;           // debug (BPX_SLICED_TIMING builds): per-CTA cycle counters, 16 per CTA
  int flags;                   // tuning experiments (BPX_SLICED_FLAGS): 1 = no L2 eviction hints, 2 = L2 prefetch of the next vertex's
                               // tensor slices (measured: +0.8 MB of DRAM traffic per vertex for no gain in time: off)
};

#ifdef BPX_SLICED_TIMING
#define TCLK() clock64()
#define TACC(slot, t0) do { if (lane == 0 && k.timing) k.timing[blockIdx.x * 16 + (slot)] += clock64() - (t0); } while (0)
#else
#define TCLK() 0ll
#define TACC(slot, t0) do { (void)(t0); } while (0)
#endif

// shared memory: ring 3 x 64 KiB | staged messages 2 x 4 x 2 KiB | raw 2 KiB | mbarriers + event counters
constexpr size_t SMEM2_DOUBLES = (size_t)3 * SLICE + 8 * MSG + MSG + 16;
constexpr size_t SMEM2_BYTES = SMEM2_DOUBLES * sizeof(double);
enum { M2_FULL = 0 /*3*/, M2_DONE = 3 /*3*/, M2_MSG = 6 /*2*/, M2_EV = 8 /* 8 plain 32-bit event counters */ };
enum { EV_B1 = 0, EV_B2 = 1, EV_DUMP = 2, EV_EPI = 3, EV_RED = 4 };  // S1 / S2 stages stored, warp dumps written, epilogues
                                                                      // finished, dumps reduced

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned int* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ bool mbar_try_once(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// one lane waits until a group counter has reached `target` (bounded: a group member that never arrives is a bug or a
// grid that is not co-resident -- trap instead of hanging the GPU)
__device__ __forceinline__ void group_wait(const unsigned int* ctr, unsigned int target) {
  if (ld_acquire_gpu(ctr) >= target) return;
  const long long t0 = clock64();
  unsigned ns = 32;
  while (ld_acquire_gpu(ctr) < target) {
    __nanosleep(ns);
    if (ns < 512) ns += ns;
    if (clock64() - t0 > 20000000000ll) __trap();
  }
}

// Strided half slices move as ONE TMA tensor copy each (cp.async.bulk.tensor, SASS UTMALDG / UTMASTG) instead of 16-32 plain
// bulk copies of 1-2 KiB: measured, a CTA's TMA unit retires one small bulk copy per ~120 clk, which made the a3-half
// slices (64 x 1 KiB per step) the bottleneck of the whole kernel.  Two views of a buffer of doubles:
//   view A0H: dims {256, 32, ceil(N / 8192)}, strides {256, 8192} doubles, box {256, 1, 16}:
//             the a0-half slice (a0 = r, a1[3] = hh) of a tensor at element offset o starts at s = o + (r << 9) + (hh << 8)
//             -> coordinates (0, (s >> 8) & 31, s >> 13)          [16 rows (a3) of 2 KiB]
//   view A3H: dims {128, 2, ceil(N / 256)}, strides {128, 256} doubles, box {128, 1, 32}:
//             the a3-half slice (a3 = r, a2[3] = hh) starts at s = o + (r << 13) + (hh << 7)
//             -> coordinates (0, (s >> 7) & 1, s >> 8)            [32 rows (a0, a1[3]) of 1 KiB]
// Both land densely in shared memory in row order: exactly the layouts L_A0H / L_A3H.
struct TensorMaps {
  CUtensorMap a0h_sites, a3h_sites, a0h_scratch, a3h_scratch;
};
__device__ __forceinline__ void tma_tensor3_g2s(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_tensor3_s2g(const CUtensorMap* tm, int c0, int c1, int c2, const void* src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];\n" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2),
               "r"(smem_u32(src))
               : "memory");
}

// L2 eviction priorities: the scratch images (P / Q) are written and re-read by the whole group within one vertex period and
// must not be written back to DRAM in between (evict_last); the tensor's last pass will not be needed again (evict_first)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_tensor3_g2s_hint(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;\n" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_tensor3_s2g_hint(const CUtensorMap* tm, int c0, int c1, int c2, const void* src, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], %5;\n" ::"l"(tm), "r"(c0),
               "r"(c1), "r"(c2), "r"(smem_u32(src)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_bulk_s2g_hint(void* gdst, const void* ssrc, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;\n" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes),
               "l"(pol)
               : "memory");
}

// In-place pair absorption split by output rows (16-warp variant: lower register footprint, and two warps can share one
// column): load the whole column once, then produce the rows x' = g + 8 mt of  buf[x', y'] = sum MX[x', x] MY[y', y] buf[x, y]
template <int LAY, int X, int Y>
__device__ __forceinline__ void load_col16(const double* buf, uint32_t base, int g, int t, double2 (&b)[4][2]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int h = 0; h < 2; ++h) b[j][h] = *reinterpret_cast<const double2*>(buf + (base ^ pos<LAY>(X, t + 4 * j) ^ pos<LAY>(Y, g + 8 * h)));
}
template <int LAY, int X, int Y>
__device__ __forceinline__ void absorb_pair16_rows(double* buf, uint32_t base, const double2 (&b)[4][2], const double (&mxr)[4], const FragB& my,
                                                   int mt, int g, int t) {
  double d1[2][2][2];  // [h][s][i]: D1[x' = g + 8 mt, y = 2t + i + 8h]
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    d1[h][0][0] = d1[h][0][1] = d1[h][1][0] = d1[h][1][1] = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dmma(d1[h][0][0], d1[h][0][1], mxr[j], b[j][h].x);
      dmma(d1[h][1][0], d1[h][1][1], mxr[j], b[j][h].y);
    }
  }
  const uint32_t a = base ^ pos<LAY>(X, g + 8 * mt);
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    double p0 = 0, p1 = 0, q0 = 0, q1 = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        dmma(p0, p1, d1[h][0][i], my.v[nt][h][i]);
        dmma(q0, q1, d1[h][1][i], my.v[nt][h][i]);
      }
    *reinterpret_cast<double2*>(buf + (a ^ pos<LAY>(Y, 2 * t + 8 * nt))) = make_double2(p0, q0);
    *reinterpret_cast<double2*>(buf + (a ^ pos<LAY>(Y, 2 * t + 1 + 8 * nt))) = make_double2(p1, q1);
  }
}

enum { K_S1 = 0, K_S2 = 1, K_S3 = 2 };
struct Step {
  int kind, v, q;
};
// step t of a member's sequence  S1(0) S2(0) | S1(1) S3(0) S2(1) | ... | S3(n-1)
__device__ __forceinline__ Step decode_step(int t, int n, int nS1, int nS2) {
  Step s;
  if (t < nS1) {
    s.kind = K_S1; s.v = 0; s.q = t;
    return s;
  }
  const int per = nS1 + 2 * nS2;
  const int u = t + nS2, v = u / per, w = u - v * per;
  if (v == n) { s.kind = K_S3; s.v = n - 1; s.q = w; }
  else if (w < nS1) { s.kind = K_S1; s.v = v; s.q = w; }
  else if (w < nS1 + nS2) { s.kind = K_S3; s.v = v - 1; s.q = w - nS1; }
  else { s.kind = K_S2; s.v = v; s.q = w - nS1 - nS2; }
  return s;
}

// canonical A_v[s, a0..a3] -> private image (run once per upload)
__global__ void swizzle_sites16v(const VItem* items, int n_items, const double* __restrict__ src, double* __restrict__ dst) {
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int64_t off = items[item].site_off;
    for (int c = threadIdx.x; c < NTENSOR / 2; c += blockDim.x) {
      const uint32_t p = global_pos(0, c & 15) ^ global_pos(1, (c >> 4) & 15) ^ global_pos(2, (c >> 8) & 15) ^ global_pos(3, c >> 12);
      *reinterpret_cast<double2*>(dst + off + p) = *reinterpret_cast<const double2*>(src + off + 2 * c);
    }
  }
}

template <int CW>  // compute warps: 8 (a column per warp: both products of a step) or 16 (two warps per column: one product each)
__global__ void __launch_bounds__(nthreads2<CW>(), 1) bp_update_sliced_c16g(Args k, const __grid_constant__ TensorMaps tm) {
  constexpr int WARP_PRODUCER = CW, WARP_COMM = CW + 1, WARP_POST = CW + 2;
  extern __shared__ __align__(128) double smem[];
  double* ring = smem;
  double* msgs = smem + 3 * SLICE;   // [2][4][256]
  double* raw = msgs + 8 * MSG;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(raw + MSG);
  volatile unsigned int* ev = reinterpret_cast<volatile unsigned int*>(&mbar[M2_EV]);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
  if (sweep_already_converged(k.resmax, k.stop_key)) return;

  // ---- who am I: group, member, members of my group ----
  const int grp = blockIdx.x / k.G, j = blockIdx.x - grp * k.G;
  const int Gm = min(k.G, (int)gridDim.x - grp * k.G);
  const bool active = grp < k.n_groups;
  const int base_item = active ? k.group_ptr[grp] : 0;
  const int n = active ? k.group_ptr[grp + 1] - base_item : 0;
  const int nS1 = 16 / Gm, nS2 = 32 / Gm;  // steps per stage (S3 like S2)
  const int T = n * (nS1 + 2 * nS2);
  double* const scratch0 = k.scratch + (size_t)grp * 2 * NTENSOR;
  double* const part0 = k.partials + (size_t)grp * PART_PER_GROUP;
  double* const part1 = k.part1 + (size_t)blockIdx.x * PART1_PER_CTA;
  unsigned int* const gs = k.gsync + (size_t)grp * GS_STRIDE;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) {
      mbar_init(&mbar[M2_FULL + i], 1);
      mbar_init(&mbar[M2_DONE + i], CW);
    }
    mbar_init(&mbar[M2_MSG + 0], 1);
    mbar_init(&mbar[M2_MSG + 1], 1);
    for (int i = 0; i < 8; ++i) ev[i] = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (n > 0) {
  if (warp == WARP_PRODUCER) {
    // ================================ producer warp ================================
    const long long tp0 = TCLK();
    peer_gate(k.peer, lane);  // multi-GPU: the peers' cut-edge messages of the previous sweep have landed
    const bool hints = !(k.flags & 1);
    const uint64_t pol_keep = l2_policy_evict_last(), pol_done = l2_policy_evict_first();
    int L = 0;                 // next step to load
    unsigned int myB1 = 0, myB2 = 0;
    int pending_post = -1;     // stage kind whose stores were committed with the previous step; post once they completed
    auto post = [&](int kind) {
      // all lanes have waited for their own bulk groups: the stage's stores are complete.  Hand the arrival on the group
      // counter to the communication warp (a gpu-scope fence under load costs microseconds: not on this warp's clock)
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        if (kind == K_S1) ev[EV_B1] = myB1 + 1u;
        else ev[EV_B2] = myB2 + 1u;
      }
      if (kind == K_S1) ++myB1;
      else ++myB2;
      __syncwarp();
    };
    // may the load of step `s` go out?  The first step of S2(v) / S3(v) needs EVERY member's stores of S1(v) / S2(v):
    // my own must have been posted (else waiting would dead-lock), and the group counter must have reached its target --
    // checked without blocking, so that this warp keeps retiring steps (stores, posts) of the stage in flight meanwhile
    auto can_load = [&](const Step& s) {
      if (s.q != 0 || s.kind == K_S1) return true;
      if ((s.kind == K_S2 ? myB1 : myB2) < (unsigned)s.v + 1u) return false;
      unsigned int c = 0;
      if (lane == 0) c = ld_acquire_gpu(gs + (s.kind == K_S2 ? GS_B1 : GS_B2));
      c = __shfl_sync(0xffffffffu, c, 0);
      return c >= (unsigned)Gm * (unsigned)(s.v + 1);
    };
    auto issue_load = [&](int t, const Step& s) {
      const VItem* d = k.items + base_item + s.v;
      const int b = t % 3;
      double* dst = ring + b * SLICE;
      const double* Aimg = k.sites + d->site_off;
      const double* scr = scratch0 + (size_t)(s.v & 1) * NTENSOR;
      if (s.q == 0) {
        if (s.kind == K_S1) {
          // the vertex's four incoming messages -> staged copy (slot v & 1)
          if (k.io.progress) {
            hostio_wait(k.io, d->need);
            asm volatile("fence.proxy.async;\n" ::: "memory");
          }
          uint64_t* mb = &mbar[M2_MSG + (s.v & 1)];
          if (lane == 0) mbar_expect_tx(mb, 4 * MSG * 8);
          __syncwarp();
          if (lane < 4) tma_bulk_g2s(msgs + ((s.v & 1) * 4 + lane) * MSG, k.msg_in + d->in_off[lane], MSG * 8, mb);
        } else {
          // (can_load has seen every member's stores of the producing stage complete, with acquire semantics)
          if (s.kind == K_S2 && s.v + 1 < n && (k.flags & 2)) {
            // this member's tensor slices of the NEXT vertex: start them on their way from DRAM into the L2 now
            const double* An = k.sites + k.items[base_item + s.v + 1].site_off;
            for (int q = lane >> 2; q < nS1; q += 8)
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(An + ((size_t)(j + q * Gm) << 13) + (lane & 3) * 2048), "r"(16384) : "memory");
          }
        }
      }
      fence_proxy_async();  // generic-proxy accesses of the slot's previous tenant happen-before the async-proxy writes
      if (lane == 0) mbar_expect_tx(&mbar[M2_FULL + b], SLICE * 8);
      __syncwarp();
      if (s.kind == K_S1) {
        const int r = j + s.q * Gm;  // a3-slice: one contiguous 64 KiB run
        if (lane < 4) tma_bulk_g2s(dst + lane * 2048, Aimg + ((size_t)r << 13) + lane * 2048, 16384, &mbar[M2_FULL + b]);
      } else if (s.kind == K_S2) {
        const int r = j + (s.q >> 1) * Gm, hh = s.q & 1;  // a0-half slice (a0 = r, a1[3] = hh): 16 rows (a3) of 2 KiB
        const size_t go = ((size_t)r << 9) + ((size_t)hh << 8);
        const size_t sp = (size_t)(scr - k.scratch) + go, sa = (size_t)d->site_off + go;
        if (lane == 0) {
          if (hints) tma_tensor3_g2s_hint(dst, &tm.a0h_scratch, 0, (int)((sp >> 8) & 31), (int)(sp >> 13), &mbar[M2_FULL + b], pol_keep);
          else tma_tensor3_g2s(dst, &tm.a0h_scratch, 0, (int)((sp >> 8) & 31), (int)(sp >> 13), &mbar[M2_FULL + b]);
        }
        if (lane == 1) tma_tensor3_g2s(dst + HALF, &tm.a0h_sites, 0, (int)((sa >> 8) & 31), (int)(sa >> 13), &mbar[M2_FULL + b]);
      } else {
        const int r = j + (s.q >> 1) * Gm, hh = s.q & 1;  // a3-half slice (a3 = r, a2[3] = hh): 32 rows (a0, a1[3]) of 1 KiB
        const size_t go = ((size_t)r << 13) + ((size_t)hh << 7);
        const size_t sp = (size_t)(scr - k.scratch) + go, sa = (size_t)d->site_off + go;
        if (lane == 0) {
          if (hints) tma_tensor3_g2s_hint(dst, &tm.a3h_scratch, 0, (int)((sp >> 7) & 1), (int)(sp >> 8), &mbar[M2_FULL + b], pol_keep);
          else tma_tensor3_g2s(dst, &tm.a3h_scratch, 0, (int)((sp >> 7) & 1), (int)(sp >> 8), &mbar[M2_FULL + b]);
        }
        if (lane == 1) {
          if (hints) tma_tensor3_g2s_hint(dst + HALF, &tm.a3h_sites, 0, (int)((sa >> 7) & 1), (int)(sa >> 8), &mbar[M2_FULL + b], pol_done);
          else tma_tensor3_g2s(dst + HALF, &tm.a3h_sites, 0, (int)((sa >> 7) & 1), (int)(sa >> 8), &mbar[M2_FULL + b]);
        }
      }
    };
    auto pump = [&](int r) {  // issue every load that may go out now: at most two steps ahead of the retired one
      while (L < T && L <= r + 2) {
        const Step s = decode_step(L, n, nS1, nS2);
        if (!can_load(s)) break;
        issue_load(L, s);
        ++L;
      }
    };
    pump(-1);
    for (int r = 0; r < T; ++r) {
      const Step s = decode_step(r, n, nS1, nS2);
      const int b = r % 3;
      const long long td = TCLK();
      {  // compute warps finished step r?  Meanwhile keep trying a load that waits for the group
        const long long t0w = clock64();
        while (!__any_sync(0xffffffffu, mbar_try_once(&mbar[M2_DONE + b], (uint32_t)(r / 3) & 1u))) {
          if (L < T && L <= r + 1) pump(r - 1);
          if (clock64() - t0w > 40000000000ll) __trap();
        }
        mbar_wait(&mbar[M2_DONE + b], (uint32_t)(r / 3) & 1u);  // (every lane observes the completed phase itself)
      }
      TACC(8, td);
      const double* src = ring + b * SLICE;
      double* scr = scratch0 + (size_t)(s.v & 1) * NTENSOR;
      // every lane that stores owns its bulk group: lanes issue, commit and wait symmetrically (one group per step)
      if (s.kind == K_S1) {
        const int rr = j + s.q * Gm;
        if (lane < 4) {
          if (hints) tma_bulk_s2g_hint(scr + ((size_t)rr << 13) + lane * 2048, src + lane * 2048, 16384, pol_keep);
          else tma_bulk_s2g(scr + ((size_t)rr << 13) + lane * 2048, src + lane * 2048, 16384);
        }
      } else if (s.kind == K_S2) {
        const int rr = j + (s.q >> 1) * Gm, hh = s.q & 1;  // the A half of the slot now holds Q[a0 = rr, a1[3] = hh, ..]
        const size_t sp = (size_t)(scr - k.scratch) + ((size_t)rr << 9) + ((size_t)hh << 8);
        if (lane == 0) {
          if (hints) tma_tensor3_s2g_hint(&tm.a0h_scratch, 0, (int)((sp >> 8) & 31), (int)(sp >> 13), src + HALF, pol_keep);
          else tma_tensor3_s2g(&tm.a0h_scratch, 0, (int)((sp >> 8) & 31), (int)(sp >> 13), src + HALF);
        }
      }
      bulk_commit();
      const bool stage_end = (s.kind == K_S1 && s.q == nS1 - 1) || (s.kind == K_S2 && s.q == nS2 - 1);
      // the stage that needs this stage's stores follows immediately (first vertex, last vertex): nothing to overlap with
      bool adjacent = false;
      if (stage_end) {
        if (r + 1 >= T) adjacent = true;
        else {
          const Step nx = decode_step(r + 1, n, nS1, nS2);
          adjacent = (s.kind == K_S1 && nx.kind == K_S2 && nx.v == s.v) || (s.kind == K_S2 && nx.kind == K_S3 && nx.v == s.v);
        }
      }
      (void)adjacent;
      (void)pending_post;
      const long long tb = TCLK();
      bulk_wait_read<1>();  // the store of step r-1 has drained its slot, which step r+2 re-uses
      __syncwarp();
      pump(r);
      if (stage_end) {
        // post the stage as early as possible: the compute warps have just started step r+1, this warp has nothing else to
        // do until they finish it -- wait for the stage's last store to complete here, then hand the arrival on
        bulk_wait<0>();
        post(s.kind);
        pump(r);  // (first vertex / last vertex: the dependent stage follows immediately)
      }
      TACC(9, tb);
    }
    bulk_wait<0>();
    __syncwarp();
    TACC(10, tp0);
  } else if (warp == WARP_COMM) {
    // ================================ communication warp ================================
    // shared-memory events -> arrivals on the group's global counters.  One lane, a non-blocking service loop: every
    // pending event is forwarded as soon as it is seen (the fences this takes run here, next to nothing else).
    if (lane == 0) {
      unsigned sent_b1 = 0, sent_b2 = 0;
      const unsigned un = (unsigned)n;
      long long t0 = clock64();
      while (sent_b1 < un || sent_b2 < un) {
        bool progress = false;
        // B1(v): before anybody may overwrite the partial tiles of vertex v - 3 (S2(v) dumps into the same generation), my
        // epilogue warp must have finished with them
        if (sent_b1 < ev[EV_B1] && (sent_b1 < (unsigned)PART_GEN || ev[EV_EPI] + (unsigned)PART_GEN > sent_b1)) {
          red_release_gpu(gs + GS_B1);  // (release at gpu scope: the fence is part of it)
          ++sent_b1;
          progress = true;
        }
        if (sent_b2 < ev[EV_B2]) {
          red_release_gpu(gs + GS_B2);
          ++sent_b2;
          progress = true;
        }
        if (!progress) {
          __nanosleep(100);
          if (clock64() - t0 > 40000000000ll) __trap();  // ~20 s without any event: a member of the group is gone
        } else {
          t0 = clock64();
        }
      }
    }
  } else if (warp == WARP_POST) {
    // ================================ post-processing warp ================================
    // (a) dump d (even: S2(d / 2) -> out3, out2; odd: S3(d / 2) -> out1, out0): sum the eight contributors' partial tiles
    //     (written to this CTA's L2-resident dump slots a moment ago) in a fixed order -> this member's tile of the
    //     group's partial buffer, arrive on the group counter;
    // (b) vertex i complete in the whole group: sum the members' tiles of the out-edges this member owns and finish the
    //     messages (sum-normalisation, residual, stores, cut-edge peer stores).
    // A non-blocking service loop: a late group never holds up the reduction of this CTA's own dumps.
    const long long te0 = TCLK();
    int dmp = 0, epi = 0;
    long long t0 = clock64();
    while (dmp < 2 * n || epi < n) {
      bool progress = false;
      unsigned have = 0;
      if (lane == 0) have = ev[EV_DUMP];
      have = __shfl_sync(0xffffffffu, have, 0);
      if (dmp < 2 * n && have >= (unsigned)CW * (unsigned)(dmp + 1)) {
        __threadfence_block();
        const int i = dmp >> 1;
        const int legs[2] = {(dmp & 1) ? 1 : 3, (dmp & 1) ? 0 : 2};
        const double2* src = reinterpret_cast<const double2*>(part1 + (size_t)(dmp & 1) * 2 * 8 * MSG) + lane;
        double* dstg = part0 + ((size_t)(i % PART_GEN) * 8 + j) * 4 * MSG;
#pragma unroll 1
        for (int tile = 0; tile < 2; ++tile) {
          double2 acc[4];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            double2 v[4][4];
#pragma unroll
            for (int w = 0; w < 4; ++w)
#pragma unroll
              for (int q = 0; q < 4; ++q) v[w][q] = __ldcg(src + ((size_t)tile * 8 + half * 4 + w) * (MSG / 2) + q * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              double2 a = half == 0 ? v[0][q] : make_double2(acc[q].x + v[0][q].x, acc[q].y + v[0][q].y);
#pragma unroll
              for (int w = 1; w < 4; ++w) {
                a.x += v[w][q].x;
                a.y += v[w][q].y;
              }
              acc[q] = a;
            }
          }
          double2* out = reinterpret_cast<double2*>(dstg + (size_t)legs[tile] * MSG) + lane;
#pragma unroll
          for (int q = 0; q < 4; ++q) __stcg(out + q * 32, acc[q]);
        }
        __threadfence();  // every lane's stores, before lane 0's arrival
        __syncwarp();
        ++dmp;
        if (lane == 0) {
          red_release_gpu(gs + GS_B3 + ((dmp - 1) >> 1) % PART_GEN);
          ev[EV_RED] = (unsigned)dmp;
        }
        progress = true;
      }
      if (epi < n && dmp >= min(2 * n, 2 * epi + 2)) {
        unsigned ok = 0;
        if (lane == 0)
          ok = ld_acquire_gpu(gs + GS_B3 + epi % PART_GEN) >= 2u * (unsigned)Gm * (unsigned)(epi / PART_GEN + 1) ? 1u : 0u;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (ok) {
          const int i = epi;
          const VItem* d = k.items + base_item + i;
          if (k.io.progress) hostio_wait(k.io, d->need);
          const double* part = part0 + (size_t)(i % PART_GEN) * 8 * 4 * MSG;
#pragma unroll 1
          for (int leg = 0; leg < 4; ++leg) {
            if ((i + leg) % Gm != j) continue;  // this member finishes out-edge `leg` of vertex i
            // sum the Gm members' tiles (fragment order: double2 (i2 = 0, 1) at [(mt * 2 + h) * 32 + lane]) in member order
            double2 acc[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = make_double2(0.0, 0.0);
            for (int m = 0; m < Gm; ++m) {
              const double2* src = reinterpret_cast<const double2*>(part + ((size_t)m * 4 + leg) * MSG) + lane;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const double2 v = __ldcg(src + q * 32);
                acc[q].x += v.x;
                acc[q].y += v.y;
              }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int mt = q >> 1, h = q & 1;
              const int el = (g + 8 * mt) + CHI * (2 * t4 + 8 * h);  // out[v', v] at v' + 16 v
              raw[el] = acc[q].x;
              raw[el + CHI] = acc[q].y;
            }
            __syncwarp();
            const int64_t off = d->out_off[leg];
            double* peer_m = (k.peer.nranks > 1 && d->peer[leg] >= 0) ? k.peer.peer_out[d->peer[leg]] + off : nullptr;
            warp_epilogue<double>(raw, k.msg_in + off, k.msg_out + off, MSG, k.normalize,
                                  k.residual ? k.residual + d->out_edge[leg] : nullptr, lane, k.resmax, peer_m,
                                  k.io.host_out ? k.io.host_out + off : nullptr);
            __syncwarp();
          }
          ++epi;
          if (lane == 0) ev[EV_EPI] = (unsigned)epi;
          progress = true;
        }
      }
      if (!progress) {
        __nanosleep(100);
        if (clock64() - t0 > 40000000000ll) __trap();
      } else {
        t0 = clock64();
      }
    }
    TACC(13, te0);
  } else {
    // ================================ compute warps ================================
    int t = 0;  // running step index (slot = t % 3, mbarrier parity = (t / 3) & 1)
    const long long tc0 = TCLK();
    int n_dump = 0;  // dumps so far (S2(0), S3(0), S2(1), S3(1), ...)
    // one partial 16x16 output tile of this warp -> slot [output][contributor] of the CTA's two-deep dump buffer (L2),
    // fragment order, coalesced
    auto dump_tile = [&](int out, int contributor, const double (&acc)[2][2][2]) {
      double2* pa = reinterpret_cast<double2*>(part1 + ((((size_t)(n_dump & 1) * 2 + out) * 8 + contributor)) * MSG) + lane;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) __stcg(pa + (mt * 2 + h) * 32, make_double2(acc[mt][h][0], acc[mt][h][1]));
    };
    auto dump_begin = [&]() {
      if (n_dump >= 2) {  // the post-processing warp has finished with the dump before last (in practice: long ago)
        if (lane == 0)
          while (ev[EV_RED] + 2u <= (unsigned)n_dump) __nanosleep(64);
        __syncwarp();
      }
    };
    auto dump_end = [&]() {
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        atomicAdd_block(const_cast<unsigned int*>(&ev[EV_DUMP]), 1u);
      }
      ++n_dump;
    };
    for (int v = 0; v <= n; ++v) {
      if (v < n) {
        // ---- S1(v): absorb the pair (0, 1) in place, a3-slices ----
        const long long tm = TCLK();
        mbar_wait(&mbar[M2_MSG + (v & 1)], (uint32_t)(v >> 1) & 1u);
        if (warp == 0) TACC(5, tm);
        const double* mm = msgs + (v & 1) * 4 * MSG;
        const FragA mx = load_fragA(mm + 0 * MSG, g, t4);
        const FragB my = load_fragB(mm + 1 * MSG, g, t4);
        for (int q = 0; q < nS1; ++q, ++t) {
          const int b = t % 3, r = j + q * Gm;
          const long long tf1 = TCLK();
          mbar_wait(&mbar[M2_FULL + b], (uint32_t)(t / 3) & 1u);
          if (warp == 0) TACC(1, tf1);
          double* buf = ring + b * SLICE;
          if constexpr (CW == 8) {
#pragma unroll 1
            for (int c = warp; c < 16; c += 8) absorb_pair16<L_A3, 0, 1>(buf, pos<L_A3>(2, c) ^ pos<L_A3>(3, r), mx, my, g, t4);
          } else {  // one column per warp, rows in two passes
            const uint32_t base = pos<L_A3>(2, warp) ^ pos<L_A3>(3, r);
            double2 bc[4][2];
            load_col16<L_A3, 0, 1>(buf, base, g, t4, bc);
            __syncwarp();  // the column is overwritten below: every lane has read it first
            absorb_pair16_rows<L_A3, 0, 1>(buf, base, bc, mx.v[0], my, 0, g, t4);
            absorb_pair16_rows<L_A3, 0, 1>(buf, base, bc, mx.v[1], my, 1, g, t4);
          }
          fence_proxy_async();  // generic-proxy writes -> visible to the TMA store
          __syncwarp();
          if (lane == 0) mbar_arrive(&mbar[M2_DONE + b]);
        }
      }
      if (v >= 1) {
        // ---- S3(v-1): a3-half slices of Q and A; absorb 0 / close 1 -> out1, absorb 1 / close 0 -> out0 ----
        const int i = v - 1;
        const double* mm = msgs + (i & 1) * 4 * MSG;
        if constexpr (CW == 8) {
          const FragA mu = load_fragA(mm + 0 * MSG, g, t4);
          const FragA mv = load_fragA(mm + 1 * MSG, g, t4);
          double accA[2][2][2], accB[2][2][2];
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b2 = 0; b2 < 2; ++b2) accA[a][b2][0] = accA[a][b2][1] = accB[a][b2][0] = accB[a][b2][1] = 0.0;
          for (int q = 0; q < nS2; ++q, ++t) {
            const int b = t % 3, r = j + (q >> 1) * Gm, hh = q & 1;
            const long long tf3 = TCLK();
            mbar_wait(&mbar[M2_FULL + b], (uint32_t)(t / 3) & 1u);
            if (warp == 0) TACC(3, tf3);
            if (warp == 0 && q == 0) TACC(12, tf3);
            const double* Pb = ring + b * SLICE;
            const double* Ab = Pb + HALF;
            const int c = warp + 8 * hh;  // column a2'
            const uint32_t base = pos<L_A3H>(2, c) ^ pos<L_A3H>(3, r);
            absorb_close16<L_A3H, 0, 1>(Pb, Ab, base, mu, g, t4, accA);
            absorb_close16<L_A3H, 1, 0>(Pb, Ab, base, mv, g, t4, accB);
            __syncwarp();
            if (lane == 0) mbar_arrive(&mbar[M2_DONE + b]);
          }
          const long long tdump = TCLK();
          dump_begin();
          dump_tile(0, warp, accA);  // -> out1
          dump_tile(1, warp, accB);  // -> out0
          dump_end();
          if (warp == 0) TACC(4, tdump);
        } else {
          const int c8 = warp >> 1, role = warp & 1;  // role 0: absorb 0 / close 1 -> out1; role 1: absorb 1 / close 0 -> out0
          const FragA mu = load_fragA(mm + role * MSG, g, t4);
          double acc[2][2][2];
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b2 = 0; b2 < 2; ++b2) acc[a][b2][0] = acc[a][b2][1] = 0.0;
          for (int q = 0; q < nS2; ++q, ++t) {
            const int b = t % 3, r = j + (q >> 1) * Gm, hh = q & 1;
            const long long tf3 = TCLK();
            mbar_wait(&mbar[M2_FULL + b], (uint32_t)(t / 3) & 1u);
            if (warp == 0) TACC(3, tf3);
            if (warp == 0 && q == 0) TACC(12, tf3);
            const double* Pb = ring + b * SLICE;
            const double* Ab = Pb + HALF;
            const uint32_t base = pos<L_A3H>(2, c8 + 8 * hh) ^ pos<L_A3H>(3, r);
            if (role == 0) absorb_close16<L_A3H, 0, 1>(Pb, Ab, base, mu, g, t4, acc);
            else absorb_close16<L_A3H, 1, 0>(Pb, Ab, base, mu, g, t4, acc);
            __syncwarp();
            if (lane == 0) mbar_arrive(&mbar[M2_DONE + b]);
          }
          const long long tdump = TCLK();
          dump_begin();
          dump_tile(role, c8, acc);
          dump_end();
          if (warp == 0) TACC(4, tdump);
        }
      }
      if (v < n) {
        // ---- S2(v): a0-half slices of P and A; absorb 2 / close 3 -> out3, absorb 3 / close 2 -> out2;
        //      then the SAME tensor columns absorb (2, 3) in place: Q's first pass ----
        const double* mm = msgs + (v & 1) * 4 * MSG;
        if constexpr (CW == 8) {
          const FragA mu = load_fragA(mm + 2 * MSG, g, t4);
          const FragA mv = load_fragA(mm + 3 * MSG, g, t4);
          const FragB my3 = load_fragB(mm + 3 * MSG, g, t4);
          double accA[2][2][2], accB[2][2][2];
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b2 = 0; b2 < 2; ++b2) accA[a][b2][0] = accA[a][b2][1] = accB[a][b2][0] = accB[a][b2][1] = 0.0;
          for (int q = 0; q < nS2; ++q, ++t) {
            const int b = t % 3, r = j + (q >> 1) * Gm, hh = q & 1;
            const long long tf2 = TCLK();
            mbar_wait(&mbar[M2_FULL + b], (uint32_t)(t / 3) & 1u);
            if (warp == 0) TACC(2, tf2);
            double* Pb = ring + b * SLICE;
            double* Ab = Pb + HALF;
            const int c = warp + 8 * hh;  // column a1'
            const uint32_t base = pos<L_A0H>(1, c) ^ pos<L_A0H>(0, r);
            absorb_close16<L_A0H, 2, 3>(Pb, Ab, base, mu, g, t4, accA);
            absorb_close16<L_A0H, 3, 2>(Pb, Ab, base, mv, g, t4, accB);
            __syncwarp();  // every lane has read the column before it is overwritten (only this warp touches column c)
            absorb_pair16<L_A0H, 2, 3>(Ab, base, mu, my3, g, t4);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&mbar[M2_DONE + b]);
          }
          const long long tdump = TCLK();
          dump_begin();
          dump_tile(0, warp, accA);  // -> out3
          dump_tile(1, warp, accB);  // -> out2
          dump_end();
          if (warp == 0) TACC(4, tdump);
        } else {
          const int c8 = warp >> 1, role = warp & 1;  // role 0: absorb 2 / close 3 -> out3; role 1: absorb 3 / close 2 -> out2
          const FragA mu = load_fragA(mm + (2 + role) * MSG, g, t4);
          double acc[2][2][2];
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b2 = 0; b2 < 2; ++b2) acc[a][b2][0] = acc[a][b2][1] = 0.0;
          for (int q = 0; q < nS2; ++q, ++t) {
            const int b = t % 3, r = j + (q >> 1) * Gm, hh = q & 1;
            const long long tf2 = TCLK();
            mbar_wait(&mbar[M2_FULL + b], (uint32_t)(t / 3) & 1u);
            if (warp == 0) TACC(2, tf2);
            double* Pb = ring + b * SLICE;
            double* Ab = Pb + HALF;
            const uint32_t base = pos<L_A0H>(1, c8 + 8 * hh) ^ pos<L_A0H>(0, r);
            if (role == 0) absorb_close16<L_A0H, 2, 3>(Pb, Ab, base, mu, g, t4, acc);
            else absorb_close16<L_A0H, 3, 2>(Pb, Ab, base, mu, g, t4, acc);
            // Q's first pass on the same tensor column, rows split between the two warps of the column: both read the whole
            // column first (after their products), meet at the pair's named barrier, then write their own rows
            double2 bc[4][2];
            load_col16<L_A0H, 2, 3>(Ab, base, g, t4, bc);
            double mxr[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) mxr[jj] = mm[2 * MSG + (g + 8 * role) + CHI * (t4 + 4 * jj)];  // M2[g + 8 mt, t + 4 j], mt = role
            const FragB my3 = load_fragB(mm + 3 * MSG, g, t4);
            onchip::bar_sync(1 + c8, 64);
            absorb_pair16_rows<L_A0H, 2, 3>(Ab, base, bc, mxr, my3, role, g, t4);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&mbar[M2_DONE + b]);
          }
          const long long tdump = TCLK();
          dump_begin();
          dump_tile(role, c8, acc);
          dump_end();
          if (warp == 0) TACC(4, tdump);
        }
      }
    }
    if (warp == 0) TACC(0, tc0);
  }
  }
  __syncthreads();
  peer_post_when_last(k.peer, false);  // peer stores were released where they were issued (warp_epilogue)
  hostio_finish(k.io);
}

}  // namespace sliced2
}  // namespace bpx
