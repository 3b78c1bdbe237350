// BPX_KERNEL_ONCHIP: vertex-centric update kernel for buckets whose site tensor fits in shared memory
// (degree 4, chi = 8, d = 2, Float64: 8192 doubles = 64 KiB) -- BASELINE config 2's dominant bucket.
//
// One persistent CTA per SM loops over the bucket's vertices.  For a vertex u with link legs 0..3 it
// produces all four outgoing messages from ONE read of A_u with the leave-one-out tree
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
// (g <-> one leg of a pair, t <-> two bits of the other leg of the same pair, pairs (0,1) and (2,3))
// touches 8 distinct 16-byte bank groups per quarter warp: conflict free.  cp.async (LDGSTS, 16 B)
// scatters the canonical HBM tensor into that layout while the previous vertex is being computed.
#pragma once
#include "bpx_common.cuh"

namespace bpx {
namespace onchip {

constexpr int CHI = 8;
constexpr int NELEM = 2 * CHI * CHI * CHI * CHI;  // 8192 doubles
constexpr int NWARPS = 16;
constexpr int NTHREADS = NWARPS * 32;
constexpr int MSG = CHI * CHI;

// ---- swizzled position (in doubles, s = 0) of element (a0,a1,a2,a3); XOR-linear in every index bit ----
// bit 0: s | bits 1-3: bank group | bits 4-12: a0[2], a1[1], a1[2], a2[0..2], a3[0..2]
__device__ __forceinline__ uint32_t leg_pos(int leg, uint32_t a) {
  const uint32_t x02 = (a ^ (a >> 2)) & 1u, b1 = (a >> 1) & 1u, b2 = (a >> 2) & 1u;
  switch (leg) {
    case 0: return (x02 << 1) | (b1 << 2) | (b2 << 4);
    case 1: return (b1 << 2) | (x02 << 3) | (b1 << 5) | (b2 << 6);
    case 2: return (x02 << 1) | (b1 << 2) | (a << 7);
    default: return (b1 << 2) | (x02 << 3) | (a << 10);
  }
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// message fragments (M[bra, ket] column-major: (a', a) at a' + CHI*a)
struct MsgFrag {
  double ma[2];  // M[g, t + 4j]   : A operand of "absorb first leg", B operand of the T-GEMM
  double mb[2];  // M[g, 2t + i]   : B operand of "absorb second leg" (register-chained)
};
__device__ __forceinline__ MsgFrag load_frag(const double* __restrict__ M, int g, int t) {
  MsgFrag f;
  f.ma[0] = M[g + CHI * t];
  f.ma[1] = M[g + CHI * (t + 4)];
  f.mb[0] = M[g + CHI * (2 * t)];
  f.mb[1] = M[g + CHI * (2 * t + 1)];
  return f;
}

// Phase 1:  dst[.., x', y', ..] = sum_{x,y} MX[x', x] MY[y', y] src[.., x, y, ..]   (legs X, Y absorbed;
// U, V are the two other legs).  64 (u, v) columns are split over the warps.
template <int X, int Y, int U, int V>
__device__ __forceinline__ void absorb_pair(const double* __restrict__ src, double* __restrict__ dst, const MsgFrag& mx,
                                            const MsgFrag& my, int warp, int g, int t) {
  const uint32_t ld0 = leg_pos(X, t) ^ leg_pos(Y, g), ld1 = leg_pos(X, t + 4) ^ leg_pos(Y, g);
  const uint32_t st0 = leg_pos(X, g) ^ leg_pos(Y, 2 * t), st1 = leg_pos(X, g) ^ leg_pos(Y, 2 * t + 1);
#pragma unroll 2
  for (int col = warp; col < CHI * CHI; col += NWARPS) {
    const uint32_t base = leg_pos(U, col & 7) ^ leg_pos(V, col >> 3);
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

// Phase 2:  out[v', v] = sum_{s, x', y', u'} conj(A[.., u', v']) * ( sum_u MU[u', u] P[.., u, v] )
// (leg U absorbed on the fly, leg V left open).  Returns this warp's partial 8x8 tile: thread (g, t) holds
// out[v' = g, v = 2t], out[v' = g, v = 2t + 1].
template <int X, int Y, int U, int V>
__device__ __forceinline__ void absorb_close(const double* __restrict__ P, const double* __restrict__ A, const MsgFrag& mu, int warp,
                                             int g, int t, double& o0, double& o1) {
  const uint32_t lp0 = leg_pos(U, t) ^ leg_pos(V, g), lp1 = leg_pos(U, t + 4) ^ leg_pos(V, g);
  const uint32_t la0 = leg_pos(U, 2 * t) ^ leg_pos(V, g), la1 = leg_pos(U, 2 * t + 1) ^ leg_pos(V, g);
  double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
#pragma unroll 2
  for (int col = warp; col < CHI * CHI; col += NWARPS) {
    const uint32_t base = leg_pos(X, col & 7) ^ leg_pos(Y, col >> 3);
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

struct Args {
  const VDesc* vdesc;
  const int32_t* vertices;  // the bucket's vertices owned by this rank
  int n_vertices;
  const int64_t* msg_off;
  const double* sites;
  const double* msg_in;
  double* msg_out;
  double* residual;
  int normalize;
};

constexpr size_t SMEM_BYTES = (size_t)(3 * NELEM + 2 * NWARPS * MSG + 2 * MSG) * sizeof(double);

// scatter the canonical tensor (HBM) into the swizzled shared-memory image, asynchronously
__device__ __forceinline__ void prefetch_site(double* dst, const double* __restrict__ src) {
  for (int c = threadIdx.x; c < NELEM / 2; c += NTHREADS) {
    const uint32_t pos = leg_pos(0, c & 7) ^ leg_pos(1, (c >> 3) & 7) ^ leg_pos(2, (c >> 6) & 7) ^ leg_pos(3, c >> 9);
    cp_async16(dst + pos, src + 2 * c);
  }
  cp_async_commit();
}

// cross-warp sum of two 8x8 partial tiles, then the fused normalise/residual/store epilogue
__device__ __forceinline__ void finish_pair(double* red, double* raw, int warp, int lane, int g, int t, double a0, double a1,
                                            double b0, double b1, int e_a, int e_b, const Args& k) {
  // thread (g, t) holds out[v' = g, v = 2t + i] -> element v' + CHI * v
  red[(0 * NWARPS + warp) * MSG + g + CHI * (2 * t)] = a0;
  red[(0 * NWARPS + warp) * MSG + g + CHI * (2 * t + 1)] = a1;
  red[(1 * NWARPS + warp) * MSG + g + CHI * (2 * t)] = b0;
  red[(1 * NWARPS + warp) * MSG + g + CHI * (2 * t + 1)] = b1;
  __syncthreads();
  if (threadIdx.x < 2 * MSG) {
    const int which = threadIdx.x / MSG, el = threadIdx.x % MSG;
    double s = 0;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) s += red[(which * NWARPS + w) * MSG + el];
    raw[which * MSG + el] = s;
  }
  __syncthreads();
  if (warp < 2) {
    const int e = warp == 0 ? e_a : e_b;
    const int64_t off = k.msg_off[e];
    warp_epilogue<double>(raw + warp * MSG, k.msg_in + off, k.msg_out + off, MSG, k.normalize,
                          k.residual ? k.residual + e : nullptr, lane);
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) bp_update_onchip_z4c8(Args k) {
  extern __shared__ __align__(16) double smem[];
  double* Pbuf = smem + 2 * NELEM;
  double* red = smem + 3 * NELEM;
  double* raw = red + 2 * NWARPS * MSG;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;

  int item = blockIdx.x;
  if (item >= k.n_vertices) return;
  int cur = 0;
  prefetch_site(smem, k.sites + k.vdesc[k.vertices[item]].site_off);
  for (; item < k.n_vertices; item += gridDim.x, cur ^= 1) {
    const VDesc& vd = k.vdesc[k.vertices[item]];
    const int next = item + gridDim.x;
    if (next < k.n_vertices) {
      prefetch_site(smem + (cur ^ 1) * NELEM, k.sites + k.vdesc[k.vertices[next]].site_off);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    const MsgFrag m0 = load_frag(k.msg_in + k.msg_off[vd.in_edge[0]], g, t);
    const MsgFrag m1 = load_frag(k.msg_in + k.msg_off[vd.in_edge[1]], g, t);
    const MsgFrag m2 = load_frag(k.msg_in + k.msg_off[vd.in_edge[2]], g, t);
    const MsgFrag m3 = load_frag(k.msg_in + k.msg_off[vd.in_edge[3]], g, t);
    __syncthreads();  // A_u landed (all threads' cp.async groups), previous vertex's epilogue done
    const double* A = smem + cur * NELEM;
    double a0, a1, b0, b1;

    // branch P = A·M0·M1  ->  out3 (absorb 2, close 3), out2 (absorb 3, close 2)
    absorb_pair<0, 1, 2, 3>(A, Pbuf, m0, m1, warp, g, t);
    __syncthreads();
    absorb_close<0, 1, 2, 3>(Pbuf, A, m2, warp, g, t, a0, a1);
    absorb_close<0, 1, 3, 2>(Pbuf, A, m3, warp, g, t, b0, b1);
    finish_pair(red, raw, warp, lane, g, t, a0, a1, b0, b1, vd.out_edge[3], vd.out_edge[2], k);

    // branch Q = A·M2·M3  ->  out1 (absorb 0, close 1), out0 (absorb 1, close 0)
    absorb_pair<2, 3, 0, 1>(A, Pbuf, m2, m3, warp, g, t);
    __syncthreads();
    absorb_close<2, 3, 0, 1>(Pbuf, A, m0, warp, g, t, a0, a1);
    absorb_close<2, 3, 1, 0>(Pbuf, A, m1, warp, g, t, b0, b1);
    finish_pair(red, raw, warp, lane, g, t, a0, a1, b0, b1, vd.out_edge[1], vd.out_edge[0], k);
  }
}

}  // namespace onchip
}  // namespace bpx
