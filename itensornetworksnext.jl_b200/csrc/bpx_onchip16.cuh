// BPX_KERNEL_ONCHIP, 16-wide variant: vertex-centric kernel for site tensors [s, b0, b1, b2] with d = 2 and three
// legs of dimension 16 (8192 doubles = 64 KiB resident in shared memory).  Two kinds of vertices map onto it:
//   * degree 3, chi = 16 (the boundary bucket of BASELINE config 5);
//   * degree 6, chi = 4 (BASELINE config 4, cubic lattice) in PAIR MODE: legs (2k, 2k+1) form super-leg k with
//     b_k = a_2k + 4 a_2k+1 -- the same memory -- and the super-message M_2k (x) M_2k+1 (Kronecker product, formed
//     on the fly while loading fragments).  The closure over super-leg k yields the 16x16 tile
//     S_k[(a', c'), (a, c)], from which the two real messages follow in the epilogue:
//         out_2k[a', a] = sum_{c', c} M_2k+1[c', c] S_k[(a', c'), (a, c)],   out_2k+1 likewise with M_2k.
//     8 GEMM units of d*16^4 MACs per vertex (1.05 M) instead of 36 units of d*4^7 for six independent updates.
// Leave-one-out tree:  X = A·M0 -> out2 (absorb 1, close 2), out1 (absorb 2, close 1);  X' = A·M2 -> out0.
// Device functions, layout (L_A3: a degree-3 tensor is exactly an "a3 = 0" slice) and DMMA chaining are those of
// bpx_sliced.cuh; TMA producer warp + double-buffered A as in bpx_onchip.cuh.
#pragma once
#include "bpx_sliced.cuh"

namespace bpx {
namespace onchip16 {

using namespace sliced;  // pos<>, FragA, absorb_close16, dmma, mbarrier/TMA helpers, constants

constexpr int NCW16 = 8;
constexpr int NCT16 = NCW16 * 32;
constexpr int NEW16 = 2;                          // epilogue warps
constexpr int NTHREADS16 = NCT16 + 32 * NEW16 + 32;  // + producer warp
constexpr int NRAW16 = NCT16 + 32 * NEW16;       // participants of the raw-tile hand-over
constexpr int NEL = 8192;

struct ItemDesc {
  int64_t site_off;
  int64_t in_off[6];   // plain mode: legs 0..2; pair mode: legs 0..5
  int64_t out_off[6];
  int32_t out_edge[6];
  int32_t peer[6];
  int32_t pair_mode;   // 0: three legs of dimension 16; 1: six legs of dimension 4 paired into super-legs
  int32_t pad;
  int64_t need;        // streamed host I/O: message-set prefix (elements) that holds every message this item reads
};

struct Args {
  const ItemDesc* items;
  int n_items;
  const double* sites;  // pre-swizzled image (layout L_A3)
  const double* msg_in;
  double* msg_out;
  double* residual;
  unsigned long long* resmax;
  int normalize;
  PeerArgs peer;
  HostIO io;  // streamed host I/O (bpx_sweep_host), all NULL otherwise
  unsigned long long stop_key;  // device-side convergence test (sweep_already_converged), 0: none
};

// staged incoming messages: plain 3 x 256 doubles, pair mode 6 x 16 doubles; super-message element (b', b)
__device__ __forceinline__ double msg_elem(const double* M, int pair_mode, int leg, int bp, int b) {
  if (!pair_mode) return M[leg * MSG + bp + CHI * b];
  const double* lo = M + (2 * leg) * 16;
  const double* hi = lo + 16;
  return lo[(bp & 3) + 4 * (b & 3)] * hi[(bp >> 2) + 4 * (b >> 2)];
}
__device__ __forceinline__ FragA load_fragA16(const double* M, int pair_mode, int leg, int g, int t) {
  FragA f;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int j = 0; j < 4; ++j) f.v[mt][j] = msg_elem(M, pair_mode, leg, g + 8 * mt, t + 4 * j);
  return f;
}

// dst[x', y, c] = sum_x MX[x', x] src[x, y, c]   (one column c of the spectator leg; dst != src)
template <int X, int Y>
__device__ __forceinline__ void absorb_one16(const double* src, double* dst, uint32_t base, const FragA& mx, int g, int t) {
  double2 b[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int h = 0; h < 2; ++h) b[j][h] = *reinterpret_cast<const double2*>(src + (base ^ pos<L_A3>(X, t + 4 * j) ^ pos<L_A3>(Y, g + 8 * h)));
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double p0 = 0, p1 = 0, q0 = 0, q1 = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dmma(p0, p1, mx.v[mt][j], b[j][h].x);
        dmma(q0, q1, mx.v[mt][j], b[j][h].y);
      }
      const uint32_t a = base ^ pos<L_A3>(X, g + 8 * mt);
      *reinterpret_cast<double2*>(dst + (a ^ pos<L_A3>(Y, 2 * t + 8 * h))) = make_double2(p0, q0);
      *reinterpret_cast<double2*>(dst + (a ^ pos<L_A3>(Y, 2 * t + 1 + 8 * h))) = make_double2(p1, q1);
    }
}

// shared memory (doubles): A[2][NEL] | X[NEL] | red[NCW16][256] | raw[256] | msgs[2][768] | 2 mbarriers
constexpr size_t SMEM_DOUBLES16 = (size_t)3 * NEL + NCW16 * MSG + MSG + 2 * 3 * MSG + 2;
constexpr size_t SMEM_BYTES16 = SMEM_DOUBLES16 * sizeof(double);
enum { BAR_C16 = 1, BAR_SLOT16 = 2 /* and 3 */, BAR_RAW_FULL16 = 4, BAR_RAW_FREE16 = 5 };

__global__ void swizzle_sites_z3(const ItemDesc* items, int n_items, const double* __restrict__ src, double* __restrict__ dst) {
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int64_t off = items[item].site_off;
    for (int c = threadIdx.x; c < NEL / 2; c += blockDim.x) {
      const uint32_t p = global_pos(0, c & 15) ^ global_pos(1, (c >> 4) & 15) ^ global_pos(2, (c >> 8) & 15);
      *reinterpret_cast<double2*>(dst + off + p) = *reinterpret_cast<const double2*>(src + off + 2 * c);
    }
  }
}

// compute warps: cross-warp sum of a 16x16 partial tile -> raw, handed to the epilogue warps.
// The partial tiles are stored in FRAGMENT order (value kk = i + 2 h + 4 mt of lane l at l + 32 kk): lane-contiguous,
// conflict-free stores.  (In the natural order (g + 8 mt) + 16 (2 t + i + 8 h) the four t of a row land on one bank: every
// one of the 8 stores of every warp was a 4-way conflict -- 4.7 M excessive wavefronts per cfg4 sweep,
// profiles/r2n_onchip_smem_conflicts.txt.)  Thread el = 32 kk + l sums element el of the NCW16 tiles and puts the result
// at its natural position in `raw`, which the epilogue warps read.
__device__ __forceinline__ void publish16(double* red, double* raw, const double (&acc)[2][2][2], int warp, int g, int t) {
  double* mine = red + warp * MSG;
  const int lane = 4 * g + t;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < 2; ++i) mine[lane + 32 * (i + 2 * h + 4 * mt)] = acc[mt][h][i];
  onchip::bar_sync(BAR_C16, NCT16);  // partial tiles visible; every compute warp is done with X / A of this phase
  const int el = threadIdx.x;        // = 32 warp + lane for the compute warps: fragment value kk = warp of lane `lane`
  double s = 0;
#pragma unroll
  for (int w = 0; w < NCW16; ++w) s += red[w * MSG + el];
  onchip::bar_sync(BAR_RAW_FREE16, NRAW16);  // epilogue warps have consumed the previous tile (also: red fully read)
  const int kk = warp, ki = kk & 1, kh = (kk >> 1) & 1, kmt = kk >> 2;
  raw[(g + 8 * kmt) + CHI * (2 * t + ki + 8 * kh)] = s;
  onchip::bar_arrive(BAR_RAW_FULL16, NRAW16);
}

// epilogue warps: tile of super-leg `sleg` -> message(s): sum-normalise, residual, store (+ peer store)
__device__ __forceinline__ void epilogue16(const double* raw, const double* Mst, int which, int lane, const Args& k, const ItemDesc* d,
                                           int sleg) {
  if (!d->pair_mode) {
    // plain: one 16x16 message, handled by epilogue warp 0; lane holds elements lane + 32 j
    const int64_t off = d->out_off[sleg];
    double o[8], v[8];
    if (which == 0) {
      hostio_wait(k.io, d->need);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = k.msg_in[off + lane + 32 * j];  // issued before the hand-over
    }
    onchip::bar_sync(BAR_RAW_FULL16, NRAW16);
    if (which == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = raw[lane + 32 * j];
    }
    onchip::bar_arrive(BAR_RAW_FREE16, NRAW16);
    if (which != 0) return;
    double s = 0, dot = 0, n_old = 0, n_new = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s += v[j];
      dot += o[j] * v[j];
      n_old += o[j] * o[j];
      n_new += v[j] * v[j];
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, m);
      dot += __shfl_xor_sync(0xffffffffu, dot, m);
      n_old += __shfl_xor_sync(0xffffffffu, n_old, m);
      n_new += __shfl_xor_sync(0xffffffffu, n_new, m);
    }
    const bool scale = k.normalize && s != 0.0;
    double* peer_m = (k.peer.nranks > 1 && d->peer[sleg] >= 0) ? k.peer.peer_out[d->peer[sleg]] + off : nullptr;
    double* host_m = k.io.host_out ? k.io.host_out + off : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double x = scale ? v[j] / s : v[j];
      k.msg_out[off + lane + 32 * j] = x;
      if (peer_m) peer_m[lane + 32 * j] = x;
      if (host_m) host_m[lane + 32 * j] = x;
    }
    if (peer_m) __threadfence_system();  // released here instead of at the kernel's tail
    if (lane == 0) residual_record(k.resmax, 1.0 - dot * dot / (n_old * n_new));  // invariant under the scaling
  } else {
    // pair mode: S[(a', c'), (a, c)] at (a' + 4 c') + 16 (a + 4 c).
    // warp 0: out_2k[a', a] = sum M_2k+1[c', c] S;  warp 1: out_2k+1[c', c] = sum M_2k[a', a] S
    const int leg = 2 * sleg + which;
    const int64_t off = d->out_off[leg];
    const double* Mo = Mst + (2 * sleg + (1 - which)) * 16;  // the OTHER message of the pair (staged copy)
    hostio_wait(k.io, d->need);
    const double o = lane < 16 ? k.msg_in[off + lane] : 0.0;
    double mo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) mo[i] = Mo[i];
    onchip::bar_sync(BAR_RAW_FULL16, NRAW16);
    double v = 0.0;
    if (lane < 16) {
      const int p = lane & 3, q = lane >> 2;  // output element (p, q) of a 4x4 message
#pragma unroll
      for (int cp = 0; cp < 4; ++cp)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int row = which == 0 ? (p + 4 * cp) : (cp + 4 * p);
          const int col = which == 0 ? (q + 4 * c) : (c + 4 * q);
          v += mo[cp + 4 * c] * raw[row + 16 * col];
        }
    }
    onchip::bar_arrive(BAR_RAW_FREE16, NRAW16);
    double s = v, dot = o * v, n_old = o * o, n_new = v * v;
#pragma unroll
    for (int m = 8; m > 0; m >>= 1) {  // lanes >= 16 hold zeros
      s += __shfl_xor_sync(0xffffffffu, s, m);
      dot += __shfl_xor_sync(0xffffffffu, dot, m);
      n_old += __shfl_xor_sync(0xffffffffu, n_old, m);
      n_new += __shfl_xor_sync(0xffffffffu, n_new, m);
    }
    if (lane < 16) {
      const double x = (k.normalize && s != 0.0) ? v / s : v;
      k.msg_out[off + lane] = x;
      if (k.peer.nranks > 1 && d->peer[leg] >= 0) {
        k.peer.peer_out[d->peer[leg]][off + lane] = x;
        __threadfence_system();
      }
      if (k.io.host_out) k.io.host_out[off + lane] = x;
    }
    if (lane == 0) residual_record(k.resmax, 1.0 - dot * dot / (n_old * n_new));
  }
}

__global__ void __launch_bounds__(NTHREADS16, 1) bp_update_onchip_c16(Args k) {
  extern __shared__ __align__(128) double smem[];
  double* Xbuf = smem + 2 * NEL;
  double* red = smem + 3 * NEL;
  double* raw = red + NCW16 * MSG;
  double* msgs = raw + MSG;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(msgs + 2 * 3 * MSG);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  if (sweep_already_converged(k.resmax, k.stop_key)) return;
  const int G = gridDim.x;
  if ((int)blockIdx.x >= k.n_items) return;
  if (threadIdx.x == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp == NCW16 + NEW16) {
    // ===== producer: A (one 64 KiB run) + incoming messages of item n into slot n & 1 =====
    peer_gate(k.peer, lane);
    int n = 0;
    for (int item = blockIdx.x; item < k.n_items; item += G, ++n) {
      const int sl = n & 1;
      if (n >= 2) onchip::bar_sync(BAR_SLOT16 + sl, NRAW16 + 32);  // compute AND epilogue warps released the slot
      const ItemDesc* d = k.items + item;
      hostio_wait(k.io, d->need);  // streamed upload: the item's messages have arrived
      fence_proxy_async();
      if (lane == 0) mbar_expect_tx(&mbar[sl], NEL * 8 + (d->pair_mode ? 6 * 16 * 8 : 3 * MSG * 8));
      __syncwarp();
      if (lane < 4) tma_bulk_g2s(smem + sl * NEL + lane * 2048, k.sites + d->site_off + lane * 2048, 16384, &mbar[sl]);
      double* Md = msgs + sl * 3 * MSG;
      if (d->pair_mode) {
        if (lane >= 8 && lane < 14) tma_bulk_g2s(Md + (lane - 8) * 16, k.msg_in + d->in_off[lane - 8], 128, &mbar[sl]);
      } else {
        if (lane >= 8 && lane < 11) tma_bulk_g2s(Md + (lane - 8) * MSG, k.msg_in + d->in_off[lane - 8], MSG * 8, &mbar[sl]);
      }
    }
  } else if (warp >= NCW16) {
    // ===== epilogue warps: tiles arrive in the order out2, out1, out0 of every item =====
    const int which = warp - NCW16;
    onchip::bar_arrive(BAR_RAW_FREE16, NRAW16);  // raw starts free
    int n = 0;
    for (int item = blockIdx.x; item < k.n_items; item += G, ++n) {
      const ItemDesc* d = k.items + item;
      const double* Mst = msgs + (n & 1) * 3 * MSG;  // staged incoming messages (valid until the item's slot is released)
      if (d->pair_mode) mbar_wait(&mbar[n & 1], (n >> 1) & 1);  // pair mode reads the staged messages
      epilogue16(raw, Mst, which, lane, k, d, 2);
      epilogue16(raw, Mst, which, lane, k, d, 1);
      epilogue16(raw, Mst, which, lane, k, d, 0);
      if (item + 2 * G < k.n_items) onchip::bar_arrive(BAR_SLOT16 + (n & 1), NRAW16 + 32);
    }
  } else {
    // ===== compute warps =====
    int n = 0;
    for (int item = blockIdx.x; item < k.n_items; item += G, ++n) {
      const int sl = n & 1;
      const ItemDesc* d = k.items + item;
      const int pm = d->pair_mode;
      mbar_wait(&mbar[sl], (n >> 1) & 1);
      const double* A = smem + sl * NEL;
      const double* M = msgs + sl * 3 * MSG;
      double accA[2][2][2], accB[2][2][2];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) accA[a][b][0] = accA[a][b][1] = accB[a][b][0] = accB[a][b][1] = 0.0;
      {
        // X = A·M0 (columns: leg 2)  ->  out2 (absorb 1, close 2), out1 (absorb 2, close 1)  (columns: leg 0')
        const FragA m0 = load_fragA16(M, pm, 0, g, t);
#pragma unroll 1
        for (int c = warp; c < 16; c += NCW16) absorb_one16<0, 1>(A, Xbuf, pos<L_A3>(2, c), m0, g, t);
        onchip::bar_sync(BAR_C16, NCT16);
        const FragA m1 = load_fragA16(M, pm, 1, g, t), m2 = load_fragA16(M, pm, 2, g, t);
#pragma unroll 1
        for (int c = warp; c < 16; c += NCW16) {
          absorb_close16<L_A3, 1, 2>(Xbuf, A, pos<L_A3>(0, c), m1, g, t, accA);
          absorb_close16<L_A3, 2, 1>(Xbuf, A, pos<L_A3>(0, c), m2, g, t, accB);
        }
      }
      publish16(red, raw, accA, warp, g, t);  // out2; its first barrier also frees Xbuf
      publish16(red, raw, accB, warp, g, t);  // out1
      {
        // X' = A·M2 (columns: leg 0)  ->  out0 (absorb 1, close 0)  (columns: leg 2')
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b) accA[a][b][0] = accA[a][b][1] = 0.0;
        const FragA m2 = load_fragA16(M, pm, 2, g, t);
#pragma unroll 1
        for (int c = warp; c < 16; c += NCW16) absorb_one16<2, 1>(A, Xbuf, pos<L_A3>(0, c), m2, g, t);
        onchip::bar_sync(BAR_C16, NCT16);
        const FragA m1 = load_fragA16(M, pm, 1, g, t);
#pragma unroll 1
        for (int c = warp; c < 16; c += NCW16) absorb_close16<L_A3, 1, 0>(Xbuf, A, pos<L_A3>(2, c), m1, g, t, accA);
      }
      publish16(red, raw, accA, warp, g, t);  // out0
      if (item + 2 * G < k.n_items) onchip::bar_arrive(BAR_SLOT16 + sl, NRAW16 + 32);
    }
    onchip::bar_sync(BAR_RAW_FREE16, NRAW16);  // let the epilogue warps' last arrive complete
  }
  peer_post_when_last(k.peer, false);  // peer stores were released where they were issued
  hostio_finish(k.io);
}

}  // namespace onchip16
}  // namespace bpx
