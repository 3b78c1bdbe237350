// Two-site expectation values in the BP environment: <O_e> for a batch of edges e = (v1, v2), the quantity a simple-update
// evolution monitors (bond energies).  BUILD-DEFINED EXTENSION like the one-site `bpx_vertex_expect_numerators`: the
// reference has no `expect`; the contraction is the two-vertex analogue of `vertex_scalar`
// (src/beliefpropagation/messagecache.jl:139-143) -- both factors of the norm network (normnetwork.jl:49-54) with every
// incoming message EXCEPT the two on the shared link, the operator applied to the two ket site legs:
//
//   N_a[(s, b), (s', b')] = sum_{ext, ext'} A_a[s, b, ext] conj(A_a[s', b', ext']) prod_i M_{w_i -> a}[ext'_i, ext_i]
//   rho[(s1 s2), (s1' s2')] = sum_{b, b'} N_1[(s1, b), (s1', b')] N_2[(s2, b), (s2', b')]
//   num = sum O[s1', s2', s1, s2] rho[(s1 s2), (s1' s2')],     den = tr rho,     <O_e> = num / den
//
// No factorisation of the messages is needed (the gates' gauges X come from an eigen-decomposition; here the messages are
// absorbed as they are): mode products on the external legs, one Gram-like product per vertex, a small double sum.
// Written against the same `Team` abstraction as bpx_apply.cuh (host-compiled tests, ThreadSanitizer schedule check);
// one CTA per edge; read-only on tensors and messages, so the edges of a batch may share vertices.
// STATUS: written after round 1's GPU budget had ended -- verified on the host only.
#pragma once
#include "bpx_apply.cuh"

namespace bpx {
namespace expect2 {

using namespace bpx::applyk;

template <typename T>
__host__ __device__ __forceinline__ T from_parts(double re, double im);
template <>
__host__ __device__ __forceinline__ double from_parts<double>(double re, double) { return re; }
template <>
__host__ __device__ __forceinline__ c64 from_parts<c64>(double re, double im) { return make_c64(re, im); }

struct EdgeDesc {
  Side s[2];
  int64_t op_off;  // element offset of O[o1, o2, i1, i2] in the packed operator buffer
  int64_t ws_off;
  int32_t chi_b;
  int32_t pad_;
};

// work space per edge (elements of T): per side the matrix view of A, two ping-pong buffers for the absorbed tensor,
// and N_a (cols x cols)
struct LayoutE {
  int64_t a0[2], t0[2], t1[2], n[2];
  int64_t total;
};
__host__ __device__ inline LayoutE layout_of(const EdgeDesc& g) {
  LayoutE L;
  int64_t o = 0;
  for (int a = 0; a < 2; ++a) {
    const Side& s = g.s[a];
    L.a0[a] = o; o += s.n;
    L.t0[a] = o; o += s.n;
    L.t1[a] = o; o += s.n;
    L.n[a] = o; o += (int64_t)s.cols * s.cols;
  }
  L.total = o;
  return L;
}

// N[c, c'] = sum_row T[row, c] conj(P[row, c']): one warp per (c, c') pair, lanes over the rows
template <typename T>
__host__ __device__ void gram_pair(const Team& tm, const T* Tm, const T* P, int64_t rows, int cols, T* N) {
  using E = Elem<T>;
  const int L = tm.lanes();
  for (int i = tm.wid; i < cols * cols; i += tm.nw) {
    const int c = i % cols, cp = i / cols;
    const T* x = Tm + rows * c;
    const T* y = P + rows * cp;
    T acc = E::zero();
    for (int64_t r = tm.lane; r < rows; r += L) acc = E::fma(x[r], E::conj(y[r]), acc);
    acc = tm.template sum_t<T>(acc);
    if (tm.lane == 0) N[i] = acc;
  }
  tm.sync();
}

// accum: 4 doubles visible to the whole team (shared memory on the device): num.re, num.im, den.re, den.im
template <typename T>
__host__ __device__ void run_edge(const Team& tm, const EdgeDesc& gd, const T* sites, const T* msgs, const T* ops, T* ws,
                                  T* num_out, T* den_out, double* accum) {
  using E = Elem<T>;
  const LayoutE L = layout_of(gd);
  T* w = ws + gd.ws_off;
  for (int a = 0; a < 2; ++a) {
    const Side& sd = gd.s[a];
    T* P = w + L.a0[a];
    T* b[2] = {w + L.t0[a], w + L.t1[a]};
    tensor_to_matrix<T>(tm, sd, sites + sd.site_off, P);
    const T* cur = P;
    int nxt = 0;
    int64_t st = 1;
    for (int i = 0; i < sd.z; ++i) {
      if (i == sd.bond_slot) continue;
      const int chi = sd.dim[i];
      // out[.., a', ..] = sum_a M[a', a] in[.., a, ..]: the message [bra, ket] absorbed on the ket leg
      mode_product<T>(tm, cur, b[nxt], sd.rows, sd.cols, st, chi, msgs + sd.in_msg[i]);
      cur = b[nxt];
      nxt ^= 1;
      st *= chi;
    }
    gram_pair<T>(tm, cur, P, sd.rows, sd.cols, w + L.n[a]);
  }
  if (tm.tid() == 0) accum[0] = accum[1] = accum[2] = accum[3] = 0.0;
  tm.sync();
  const Side& s1 = gd.s[0];
  const Side& s2 = gd.s[1];
  const int d1 = s1.d, d2 = s2.d, chi = gd.chi_b, c1 = s1.cols, c2 = s2.cols, dd = d1 * d2;
  const T* N1 = w + L.n[0];
  const T* N2 = w + L.n[1];
  const T* op = ops + gd.op_off;
  T num = E::zero(), den = E::zero();
  // one (b, b') pair per thread: rho's contribution, contracted with O (num) and with the identity (den)
  for (int i = tm.tid(); i < chi * chi; i += tm.nt()) {
    const int b = i % chi, bp = i / chi;
    for (int x2 = 0; x2 < d2; ++x2)
      for (int x1 = 0; x1 < d1; ++x1)
        for (int y2 = 0; y2 < d2; ++y2)
          for (int y1 = 0; y1 < d1; ++y1) {
            // ket (x1, x2), bra (y1, y2)
            const T r = E::mul(N1[(x1 + d1 * b) + c1 * (y1 + d1 * bp)], N2[(x2 + d2 * b) + c2 * (y2 + d2 * bp)]);
            num = E::fma(op[y1 + d1 * y2 + dd * (x1 + d1 * x2)], r, num);
            if (x1 == y1 && x2 == y2) den = E::add(den, r);
          }
  }
  num = tm.template sum_t<T>(num);
  den = tm.template sum_t<T>(den);
  if (tm.lane == 0) {
#ifdef __CUDA_ARCH__
    atomicAdd(accum + 0, real_of(num));
    atomicAdd(accum + 1, imag_of(num));
    atomicAdd(accum + 2, real_of(den));
    atomicAdd(accum + 3, imag_of(den));
#else
    BPX_HOST_ATOMIC_ADD(accum + 0, real_of(num));
    BPX_HOST_ATOMIC_ADD(accum + 1, imag_of(num));
    BPX_HOST_ATOMIC_ADD(accum + 2, real_of(den));
    BPX_HOST_ATOMIC_ADD(accum + 3, imag_of(den));
#endif
  }
  tm.sync();
  if (tm.tid() == 0) {
    *num_out = from_parts<T>(accum[0], accum[1]);
    *den_out = from_parts<T>(accum[2], accum[3]);
  }
  tm.sync();
}

#ifdef __CUDACC__
struct ExpectArgs {
  const EdgeDesc* edges;
  int64_t n_edges;
  const void* sites;
  const void* msgs;
  const void* ops;
  void* ws;
  void* num_out;
  void* den_out;
};

template <typename T>
__global__ void __launch_bounds__(NT) bp_edge_expect(ExpectArgs a) {
  __shared__ double accum[4];
  Team tm;
  tm.lane = threadIdx.x & 31;
  tm.wid = threadIdx.x >> 5;
  tm.nw = NT / 32;
  for (int64_t g = blockIdx.x; g < a.n_edges; g += gridDim.x) {
    run_edge<T>(tm, a.edges[g], static_cast<const T*>(a.sites), static_cast<const T*>(a.msgs), static_cast<const T*>(a.ops),
                static_cast<T*>(a.ws), static_cast<T*>(a.num_out) + g, static_cast<T*>(a.den_out) + g, accum);
    __syncthreads();
  }
}
#endif

}  // namespace expect2
}  // namespace bpx
