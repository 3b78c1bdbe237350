// Generic (any degree / any dims / real or complex / NORM or SINGLE) kernels: the correctness
// baseline every bucket can fall back to ON THE GPU.  One CTA per work item (a directed edge for
// updates, a vertex for scalars); the partially absorbed tensor lives in shared memory when two copies
// fit, otherwise in a per-CTA global scratch that stays L2 resident.
//
// Restates, per work item, message_update!(::SimpleMessageUpdate, ...) of
// /root/reference/src/beliefpropagation/beliefpropagation.jl:242-257 with the absorption order
// (SURVEY.md §8 d3): absorb the k = z-1 incoming messages into the ket tensor one leg at a time, then
// close with conj(A) over the site leg and all absorbed legs.
#pragma once
#include "bpx_common.cuh"

namespace bpx {

struct GenericArgs {
  const VDesc* vdesc;
  const int32_t* src;      // per directed edge
  const int32_t* slot;     // per directed edge
  const int64_t* msg_off;  // per directed edge (+ total at [ne])
  const void* sites;
  const void* msg_in;
  void* msg_out;           // may alias msg_in (sequential schedule: in place)
  double* residual;        // per directed edge, may be NULL
  unsigned long long* resmax;  // the sweep's residual key (atomicMax), may be NULL
  const int32_t* work;     // list of work items (edge ids / vertex ids) or NULL for identity
  int64_t n_work;
  void* scratch;           // per-CTA global scratch: 2 * scratch_elems elements each
  int64_t scratch_elems;
  int smem_elems;          // elements per shared-memory tensor buffer (0: use global scratch)
  int normalize;
  int mode;                // BPX_MODE_*
  const void* ops;         // scalars: optional per-vertex d x d operators (packed), else NULL
  const int64_t* op_off;   // per-vertex element offsets into ops
  void* scalars_out;       // scalars: per work item
  unsigned long long stop_key;  // device-side convergence test (sweep_already_converged), 0: none
};

// out[l, a', r] = sum_a M[a' + chi_out * a] * cur[l + L * (a + chi * r)]
template <typename T>
__device__ __forceinline__ void absorb_leg(const T* __restrict__ cur, T* __restrict__ nxt, int64_t L, int chi,
                                           int chi_out, int64_t R, const T* __restrict__ M) {
  using E = Elem<T>;
  const int64_t n_out = L * chi_out * R;
  for (int64_t o = threadIdx.x; o < n_out; o += blockDim.x) {
    const int64_t l = o % L;
    const int64_t t = o / L;
    const int ap = (int)(t % chi_out);
    const int64_t r = t / chi_out;
    const T* c = cur + l + L * chi * r;
    T acc = E::zero();
    for (int a = 0; a < chi; ++a) acc = E::fma(M[ap + chi_out * a], c[L * a], acc);
    nxt[o] = acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bp_update_generic(GenericArgs g) {
  using E = Elem<T>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ VDesc vd;
  __shared__ int64_t cur_dim[BPX_MAX_DEGREE + 1];  // dims of the running tensor: [d, l_0..l_{z-1}]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const bool norm_mode = (g.mode == BPX_MODE_NORM);
  if (sweep_already_converged(g.resmax, g.stop_key)) return;

  T* buf[2];
  T* out_s;  // raw output message staged in shared memory (chi_s^2 or chi_s entries)
  if (g.smem_elems > 0) {
    buf[0] = reinterpret_cast<T*>(smem_raw);
    buf[1] = buf[0] + g.smem_elems;
    out_s = buf[1] + g.smem_elems;
  } else {
    buf[0] = reinterpret_cast<T*>(g.scratch) + (int64_t)blockIdx.x * 2 * g.scratch_elems;
    buf[1] = buf[0] + g.scratch_elems;
    out_s = reinterpret_cast<T*>(smem_raw);
  }

  for (int64_t w = blockIdx.x; w < g.n_work; w += gridDim.x) {
    const int e = g.work ? g.work[w] : (int)w;
    __syncthreads();  // previous item fully consumed
    if (threadIdx.x == 0) vd = g.vdesc[g.src[e]];
    __syncthreads();
    const int z = vd.z, slot = g.slot[e];
    if (threadIdx.x == 0) {
      cur_dim[0] = vd.d;
      for (int i = 0; i < z; ++i) cur_dim[i + 1] = vd.dim[i];
    }
    __syncthreads();
    const T* A = reinterpret_cast<const T*>(g.sites) + vd.site_off;
    const T* msg_in = reinterpret_cast<const T*>(g.msg_in);
    const T* cur = A;
    int which = 0;
    for (int i = 0; i < z; ++i) {
      if (i == slot) continue;
      int64_t L = 1, R = 1;
      for (int j = 0; j <= i; ++j) L *= cur_dim[j];
      for (int j = i + 2; j <= z; ++j) R *= cur_dim[j];
      const int chi = vd.dim[i];
      const int chi_out = norm_mode ? chi : 1;
      absorb_leg<T>(cur, buf[which], L, chi, chi_out, R, msg_in + g.msg_off[vd.in_edge[i]]);
      __syncthreads();
      if (threadIdx.x == 0) cur_dim[i + 1] = chi_out;
      cur = buf[which];
      which ^= 1;
      __syncthreads();
    }
    const int chi_s = vd.dim[slot];
    int nelem;
    if (norm_mode) {
      // out[b', b] = sum_{l, r} cur[l, b, r] * conj(A[l, b', r])
      int64_t L = vd.d, R = 1;
      for (int j = 0; j < slot; ++j) L *= vd.dim[j];
      for (int j = slot + 1; j < z; ++j) R *= vd.dim[j];
      const int64_t LR = L * R;
      nelem = chi_s * chi_s;
      for (int pair = warp; pair < nelem; pair += nwarps) {
        const int bp = pair % chi_s, b = pair / chi_s;
        T acc = E::zero();
        for (int64_t idx = lane; idx < LR; idx += 32) {
          const int64_t l = idx % L, r = idx / L;
          acc = E::fma(cur[l + L * (b + (int64_t)chi_s * r)], E::conj(A[l + L * (bp + (int64_t)chi_s * r)]), acc);
        }
        acc = warp_sum<T>(acc);
        if (lane == 0) out_s[pair] = acc;
      }
    } else {
      nelem = chi_s;
      for (int i = threadIdx.x; i < nelem; i += blockDim.x) out_s[i] = cur[i];
    }
    __syncthreads();
    if (warp == 0) {
      const int64_t off = g.msg_off[e];
      warp_epilogue<T>(out_s, msg_in + off, reinterpret_cast<T*>(g.msg_out) + off, nelem, g.normalize,
                       g.residual ? g.residual + e : nullptr, lane, g.resmax);
    }
  }
}

// vertex_scalar (messagecache.jl:139-143): factor with ALL z incoming messages absorbed; optional
// d x d operator on the ket site leg (numerator of a local expectation value).
template <typename T>
__global__ void __launch_bounds__(256) bp_vertex_scalar_generic(GenericArgs g) {
  using E = Elem<T>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ VDesc vd;
  __shared__ int64_t cur_dim[BPX_MAX_DEGREE + 1];
  __shared__ T red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const bool norm_mode = (g.mode == BPX_MODE_NORM);
  T* buf[2];
  if (g.smem_elems > 0) {
    buf[0] = reinterpret_cast<T*>(smem_raw);
    buf[1] = buf[0] + g.smem_elems;
  } else {
    buf[0] = reinterpret_cast<T*>(g.scratch) + (int64_t)blockIdx.x * 2 * g.scratch_elems;
    buf[1] = buf[0] + g.scratch_elems;
  }
  for (int64_t w = blockIdx.x; w < g.n_work; w += gridDim.x) {
    const int v = g.work ? g.work[w] : (int)w;
    __syncthreads();
    if (threadIdx.x == 0) {
      vd = g.vdesc[v];
      cur_dim[0] = vd.d;
      for (int i = 0; i < vd.z; ++i) cur_dim[i + 1] = vd.dim[i];
    }
    __syncthreads();
    const int z = vd.z;
    const T* A = reinterpret_cast<const T*>(g.sites) + vd.site_off;
    const T* msg_in = reinterpret_cast<const T*>(g.msg_in);
    const T* cur = A;
    int which = 0;
    for (int i = 0; i < z; ++i) {
      int64_t L = 1, R = 1;
      for (int j = 0; j <= i; ++j) L *= cur_dim[j];
      for (int j = i + 2; j <= z; ++j) R *= cur_dim[j];
      const int chi = vd.dim[i];
      const int chi_out = norm_mode ? chi : 1;
      absorb_leg<T>(cur, buf[which], L, chi, chi_out, R, msg_in + g.msg_off[vd.in_edge[i]]);
      __syncthreads();
      if (threadIdx.x == 0) cur_dim[i + 1] = chi_out;
      cur = buf[which];
      which ^= 1;
      __syncthreads();
    }
    if (norm_mode && g.ops) {
      // T'[s', rest] = sum_s op[s', s] T[s, rest]
      absorb_leg<T>(cur, buf[which], 1, vd.d, vd.d, vd.n / vd.d, reinterpret_cast<const T*>(g.ops) + g.op_off[v]);
      __syncthreads();
      cur = buf[which];
      which ^= 1;
    }
    T acc = E::zero();
    if (norm_mode) {
      for (int64_t i = threadIdx.x; i < vd.n; i += blockDim.x) acc = E::fma(cur[i], E::conj(A[i]), acc);
    } else {
      if (threadIdx.x == 0) acc = cur[0];
    }
    acc = warp_sum<T>(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      T s = E::zero();
      for (int i = 0; i < nwarps; ++i) s = E::add(s, red[i]);
      reinterpret_cast<T*>(g.scalars_out)[v] = s;  // indexed by vertex
    }
  }
}

// edge_scalar (messagecache.jl:153-157): sum_i M_e[i] * M_rev(e)[i] (no conjugation: same names).
template <typename T>
__global__ void bp_edge_scalar(const T* __restrict__ msgs, const int64_t* __restrict__ msg_off,
                               const int32_t* __restrict__ und_edge, const int32_t* __restrict__ rev, int64_t n_und,
                               T* __restrict__ out) {
  using E = Elem<T>;
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_und) return;
  const int e = und_edge[w], r = rev[e];
  const int64_t n = msg_off[e + 1] - msg_off[e];
  const T* a = msgs + msg_off[e];
  const T* b = msgs + msg_off[r];
  T acc = E::zero();
  for (int64_t i = lane; i < n; i += 32) acc = E::fma(a[i], b[i], acc);
  acc = warp_sum<T>(acc);
  if (lane == 0) out[w] = acc;
}

// vertex_scalar through the update kernels (messagecache.jl:139-143): the factor of v contracted with ALL its incoming
// messages equals  sum_i Mtilde_{v->w}[i] * M_{w->v}[i]  for any neighbour w, where Mtilde is the UNNORMALISED update output
// on the out-edge (everything but M_{w->v} absorbed) -- so one sweep of the bucket kernels with normalize = 0 into the idle
// message set plus this dot product gives every vertex scalar.  One warp per vertex: its first out-edge.
template <typename T>
__global__ void bp_vertex_belief(const T* __restrict__ unnormalised_out, const T* __restrict__ msgs_in, const int64_t* __restrict__ msg_off,
                                 const int32_t* __restrict__ rev, const int32_t* __restrict__ first_out_edge /* per vertex, -1: none */,
                                 const int32_t* __restrict__ vertices /* work list or NULL */, int64_t n_work, T* __restrict__ out) {
  using E = Elem<T>;
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_work) return;
  const int64_t v = vertices ? vertices[w] : w;
  const int e = first_out_edge[v];
  if (e < 0) return;  // isolated vertex: left to the generic scalar kernel
  const int r = rev[e];
  const int64_t n = msg_off[e + 1] - msg_off[e];
  const T* a = unnormalised_out + msg_off[e];
  const T* b = msgs_in + msg_off[r];
  T acc = E::zero();
  for (int64_t i = lane; i < n; i += 32) acc = E::fma(a[i], b[i], acc);
  acc = warp_sum<T>(acc);
  if (lane == 0) out[v] = acc;
}

// Per-edge term of iterate_diff (beliefpropagation.jl:261-267) between two message sets.
template <typename T>
__global__ void bp_edge_residual(const T* __restrict__ m1, const T* __restrict__ m2, const int64_t* __restrict__ msg_off,
                                 int64_t ne, double* __restrict__ residual) {
  using E = Elem<T>;
  const int lane = threadIdx.x & 31;
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= ne) return;
  const int64_t n = msg_off[e + 1] - msg_off[e];
  const T* a = m1 + msg_off[e];
  const T* b = m2 + msg_off[e];
  T dot = E::zero();
  double na = 0.0, nb = 0.0;
  for (int64_t i = lane; i < n; i += 32) {
    dot = E::fma(E::conj(a[i]), b[i], dot);
    na += E::abs2(a[i]);
    nb += E::abs2(b[i]);
  }
  dot = warp_sum<T>(dot);
  na = warp_sum_d(na);
  nb = warp_sum_d(nb);
  if (lane == 0) residual[e] = 1.0 - E::abs2(dot) / (na * nb);
}

// max over per-edge residuals (Julia `maximum`: NaN propagates), folded into the sweep's residual key (atomicMax,
// order independent => deterministic).  Only needed for edges updated by kernels that do not record the key
// themselves (the generic kernels and bp_edge_residual).
__global__ void __launch_bounds__(1024) bp_residual_max(const double* __restrict__ residual, const int32_t* __restrict__ list,
                                                        int64_t n, unsigned long long* __restrict__ slot) {
  __shared__ double red[32];
  __shared__ int nan_seen[32];
  double m = -INFINITY;
  int has_nan = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = residual[list ? list[i] : i];
    if (v != v) has_nan = 1;
    m = fmax(m, v);
  }
  for (int s = 16; s > 0; s >>= 1) {
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
    has_nan |= __shfl_xor_sync(0xffffffffu, has_nan, s);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    red[warp] = m;
    nan_seen[warp] = has_nan;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    m = lane < nw ? red[lane] : -INFINITY;
    has_nan = lane < nw ? nan_seen[lane] : 0;
    for (int s = 16; s > 0; s >>= 1) {
      m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
      has_nan |= __shfl_xor_sync(0xffffffffu, has_nan, s);
    }
    if (lane == 0 && n > 0) {
      if (has_nan) m = nan("");
      residual_record(slot, m);
    }
  }
}

// ---- device-side synthetic inputs (same counter-based recipe as the host bpx_fill_randn; libdevice log/cos may
// differ from libm in the last ulp, so parity tests always use host-generated data) ---------------------------
__device__ __forceinline__ unsigned long long splitmix64_dev(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ double randn_at_dev(unsigned long long seed, unsigned long long stream, unsigned long long i) {
  const unsigned long long base = splitmix64_dev(seed ^ splitmix64_dev(stream + 0x632BE59BD9B4E019ull));
  const unsigned long long a = splitmix64_dev(base + 2 * i), b = splitmix64_dev(base + 2 * i + 1);
  const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);
  const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

// site tensor v: randn(seed, stream = v, i) / sqrt(n_v)   (problems.synthetic_peps recipe); one CTA per vertex chunk
__global__ void fill_sites_randn(const VDesc* __restrict__ vdesc, int64_t nv, unsigned long long seed, int doubles_per_elem,
                                 double* __restrict__ sites) {
  for (int64_t v = blockIdx.y; v < nv; v += gridDim.y) {
    if (!vdesc[v].owned) continue;
    const int64_t n = vdesc[v].n * doubles_per_elem;
    const double scale = (doubles_per_elem == 2 ? 0.70710678118654752440 : 1.0) / sqrt((double)vdesc[v].n);
    double* dst = sites + vdesc[v].site_off * doubles_per_elem;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      dst[i] = scale * randn_at_dev(seed, (unsigned long long)v, (unsigned long long)i);
  }
}

// message e: (I + 0.1 |randn(seed, stream = nv + e, i)|) / sum   (NORM mode, "positive" init); warp per edge
__global__ void fill_messages_positive(const int64_t* __restrict__ msg_off, const int32_t* __restrict__ src_unused, int64_t ne,
                                       int64_t nv, unsigned long long seed, int doubles_per_elem, const int32_t* __restrict__ chi,
                                       double* __restrict__ msgs) {
  const int lane = threadIdx.x & 31;
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= ne) return;
  const int c = chi[e];
  const int64_t n = (int64_t)c * c;
  double* dst = msgs + msg_off[e] * doubles_per_elem;
  double s = 0.0;
  for (int64_t i = lane; i < n; i += 32) {
    const double v = ((i % c) == (i / c) ? 1.0 : 0.0) + 0.1 * fabs(randn_at_dev(seed, (unsigned long long)(nv + e), (unsigned long long)i));
    s += v;
  }
  s = warp_sum_d(s);
  for (int64_t i = lane; i < n; i += 32) {
    const double v = ((i % c) == (i / c) ? 1.0 : 0.0) + 0.1 * fabs(randn_at_dev(seed, (unsigned long long)(nv + e), (unsigned long long)i));
    if (doubles_per_elem == 1) {
      dst[i] = v / s;
    } else {
      dst[2 * i] = v / s;
      dst[2 * i + 1] = 0.0;
    }
  }
}

}  // namespace bpx
