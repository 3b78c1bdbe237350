// Vertex-centric register kernel for SINGLE-layer networks with small uniform link dimensions: the HBM-bound bucket
// family of the path (Ising / spin-ice / delta networks of src/ITensorNetworkGenerators, chi = 2).
//
// One update on a single-layer network (beliefpropagation.jl:242-257 with plain ITensorNetwork factors):
//     m~_{u->v_j}[b] = sum_{l : l_j = b} T_u[l_0..l_{z-1}] prod_{k != j} m_{w_k->u}[l_k]
// followed by the sum-normalisation (:248-253) and the per-edge term of iterate_diff (:261-267).
//
// Arithmetic intensity is O(z) flop per byte of T_u (a chi = 2, degree-4 factor is 128 B and costs ~160 flops for all
// four out-messages), far below the FP64 ridge, so the kernel is organised around memory traffic only:
//   * ONE THREAD PER VERTEX.  The factor is read once (16-byte loads, one full 128 B line per thread at chi = 2,
//     z = 4) and ALL z leave-one-out contractions are formed from it in registers (prefix products over the legs,
//     fully unrolled: every index of the loops below is a compile-time constant), so T_u is read once per sweep
//     instead of z times and no intermediate touches shared or global memory.
//   * Descriptors are a structure of arrays indexed by the position in the bucket (site offset, z in-message offsets,
//     z out-message offsets: 8 + 8 z bytes per vertex), read fully coalesced -- the generic kernel's 176-byte VDesc per
//     update would cost more traffic than the factor itself.
//   * The residual maximum is reduced per thread -> warp (shuffles) -> CTA (shared memory) and leaves the CTA as ONE
//     atomicMax on the sweep's order-preserving key (bpx_common.cuh); a per-message atomic on one address would
//     serialise millions of updates.
// Algorithmic bytes per vertex: chi^z w (factor) + 3 z chi w (message in, old, out) -- the roofline `bench.py
// --workload ising` reports against (plus the 8 + 8 z descriptor bytes, which are real traffic but not algorithmic).
#pragma once
#include <stdlib.h>

#include "bpx_common.cuh"

namespace bpx {
namespace vertexk {

constexpr int NT = 128;  // threads per CTA
#ifndef BPX_VERTEX_MIN_CTAS
#define BPX_VERTEX_MIN_CTAS 6  // resident CTAs per SM the small Float64 shapes are compiled for (register cap 80; 5 -> 96 measured 2 % slower)
#endif

struct Args {
  const int64_t* site;     // [n]      element offset of T_v in the device site buffer
  const int32_t* moff;     // [2 z][n] element offsets of the z incoming messages, then of the z outgoing messages
  const void* sites;
  const void* msg_in;
  void* msg_out;
  unsigned long long* resmax;  // the sweep's residual key (may be NULL)
  int64_t n;               // vertices in this launch
  int normalize;
  int out_contig;          // the z out-messages of every vertex are adjacent in the packed message layout, slot order
  unsigned long long stop_key;  // device-side convergence test (sweep_already_converged), 0: none
};

__host__ __device__ constexpr int ipow(int b, int e) { return e <= 0 ? 1 : b * ipow(b, e - 1); }

// shapes that stay in registers: chi^z values of at most 64 doubles (complex counts twice)
template <typename T, int Z, int N>
__host__ __device__ constexpr bool supported() {
  return Z >= 1 && Z <= 6 && N >= 2 && N <= 4 && ipow(N, Z) * (int)(sizeof(T) / 8) <= 64;
}
inline bool shape_supported(bool is_complex, int z, int n) {
  if (z < 1 || z > 6 || n < 2 || n > 4) return false;
  int64_t e = 1;
  for (int i = 0; i < z; ++i) e *= n;
  return e * (is_complex ? 2 : 1) <= 64;
}

// COUNT elements from global memory into registers; 16-byte loads when the run is 16-byte aligned
template <typename T, int COUNT>
__device__ __forceinline__ void load_run(const T* __restrict__ p, T (&dst)[COUNT]) {
  if constexpr (sizeof(T) == 16) {
    const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int i = 0; i < COUNT; ++i) {
      const double2 v = __ldg(q + i);
      dst[i] = make_c64(v.x, v.y);
    }
  } else {
    if constexpr (COUNT % 2 == 0) {
      if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
        for (int i = 0; i < COUNT / 2; ++i) {
          const double2 v = __ldg(q + i);
          dst[2 * i] = v.x;
          dst[2 * i + 1] = v.y;
        }
        return;
      }
    }
#pragma unroll
    for (int i = 0; i < COUNT; ++i) dst[i] = __ldg(p + i);
  }
}

template <typename T, int COUNT>
__device__ __forceinline__ void store_run(T* __restrict__ p, const T (&src)[COUNT]) {
  if constexpr (sizeof(T) == 16) {
    double2* q = reinterpret_cast<double2*>(p);
#pragma unroll
    for (int i = 0; i < COUNT; ++i) q[i] = make_double2(src[i].re, src[i].im);
  } else {
    if constexpr (COUNT % 2 == 0) {
      if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        double2* q = reinterpret_cast<double2*>(p);
#pragma unroll
        for (int i = 0; i < COUNT / 2; ++i) q[i] = make_double2(src[2 * i], src[2 * i + 1]);
        return;
      }
    }
#pragma unroll
    for (int i = 0; i < COUNT; ++i) p[i] = src[i];
  }
}

// geometry of the staged (warp-cooperative) path: a warp's 32 factors (and its 32 runs of z out-messages) are contiguous
// in HBM when the bucket's vertices are consecutive, so the warp moves them with fully coalesced 16-byte accesses and
// transposes through shared memory (row = one vertex; odd row stride in 16-byte units: conflict-free LDS/STS.128)
template <typename T, int Z, int N>
struct Shape {
  static constexpr int NE = ipow(N, Z);
  static constexpr int RB = NE * (int)sizeof(T);       // bytes of one factor
  static constexpr int OB = Z * N * (int)sizeof(T);    // bytes of the z out-messages of one vertex
  static constexpr bool STAGED = (RB % 16 == 0) && (OB % 16 == 0);
  static constexpr int RC = RB / 16, RS = RC + ((RC & 1) ? 2 : 1);
  static constexpr int OC = OB / 16, OS = OC + ((OC & 1) ? 2 : 1);
  static constexpr int WARP_UNITS = STAGED ? 32 * (RS + OS) : 0;  // 16-byte units of shared memory per warp
  static constexpr int SMEM_BYTES = (NT / 32) * WARP_UNITS * 16;
  // small factors: cap the registers at 96 (register files are allocated in units of 256 per warp: 100 registers would
  // cost a fifth resident CTA)
  static constexpr int MIN_CTAS = (NE * (int)(sizeof(T) / 8) <= 16) ? (sizeof(T) == 8 ? BPX_VERTEX_MIN_CTAS : 5) : 1;
};

// all z leave-one-out contractions in one pass over the factor:
//   pre_j = T[l] prod_{k < j} m_k[l_k],  suf_j = prod_{k > j} m_k[l_k],  raw_j[l_j] += pre_j suf_j
template <typename T, int Z, int N>
__device__ __forceinline__ void leave_one_out(const T (&t)[Shape<T, Z, N>::NE], const T (&m)[Z][N], T (&raw)[Z][N]) {
  using E = Elem<T>;
  constexpr int NE = Shape<T, Z, N>::NE;
#pragma unroll
  for (int j = 0; j < Z; ++j)
#pragma unroll
    for (int b = 0; b < N; ++b) raw[j][b] = E::zero();
#pragma unroll
  for (int x = 0; x < NE; ++x) {
    int dig[Z];  // l_k of element x: constants once the loop is unrolled
    {
      int r = x;
#pragma unroll
      for (int k = 0; k < Z; ++k) {
        dig[k] = r % N;
        r /= N;
      }
    }
    T pre[Z];
    pre[0] = t[x];
#pragma unroll
    for (int k = 1; k < Z; ++k) pre[k] = E::mul(pre[k - 1], m[k - 1][dig[k - 1]]);
    // j = Z - 1: empty suffix
    raw[Z - 1][dig[Z - 1]] = E::add(raw[Z - 1][dig[Z - 1]], pre[Z - 1]);
    if constexpr (Z >= 2) {
      T suf = m[Z - 1][dig[Z - 1]];
#pragma unroll
      for (int j = Z - 2; j >= 0; --j) {
        raw[j][dig[j]] = E::fma(pre[j], suf, raw[j][dig[j]]);
        if (j > 0) suf = E::mul(suf, m[j][dig[j]]);
      }
    }
  }
}

// epilogue of one out-message: sum-normalise (beliefpropagation.jl:248-253), residual term against the previous message
// on that edge (:261-267).  Returns the order-preserving key of the residual.
template <typename T, int N>
__device__ __forceinline__ unsigned long long finish_message(const T* raw, const T* old_m, int normalize, T* v) {
  using E = Elem<T>;
  T s = E::zero();
#pragma unroll
  for (int b = 0; b < N; ++b) s = E::add(s, raw[b]);
  const bool scale = normalize && !E::is_zero(s);
  T dot = E::zero();
  double n_old = 0.0, n_new = 0.0;
#pragma unroll
  for (int b = 0; b < N; ++b) {
    v[b] = scale ? E::div(raw[b], s) : raw[b];
    dot = E::fma(E::conj(old_m[b]), v[b], dot);
    n_old += E::abs2(old_m[b]);
    n_new += E::abs2(v[b]);
  }
  return residual_key(1.0 - E::abs2(dot) / (n_old * n_new));
}

// 16 bytes global -> shared without a register round trip (LDGSTS, L2 only)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// a shared-memory row of 16-byte units <-> COUNT elements in registers (COUNT * sizeof(T) is a multiple of 16)
template <typename T, int COUNT>
__device__ __forceinline__ void row_load(const double2* row, T* dst) {
  if constexpr (sizeof(T) == 16) {
#pragma unroll
    for (int e = 0; e < COUNT; ++e) {
      const double2 x = row[e];
      dst[e] = make_c64(x.x, x.y);
    }
  } else {
#pragma unroll
    for (int u = 0; u < COUNT / 2; ++u) {
      const double2 x = row[u];
      reinterpret_cast<double*>(dst)[2 * u] = x.x;
      reinterpret_cast<double*>(dst)[2 * u + 1] = x.y;
    }
  }
}
template <typename T, int COUNT>
__device__ __forceinline__ void row_store(double2* row, const T* src) {
  if constexpr (sizeof(T) == 16) {
#pragma unroll
    for (int e = 0; e < COUNT; ++e) row[e] = make_double2(src[e].re, src[e].im);
  } else {
#pragma unroll
    for (int u = 0; u < COUNT / 2; ++u)
      row[u] = make_double2(reinterpret_cast<const double*>(src)[2 * u], reinterpret_cast<const double*>(src)[2 * u + 1]);
  }
}

template <typename T, int Z, int N>
__global__ void __launch_bounds__(NT, Shape<T, Z, N>::MIN_CTAS) bp_update_single_vertex(Args a) {
  using S = Shape<T, Z, N>;
  constexpr int NE = S::NE;
  extern __shared__ __align__(16) unsigned char vx_smem[];
  const T* __restrict__ sites = reinterpret_cast<const T*>(a.sites);
  const T* __restrict__ msg_in = reinterpret_cast<const T*>(a.msg_in);
  T* __restrict__ msg_out = reinterpret_cast<T*>(a.msg_out);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (sweep_already_converged(a.resmax, a.stop_key)) return;
  unsigned long long key = 0ull;  // 0 = nothing recorded (residual_key never returns 0)

  // one warp = 32 consecutive vertices of the bucket per iteration (all lanes stay in the loop: warp collectives inside).
  // The grid is persistent (resident CTAs only), so a warp runs many iterations and the descriptors of iteration k + 1
  // are fetched while iteration k computes: one exposed memory latency per iteration instead of two.
  const int64_t stride = (int64_t)gridDim.x * NT;
  int64_t w0 = ((int64_t)blockIdx.x * (NT / 32) + warp) * 32;
  int64_t nx_site = 0;
  int32_t nx_in[Z], nx_out0 = 0;
#pragma unroll
  for (int k = 0; k < Z; ++k) nx_in[k] = 0;
  if (w0 + lane < a.n) {
    nx_site = a.site[w0 + lane];
#pragma unroll
    for (int k = 0; k < Z; ++k) nx_in[k] = a.moff[(int64_t)k * a.n + w0 + lane];
    nx_out0 = a.moff[(int64_t)Z * a.n + w0 + lane];
  }
  for (; w0 < a.n; w0 += stride) {
    const int64_t i = w0 + lane;
    const bool valid = i < a.n;
    const int64_t site = nx_site;
    int32_t in_off[Z], out_off[Z];
#pragma unroll
    for (int k = 0; k < Z; ++k) in_off[k] = nx_in[k];
    out_off[0] = nx_out0;
    if (i + stride < a.n) {  // prefetch the next iteration's descriptors (consumed after this iteration's stores)
      nx_site = a.site[i + stride];
#pragma unroll
      for (int k = 0; k < Z; ++k) nx_in[k] = a.moff[(int64_t)k * a.n + i + stride];
      nx_out0 = a.moff[(int64_t)Z * a.n + i + stride];
    }
    bool staged = false;
    int64_t site0 = 0;
    int32_t out0 = 0;
    if constexpr (S::STAGED) {
      site0 = __shfl_sync(0xffffffffu, site, 0);
      out0 = __shfl_sync(0xffffffffu, out_off[0], 0);
      const bool mine = valid && site == site0 + (int64_t)lane * NE && out_off[0] == out0 + lane * (Z * N);
      staged = a.out_contig && __all_sync(0xffffffffu, mine) && ((reinterpret_cast<uintptr_t>(sites + site0) & 15) == 0) &&
               ((reinterpret_cast<uintptr_t>(msg_in + out0) & 15) == 0) && ((reinterpret_cast<uintptr_t>(msg_out + out0) & 15) == 0);
    }
    if (staged) {
      if constexpr (S::STAGED) {
        double2* wt = reinterpret_cast<double2*>(vx_smem) + (size_t)warp * S::WARP_UNITS;  // 32 factor rows
        double2* wo = wt + 32 * S::RS;                                                     // 32 rows of z out-messages
        const double2* gt = reinterpret_cast<const double2*>(sites + site0);
        const double2* go = reinterpret_cast<const double2*>(msg_in + out0);
        // every global read of the iteration is issued here, back to back: the factor rows and the old out-messages go
        // straight from HBM into their transposed shared-memory slots (cp.async, 16 bytes per lane and request: each
        // request covers 512 contiguous bytes), the z gathered in-messages into registers
#pragma unroll
        for (int q = 0; q < S::RC; ++q) {
          const int c = q * 32 + lane;
          cp_async16(wt + (c / S::RC) * S::RS + (c % S::RC), gt + c);
        }
#pragma unroll
        for (int q = 0; q < S::OC; ++q) {
          const int c = q * 32 + lane;
          cp_async16(wo + (c / S::OC) * S::OS + (c % S::OC), go + c);
        }
        T m[Z][N];
#pragma unroll
        for (int k = 0; k < Z; ++k) load_run<T, N>(msg_in + in_off[k], m[k]);
        cp_async_wait_all();
        __syncwarp();
        T t[NE], old_m[Z * N];
        row_load<T, NE>(wt + lane * S::RS, t);
        row_load<T, Z * N>(wo + lane * S::OS, old_m);
        T raw[Z][N], v[Z * N];
        leave_one_out<T, Z, N>(t, m, raw);
#pragma unroll
        for (int j = 0; j < Z; ++j) {
          const unsigned long long kj = finish_message<T, N>(raw[j], old_m + j * N, a.normalize, v + j * N);
          key = kj > key ? kj : key;
        }
        row_store<T, Z * N>(wo + lane * S::OS, v);  // (only this lane touched its row since the barrier above)
        __syncwarp();
        double2* gout = reinterpret_cast<double2*>(msg_out + out0);
#pragma unroll
        for (int q = 0; q < S::OC; ++q) {
          const int c = q * 32 + lane;
          gout[c] = wo[(c / S::OC) * S::OS + (c % S::OC)];
        }
        __syncwarp();  // rows are re-filled by other lanes in the next iteration
      }
    } else if (valid) {
      // per-thread path: a tail warp, a bucket whose vertices are not consecutive in memory, or an odd row size
#pragma unroll
      for (int j = 1; j < Z; ++j) out_off[j] = a.moff[(int64_t)(Z + j) * a.n + i];
      T t[NE];
      load_run<T, NE>(sites + site, t);
      T m[Z][N], old_m[Z][N];
#pragma unroll
      for (int k = 0; k < Z; ++k) load_run<T, N>(msg_in + in_off[k], m[k]);
#pragma unroll
      for (int j = 0; j < Z; ++j) load_run<T, N>(msg_in + out_off[j], old_m[j]);
      T raw[Z][N];
      leave_one_out<T, Z, N>(t, m, raw);
#pragma unroll
      for (int j = 0; j < Z; ++j) {
        T v[N];
        const unsigned long long kj = finish_message<T, N>(raw[j], old_m[j], a.normalize, v);
        key = kj > key ? kj : key;
        store_run<T, N>(msg_out + out_off[j], v);
      }
    }
  }

  // thread -> warp -> CTA -> one atomicMax
  if (a.resmax == nullptr) return;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, s);
    key = o > key ? o : key;
  }
  __shared__ unsigned long long wkey[NT / 32];
  if (lane == 0) wkey[warp] = key;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) key = wkey[w] > key ? wkey[w] : key;
    if (key != 0ull) atomicMax(a.resmax, key);
  }
}

// `ctas_per_sm` > 0: launch a persistent grid of num_sms * min(ctas_per_sm, occupancy) CTAs (at most one CTA per NT vertices)
template <typename T, int Z, int N>
inline cudaError_t launch_zn(const Args& a, int num_sms, cudaStream_t stream) {
  if constexpr (supported<T, Z, N>()) {
    constexpr int smem = Shape<T, Z, N>::SMEM_BYTES;
    static int resident = 0;  // CTAs per SM of this instantiation (queried once)
    if (resident == 0) {
      if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(bp_update_single_vertex<T, Z, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
      }
      int occ = 0;
      const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bp_update_single_vertex<T, Z, N>, NT, smem);
      if (e != cudaSuccess) return e;
      resident = occ > 0 ? occ : 1;
      if (const char* env = getenv("BPX_VERTEX_CTAS_PER_SM")) {  // tuning knob: CTAs per SM of the launch (> occupancy: not persistent)
        const int v = atoi(env);
        if (v > 0) resident = v;
      }
    }
    const int64_t want = (a.n + NT - 1) / NT, cap = (int64_t)num_sms * resident;
    const int grid = (int)(want < cap ? want : cap);
    bp_update_single_vertex<T, Z, N><<<grid, NT, smem, stream>>>(a);
    return cudaGetLastError();
  } else {
    return cudaErrorInvalidValue;
  }
}

template <typename T, int N>
inline cudaError_t launch_n(const Args& a, int z, int num_sms, cudaStream_t stream) {
  switch (z) {
    case 1: return launch_zn<T, 1, N>(a, num_sms, stream);
    case 2: return launch_zn<T, 2, N>(a, num_sms, stream);
    case 3: return launch_zn<T, 3, N>(a, num_sms, stream);
    case 4: return launch_zn<T, 4, N>(a, num_sms, stream);
    case 5: return launch_zn<T, 5, N>(a, num_sms, stream);
    case 6: return launch_zn<T, 6, N>(a, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

template <typename T>
inline cudaError_t launch(const Args& a, int z, int n, int num_sms, cudaStream_t stream) {
  switch (n) {
    case 2: return launch_n<T, 2>(a, z, num_sms, stream);
    case 3: return launch_n<T, 3>(a, z, num_sms, stream);
    case 4: return launch_n<T, 4>(a, z, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace vertexk
}  // namespace bpx
