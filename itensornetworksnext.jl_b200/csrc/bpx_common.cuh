// Shared device/host types for libbpx (B200 / sm_100a).  See include/bpx.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/bpx.h"

namespace bpx {

// ---- element types -------------------------------------------------------------------------------
struct c64 {  // Julia ComplexF64: interleaved (re, im)
  double re, im;
};

__host__ __device__ __forceinline__ c64 make_c64(double r, double i) {
  c64 z;
  z.re = r;
  z.im = i;
  return z;
}

template <typename T>
struct Elem;

template <>
struct Elem<double> {
  static constexpr bool is_complex = false;
  __host__ __device__ static __forceinline__ double zero() { return 0.0; }
  __host__ __device__ static __forceinline__ double conj(double a) { return a; }
  // acc + a*b
  __host__ __device__ static __forceinline__ double fma(double a, double b, double acc) {
    return ::fma(a, b, acc);
  }
  __host__ __device__ static __forceinline__ double add(double a, double b) { return a + b; }
  __host__ __device__ static __forceinline__ double mul(double a, double b) { return a * b; }
  __host__ __device__ static __forceinline__ bool is_zero(double a) { return a == 0.0; }
  __host__ __device__ static __forceinline__ double div(double a, double b) { return a / b; }
  __host__ __device__ static __forceinline__ double abs2(double a) { return a * a; }
  __host__ __device__ static __forceinline__ double real(double a) { return a; }
  __host__ __device__ static __forceinline__ double imag(double) { return 0.0; }
  __device__ static __forceinline__ double shfl_xor(double a, int m) { return __shfl_xor_sync(0xffffffffu, a, m); }
};

template <>
struct Elem<c64> {
  static constexpr bool is_complex = true;
  __host__ __device__ static __forceinline__ c64 zero() { return make_c64(0.0, 0.0); }
  __host__ __device__ static __forceinline__ c64 conj(c64 a) { return make_c64(a.re, -a.im); }
  __host__ __device__ static __forceinline__ c64 fma(c64 a, c64 b, c64 acc) {
    acc.re = ::fma(a.re, b.re, acc.re);
    acc.re = ::fma(-a.im, b.im, acc.re);
    acc.im = ::fma(a.re, b.im, acc.im);
    acc.im = ::fma(a.im, b.re, acc.im);
    return acc;
  }
  __host__ __device__ static __forceinline__ c64 add(c64 a, c64 b) { return make_c64(a.re + b.re, a.im + b.im); }
  __host__ __device__ static __forceinline__ c64 mul(c64 a, c64 b) {
    return make_c64(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
  }
  __host__ __device__ static __forceinline__ bool is_zero(c64 a) { return a.re == 0.0 && a.im == 0.0; }
  __host__ __device__ static __forceinline__ c64 div(c64 a, c64 b) {
    // Smith's algorithm (what Julia's complex division is built on): robust against overflow
    if (fabs(b.re) >= fabs(b.im)) {
      double r = b.im / b.re, den = b.re + b.im * r;
      return make_c64((a.re + a.im * r) / den, (a.im - a.re * r) / den);
    } else {
      double r = b.re / b.im, den = b.re * r + b.im;
      return make_c64((a.re * r + a.im) / den, (a.im * r - a.re) / den);
    }
  }
  __host__ __device__ static __forceinline__ double abs2(c64 a) { return a.re * a.re + a.im * a.im; }
  __host__ __device__ static __forceinline__ double real(c64 a) { return a.re; }
  __host__ __device__ static __forceinline__ double imag(c64 a) { return a.im; }
  __device__ static __forceinline__ c64 shfl_xor(c64 a, int m) {
    return make_c64(__shfl_xor_sync(0xffffffffu, a.re, m), __shfl_xor_sync(0xffffffffu, a.im, m));
  }
};

// ---- per-vertex descriptor (device copy of the graph walk the reference does per update) ---------
struct VDesc {
  int64_t site_off;                 // element offset of A_v in the packed site buffer
  int64_t n;                        // elements of A_v
  int32_t z;                        // degree
  int32_t d;                        // physical dimension (1 in SINGLE mode)
  int32_t owned;                    // tensor resident on this rank (site_off valid)
  int32_t pad_;
  int32_t dim[BPX_MAX_DEGREE];      // link dims, slot order
  int32_t out_edge[BPX_MAX_DEGREE]; // directed edge v -> w_i
  int32_t in_edge[BPX_MAX_DEGREE];  // directed edge w_i -> v  (the message that arrives on leg i)
};

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v = Elem<T>::add(v, Elem<T>::shfl_xor(v, m));
  return v;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// ---- per-sweep residual maximum without an extra kernel: order-preserving 64-bit keys + atomicMax ----------
// key(x) is monotone in x; NaN maps to the largest key (Julia's `maximum` propagates NaN); key 0 = "no value yet".
__host__ __device__ __forceinline__ unsigned long long residual_key(double x) {
  if (x != x) return ~0ull;
  long long b;
#ifdef __CUDA_ARCH__
  b = __double_as_longlong(x);
#else
  memcpy(&b, &x, sizeof(b));
#endif
  return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double residual_from_key(unsigned long long k) {
  if (k == 0ull) return -INFINITY;  // nothing was recorded (e.g. a graph without edges)
  if (k == ~0ull) return NAN;
  const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  double x;
#ifdef __CUDA_ARCH__
  x = __longlong_as_double((long long)b);
#else
  memcpy(&x, &b, sizeof(x));
#endif
  return x;
}
__device__ __forceinline__ void residual_record(unsigned long long* slot, double r) {
  if (slot) atomicMax(slot, residual_key(r));
}

// Device-side StopWhenConverged (AlgorithmsInterfaceExtensions.jl:84-119) for sweeps that are enqueued ahead of the host:
// `slot` is this sweep's entry of the residual-key ring, so slot[-1] is the previous sweep's final key (stream order).  If
// that residual is already below the tolerance (key < stop_key; NaN has the largest key and never stops), the sweep must
// not happen: every CTA of every launch of the sweep returns at once, and the key is handed on (slot[0] = slot[-1]) so
// that all later pre-enqueued sweeps stop as well.  stop_key == 0: no test (first sweep of a call, tol <= 0).
__device__ __forceinline__ bool sweep_already_converged(unsigned long long* slot, unsigned long long stop_key) {
  if (stop_key == 0ull || slot == nullptr) return false;
  const unsigned long long prev = *reinterpret_cast<const volatile unsigned long long*>(slot - 1);
  if (prev == 0ull || prev >= stop_key) return false;
  if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(slot) = prev;
  return true;
}

// Epilogue shared by every update kernel: sum-normalise (beliefpropagation.jl:248-253), residual term
// 1 - |<old^, new^>|^2 (beliefpropagation.jl:261-267), store.  Executed by ONE warp; `raw` holds the
// unnormalised message (nelem entries, any address space), `old_m` / `new_m` the global slots.
template <typename T>
__device__ __forceinline__ void warp_epilogue(const T* raw, const T* old_m, T* new_m, int nelem, int normalize,
                                              double* residual_slot, int lane, unsigned long long* resmax = nullptr,
                                              T* peer_m = nullptr, T* host_m = nullptr) {
  using E = Elem<T>;
  T s = E::zero();
  for (int i = lane; i < nelem; i += 32) s = E::add(s, raw[i]);
  s = warp_sum<T>(s);
  const bool scale = normalize && !E::is_zero(s);
  T dot = E::zero();
  double n_old = 0.0, n_new = 0.0;
  for (int i = lane; i < nelem; i += 32) {
    T v = raw[i];
    if (scale) v = E::div(v, s);
    T o = old_m[i];
    dot = E::fma(E::conj(o), v, dot);
    n_old += E::abs2(o);
    n_new += E::abs2(v);
    new_m[i] = v;
    if (peer_m) peer_m[i] = v;  // cut edge: also store into the owning rank's message set (NVLink peer memory)
    if (host_m) host_m[i] = v;  // streamed host I/O: the caller's mapped host buffer
  }
  if (peer_m) __threadfence_system();  // release the peer stores here, off the kernel's tail (see peer_post_when_last)
  dot = warp_sum<T>(dot);
  n_old = warp_sum_d(n_old);
  n_new = warp_sum_d(n_new);
  if (lane == 0) {
    const double r = 1.0 - E::abs2(dot) / (n_old * n_new);
    if (residual_slot) *residual_slot = r;
    residual_record(resmax, r);
  }
}

}  // namespace bpx
