/*
 * CPU ORACLE (plain C) for the belief-propagation message-update path.  TEST INFRASTRUCTURE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Independent restatement (plain loops, no BLAS) of
 *   /root/reference/src/beliefpropagation/beliefpropagation.jl:242-257   per-edge update + sum-normalise
 *   /root/reference/src/beliefpropagation/messagecache.jl:124-131        incoming = in-edges minus reverse
 *   /root/reference/src/normnetwork.jl:49-54                             factor = ket (x) conj(bra)
 *   /root/reference/src/beliefpropagation/beliefpropagation.jl:261-267   residual
 * used to cross-check oracle/bp_oracle.py and as the multi-threaded "reference CPU path (restated)"
 * baseline (OpenMP over edges; the reference itself is single-threaded Julia).
 *
 * Pinning: see the header of oracle/bp_oracle.py (known-answer tests of the reference; the synchronous
 * schedule and per-sweep loopy values are "parity unpinned").
 *
 * Layout = include/bpx.h: column-major, A[s, l_0..l_{z-1}], M[bra, ket], complex interleaved.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXZ 12

#define DEFINE_ORACLE(T, SFX, CONJ)                                                                          \
  /* nxt[l, a', r] = sum_a M[a', a] cur[l, a, r] */                                                          \
  static void absorb_##SFX(const T* cur, T* nxt, int64_t L, int chi, int64_t R, const T* M) {                \
    for (int64_t r = 0; r < R; ++r)                                                                          \
      for (int ap = 0; ap < chi; ++ap) {                                                                     \
        T* o = nxt + L * (ap + (int64_t)chi * r);                                                            \
        for (int64_t l = 0; l < L; ++l) o[l] = 0;                                                            \
        for (int a = 0; a < chi; ++a) {                                                                      \
          const T m = M[ap + (int64_t)chi * a];                                                              \
          const T* c = cur + L * (a + (int64_t)chi * r);                                                     \
          for (int64_t l = 0; l < L; ++l) o[l] += m * c[l];                                                  \
        }                                                                                                    \
      }                                                                                                      \
  }                                                                                                          \
  /* absorption order: z-1 absorbs, then close with conj(A); out[b', b] unnormalised */                      \
  static void contract_##SFX(int z, int d, const int32_t* dim, int slot, const T* A, const T* const* in,     \
                             T* out, T* work0, T* work1) {                                                   \
    int64_t n = d;                                                                                           \
    for (int i = 0; i < z; ++i) n *= dim[i];                                                                 \
    const T* cur = A;                                                                                        \
    T* bufs[2] = {work0, work1};                                                                             \
    int which = 0;                                                                                           \
    int64_t L = d;                                                                                           \
    for (int i = 0; i < z; ++i) {                                                                            \
      if (i != slot) {                                                                                       \
        absorb_##SFX(cur, bufs[which], L, dim[i], n / (L * dim[i]), in[i]);                                  \
        cur = bufs[which];                                                                                   \
        which ^= 1;                                                                                          \
      }                                                                                                      \
      L *= dim[i];                                                                                           \
    }                                                                                                        \
    const int chi = dim[slot];                                                                               \
    L = d;                                                                                                   \
    for (int i = 0; i < slot; ++i) L *= dim[i];                                                              \
    const int64_t R = n / (L * chi);                                                                         \
    for (int b = 0; b < chi; ++b)                                                                            \
      for (int bp = 0; bp < chi; ++bp) {                                                                     \
        T acc = 0;                                                                                           \
        for (int64_t r = 0; r < R; ++r) {                                                                    \
          const T* t = cur + L * (b + (int64_t)chi * r);                                                     \
          const T* a = A + L * (bp + (int64_t)chi * r);                                                      \
          for (int64_t l = 0; l < L; ++l) acc += t[l] * CONJ(a[l]);                                          \
        }                                                                                                    \
        out[bp + (int64_t)chi * b] = acc;                                                                    \
      }                                                                                                      \
  }                                                                                                          \
  /* the reference's literal evaluation: materialise the chi^(2z) double-layer factor first */   \
  static int contract_literal_##SFX(int z, int d, const int32_t* dim, int slot, const T* A,                  \
                                    const T* const* in, T* out) {                                            \
    int64_t nk = 1;                                                                                          \
    for (int i = 0; i < z; ++i) nk *= dim[i];                                                                \
    if ((double)nk * (double)nk > 3.0e8) return -1;                                                          \
    T* D = (T*)malloc(sizeof(T) * (size_t)(nk * nk)); /* D[ket multi-index, bra multi-index] */              \
    if (!D) return -2;                                                                                       \
    for (int64_t kb = 0; kb < nk; ++kb)                                                                      \
      for (int64_t kk = 0; kk < nk; ++kk) {                                                                  \
        T acc = 0;                                                                                           \
        for (int s = 0; s < d; ++s) acc += A[s + d * kk] * CONJ(A[s + d * kb]);                              \
        D[kk + nk * kb] = acc;                                                                               \
      }                                                                                                      \
    const int chi = dim[slot];                                                                               \
    for (int i = 0; i < chi * chi; ++i) out[i] = 0;                                                          \
    int idxk[MAXZ], idxb[MAXZ];                                                                              \
    for (int64_t kb = 0; kb < nk; ++kb) {                                                                    \
      int64_t t = kb;                                                                                        \
      for (int i = 0; i < z; ++i) { idxb[i] = (int)(t % dim[i]); t /= dim[i]; }                              \
      for (int64_t kk = 0; kk < nk; ++kk) {                                                                  \
        t = kk;                                                                                              \
        for (int i = 0; i < z; ++i) { idxk[i] = (int)(t % dim[i]); t /= dim[i]; }                            \
        T w = D[kk + nk * kb];                                                                               \
        for (int i = 0; i < z; ++i)                                                                          \
          if (i != slot) w *= in[i][idxb[i] + (int64_t)dim[i] * idxk[i]];                                    \
        out[idxb[slot] + (int64_t)chi * idxk[slot]] += w;                                                    \
      }                                                                                                      \
    }                                                                                                        \
    free(D);                                                                                                 \
    return 0;                                                                                                \
  }                                                                                                          \
  static void normalize_##SFX(T* m, int n) {                                                                 \
    T s = 0;                                                                                                 \
    for (int i = 0; i < n; ++i) s += m[i];                                                                   \
    if (s != 0)                                                                                              \
      for (int i = 0; i < n; ++i) m[i] /= s;                                                                 \
  }

#define CONJ_REAL(x) (x)
DEFINE_ORACLE(double, f64, CONJ_REAL)
DEFINE_ORACLE(double _Complex, c64, conj)

/* One update of directed edge e from `msgs_in` into out (chi^2 elements). variant: 0 absorption order, 1 literal. */
static int update_edge(int is_complex, int64_t e, const int64_t* src, const int64_t* rev, const int64_t* slot,
                       const int64_t* row_ptr, const int32_t* phys_dim, const int32_t* link_dim, const int64_t* site_off,
                       const int64_t* msg_off, const void* sites, const void* msgs_in, void* out, int normalize, int variant,
                       void* work0, void* work1) {
  const int64_t u = src[e];
  const int z = (int)(row_ptr[u + 1] - row_ptr[u]);
  int32_t dim[MAXZ] = {0};
  const void* in[MAXZ];
  const size_t w = is_complex ? 16 : 8;
  for (int i = 0; i < z; ++i) {
    const int64_t f = row_ptr[u] + i; /* out-edge on leg i; its reverse carries the incoming message */
    dim[i] = link_dim[f];
    in[i] = (const char*)msgs_in + w * (size_t)msg_off[rev[f]];
  }
  const int s = (int)slot[e];
  const int chi = dim[s];
  int rc = 0;
  if (!is_complex) {
    const double* A = (const double*)sites + site_off[u];
    if (variant == 0)
      contract_f64(z, phys_dim[u], dim, s, A, (const double* const*)in, (double*)out, (double*)work0, (double*)work1);
    else
      rc = contract_literal_f64(z, phys_dim[u], dim, s, A, (const double* const*)in, (double*)out);
    if (!rc && normalize) normalize_f64((double*)out, chi * chi);
  } else {
    const double _Complex* A = (const double _Complex*)sites + site_off[u];
    if (variant == 0)
      contract_c64(z, phys_dim[u], dim, s, A, (const double _Complex* const*)in, (double _Complex*)out, (double _Complex*)work0,
                   (double _Complex*)work1);
    else
      rc = contract_literal_c64(z, phys_dim[u], dim, s, A, (const double _Complex* const*)in, (double _Complex*)out);
    if (!rc && normalize) normalize_c64((double _Complex*)out, chi * chi);
  }
  return rc;
}

/* Synchronous sweep over `n_list` edges (edge_list NULL: all ne edges); msgs_out must not alias msgs_in.
 * Edges not in the list are copied through.  nthreads <= 0: all cores. */
int oracle_sweep_jacobi(int is_complex, int64_t nv, int64_t ne, const int64_t* src, const int64_t* rev, const int64_t* slot,
                        const int64_t* row_ptr, const int32_t* phys_dim, const int32_t* link_dim, const int64_t* site_off,
                        const int64_t* msg_off, const void* sites, const void* msgs_in, void* msgs_out, const int64_t* edge_list,
                        int64_t n_list, int normalize, int variant, int nthreads) {
  const size_t w = is_complex ? 16 : 8;
  int64_t max_n = 1;
  for (int64_t v = 0; v < nv; ++v)
    if (site_off[v + 1] - site_off[v] > max_n) max_n = site_off[v + 1] - site_off[v];
  if (edge_list) memcpy(msgs_out, msgs_in, w * (size_t)msg_off[ne]);
  const int64_t n = edge_list ? n_list : ne;
  int err = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    void* w0 = malloc(w * (size_t)max_n);
    void* w1 = malloc(w * (size_t)max_n);
#pragma omp for schedule(dynamic, 1)
    for (int64_t i = 0; i < n; ++i) {
      const int64_t e = edge_list ? edge_list[i] : i;
      int rc = update_edge(is_complex, e, src, rev, slot, row_ptr, phys_dim, link_dim, site_off, msg_off, sites, msgs_in,
                           (char*)msgs_out + w * (size_t)msg_off[e], normalize, variant, w0, w1);
      if (rc) {
#pragma omp atomic write
        err = rc;
      }
    }
    free(w0);
    free(w1);
  }
  return err;
}

/* Reference schedule: in place along edge_seq (single thread, like the reference). */
int oracle_sweep_sequential(int is_complex, int64_t nv, int64_t ne, const int64_t* src, const int64_t* rev, const int64_t* slot,
                            const int64_t* row_ptr, const int32_t* phys_dim, const int32_t* link_dim, const int64_t* site_off,
                            const int64_t* msg_off, const void* sites, void* msgs, const int64_t* edge_seq, int64_t n_seq,
                            int normalize, int variant) {
  const size_t w = is_complex ? 16 : 8;
  int64_t max_n = 1, max_m = 1;
  for (int64_t v = 0; v < nv; ++v)
    if (site_off[v + 1] - site_off[v] > max_n) max_n = site_off[v + 1] - site_off[v];
  for (int64_t e = 0; e < ne; ++e)
    if (msg_off[e + 1] - msg_off[e] > max_m) max_m = msg_off[e + 1] - msg_off[e];
  void* w0 = malloc(w * (size_t)max_n);
  void* w1 = malloc(w * (size_t)max_n);
  void* tmp = malloc(w * (size_t)max_m);
  int rc = 0;
  for (int64_t i = 0; i < n_seq && !rc; ++i) {
    const int64_t e = edge_seq[i];
    rc = update_edge(is_complex, e, src, rev, slot, row_ptr, phys_dim, link_dim, site_off, msg_off, sites, msgs, tmp, normalize,
                     variant, w0, w1);
    if (!rc) memcpy((char*)msgs + w * (size_t)msg_off[e], tmp, w * (size_t)(msg_off[e + 1] - msg_off[e]));
  }
  free(w0);
  free(w1);
  free(tmp);
  return rc;
}

/* max_e 1 - |<m1^, m2^>|^2, m^ = m / ||m||_F */
double oracle_iterate_diff(int is_complex, int64_t ne, const int64_t* msg_off, const void* m1, const void* m2) {
  double best = -INFINITY;
  for (int64_t e = 0; e < ne; ++e) {
    const int64_t n = msg_off[e + 1] - msg_off[e];
    double n1 = 0, n2 = 0, dre = 0, dim_ = 0;
    if (!is_complex) {
      const double* a = (const double*)m1 + msg_off[e];
      const double* b = (const double*)m2 + msg_off[e];
      for (int64_t i = 0; i < n; ++i) { n1 += a[i] * a[i]; n2 += b[i] * b[i]; }
      n1 = sqrt(n1); n2 = sqrt(n2);
      for (int64_t i = 0; i < n; ++i) dre += (a[i] / n1) * (b[i] / n2);
    } else {
      const double _Complex* a = (const double _Complex*)m1 + msg_off[e];
      const double _Complex* b = (const double _Complex*)m2 + msg_off[e];
      for (int64_t i = 0; i < n; ++i) {
        n1 += creal(a[i]) * creal(a[i]) + cimag(a[i]) * cimag(a[i]);
        n2 += creal(b[i]) * creal(b[i]) + cimag(b[i]) * cimag(b[i]);
      }
      n1 = sqrt(n1); n2 = sqrt(n2);
      double _Complex d = 0;
      for (int64_t i = 0; i < n; ++i) d += conj(a[i] / n1) * (b[i] / n2);
      dre = creal(d); dim_ = cimag(d);
    }
    const double r = 1.0 - (dre * dre + dim_ * dim_);
    if (r != r) return r;
    if (r > best) best = r;
  }
  return best;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
