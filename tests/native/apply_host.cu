// TEST INFRASTRUCTURE: compiles the device code of csrc/bpx_apply.cuh for the HOST (the `Team` abstraction collapses to
// one sequential lane) so that `-m "not gpu"` tests can check every numerical stage of the gate-application kernel
// against numpy without a GPU.  Nothing in the product links or loads this file; libbpx.so runs the same code as a
// CUDA kernel only (bpx_apply_gates fails without a device like every other compute entry point).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../itensornetworksnext.jl_b200/csrc/bpx_apply3.cuh"
#include "../../itensornetworksnext.jl_b200/csrc/bpx_expect2.cuh"

using namespace bpx;
using namespace bpx::applyk;

static Team host_team() {
  Team t;
  t.lane = 0;
  t.wid = 0;
  t.nw = 1;
  return t;
}

template <typename T>
static int two_site(int z1, int d1, int slot1, const int32_t* dims1, T* site1, const T* msgs1, int z2, int d2, int slot2,
                     const int32_t* dims2, T* site2, const T* msgs2, const T* op, int max_rank, int normalize,
                     T* msg_out, double* sv_out, int64_t smem_elems = 0, int version3 = 0) {
  GateDesc g;
  memset(&g, 0, sizeof(g));
  g.nsides = 2;
  const int chi = dims1[slot1];
  g.chi_b = chi;
  std::vector<T> sites, msgs;
  const int zs[2] = {z1, z2}, ds[2] = {d1, d2}, slots[2] = {slot1, slot2};
  const int32_t* dims[2] = {dims1, dims2};
  const T* site_in[2] = {site1, site2};
  const T* msg_in[2] = {msgs1, msgs2};
  for (int a = 0; a < 2; ++a) {
    Side& s = g.s[a];
    s.z = zs[a];
    s.d = ds[a];
    s.bond_slot = slots[a];
    int64_t moff = 0;
    for (int i = 0; i < s.z; ++i) {
      s.dim[i] = dims[a][i];
      s.in_msg[i] = (int64_t)msgs.size() + moff;
      moff += (int64_t)s.dim[i] * s.dim[i];
    }
    msgs.insert(msgs.end(), msg_in[a], msg_in[a] + moff);
    finish_side(s, chi);
    s.site_off = (int64_t)sites.size();
    sites.insert(sites.end(), site_in[a], site_in[a] + s.n);
  }
  g.msg12 = (int64_t)msgs.size();
  g.msg21 = g.msg12 + (int64_t)chi * chi;
  msgs.resize(msgs.size() + 2 * (size_t)chi * chi);
  const int m = g.s[0].nref * g.s[0].d, n = g.s[1].nref * g.s[1].d;
  int k = max_rank > 0 ? max_rank : chi;
  if (k > chi) k = chi;
  if (k > m) k = m;
  if (k > n) k = n;
  g.k = k;
  int flag = 0;
  double ssum = 0.0;
  int status = 0;
  if (version3) {  // version 3 (bpx_apply3.cuh): status 1 = the gate was declined (left untouched for the fallback)
    const int64_t need = applyk3::smem_need(g, Elem<T>::is_complex);
    if (need == 0) return -2;
    const applyk3::Layout3 L3 = applyk3::layout3_of(g);
    std::vector<T> ws((size_t)L3.total + 1), smem((size_t)need);
    int bad = 0;
    status = applyk3::run_two_site_v3<T>(host_team(), g, sites.data(), msgs.data(), op, ws.data(), sv_out, normalize, &flag, &bad,
                                         smem.data());
  } else if (smem_elems > 0) {  // version 2 (bpx_apply2.cuh): `smem_elems` elements of T stand in for the CTA's shared memory
    if (applyk2::block_rows(g.s[0], smem_elems) == 0 || applyk2::block_rows(g.s[1], smem_elems) == 0) return -2;
    const applyk2::Layout2 L2 = applyk2::layout2_of(g, smem_elems);
    std::vector<T> ws((size_t)L2.total + 1), smem((size_t)smem_elems);
    applyk2::run_two_site_v2<T>(host_team(), g, sites.data(), msgs.data(), op, ws.data(), sv_out, normalize, &flag,
                                smem.data(), smem_elems);
  } else {
    const Layout L = layout_of(g);
    std::vector<T> ws((size_t)L.total + 1);
    run_gate<T>(host_team(), g, sites.data(), msgs.data(), op, ws.data(), sv_out, normalize, &flag, &ssum);
  }
  memcpy(site1, sites.data() + g.s[0].site_off, sizeof(T) * g.s[0].n);
  memcpy(site2, sites.data() + g.s[1].site_off, sizeof(T) * g.s[1].n);
  memcpy(msg_out, msgs.data() + g.msg12, sizeof(T) * chi * chi);
  return status;
}

template <typename T>
static void one_site(int z, int d, const int32_t* dims, T* site, const T* msgs_in, const T* op, int normalize) {
  GateDesc g;
  memset(&g, 0, sizeof(g));
  g.nsides = 1;
  Side& s = g.s[0];
  s.z = z;
  s.d = d;
  s.bond_slot = -1;
  int64_t moff = 0;
  for (int i = 0; i < z; ++i) {
    s.dim[i] = dims[i];
    s.in_msg[i] = moff;
    moff += (int64_t)dims[i] * dims[i];
  }
  finish_side(s, 0);
  const Layout L = layout_of(g);
  std::vector<T> ws((size_t)L.total + 1);
  std::vector<T> msgs(msgs_in, msgs_in + moff);
  int flag = 0;
  double ssum = 0.0;
  run_gate<T>(host_team(), g, site, msgs.data(), op, ws.data(), nullptr, normalize, &flag, &ssum);
}

template <typename T>
static void edge_expect(int z1, int d1, int slot1, const int32_t* dims1, const T* site1, const T* msgs1, int z2, int d2,
                        int slot2, const int32_t* dims2, const T* site2, const T* msgs2, const T* op, T* num, T* den) {
  expect2::EdgeDesc g;
  memset(&g, 0, sizeof(g));
  const int chi = dims1[slot1];
  g.chi_b = chi;
  std::vector<T> sites, msgs;
  const int zs[2] = {z1, z2}, ds[2] = {d1, d2}, slots[2] = {slot1, slot2};
  const int32_t* dims[2] = {dims1, dims2};
  const T* site_in[2] = {site1, site2};
  const T* msg_in[2] = {msgs1, msgs2};
  for (int a = 0; a < 2; ++a) {
    Side& s = g.s[a];
    s.z = zs[a];
    s.d = ds[a];
    s.bond_slot = slots[a];
    int64_t moff = 0;
    for (int i = 0; i < s.z; ++i) {
      s.dim[i] = dims[a][i];
      s.in_msg[i] = (int64_t)msgs.size() + moff;
      moff += (int64_t)s.dim[i] * s.dim[i];
    }
    msgs.insert(msgs.end(), msg_in[a], msg_in[a] + moff);
    finish_side(s, chi);
    s.site_off = (int64_t)sites.size();
    sites.insert(sites.end(), site_in[a], site_in[a] + s.n);
  }
  std::vector<T> ws((size_t)expect2::layout_of(g).total + 1);
  double accum[4];
  expect2::run_edge<T>(host_team(), g, sites.data(), msgs.data(), op, ws.data(), num, den, accum);
}

extern "C" {

// two-site expectation value in the BP environment (csrc/bpx_expect2.cuh): numerator and denominator of <O_e>
int apply_host_edge_expect(int dtype, int z1, int d1, int slot1, const int32_t* dims1, const void* site1, const void* msgs1, int z2,
                           int d2, int slot2, const int32_t* dims2, const void* site2, const void* msgs2, const void* op,
                           void* num, void* den) {
  if (dims1[slot1] != dims2[slot2]) return -1;
  if (dtype == 0)
    edge_expect<double>(z1, d1, slot1, dims1, (const double*)site1, (const double*)msgs1, z2, d2, slot2, dims2,
                        (const double*)site2, (const double*)msgs2, (const double*)op, (double*)num, (double*)den);
  else
    edge_expect<c64>(z1, d1, slot1, dims1, (const c64*)site1, (const c64*)msgs1, z2, d2, slot2, dims2, (const c64*)site2,
                     (const c64*)msgs2, (const c64*)op, (c64*)num, (c64*)den);
  return 0;
}

// dtype: 0 = Float64, 1 = ComplexF64 (interleaved).  msgsN: the z_N incoming messages of vertex N packed in slot order
// (chi_i^2 each, [bra, ket] column-major; the entry of the bond slot is ignored).  Sites are updated in place.
int apply_host_two_site(int dtype, int z1, int d1, int slot1, const int32_t* dims1, void* site1, const void* msgs1, int z2,
                        int d2, int slot2, const int32_t* dims2, void* site2, const void* msgs2, const void* op,
                        int max_rank, int normalize, void* msg_out, double* sv_out) {
  if (dims1[slot1] != dims2[slot2]) return -1;
  if (dtype == 0)
    two_site<double>(z1, d1, slot1, dims1, (double*)site1, (const double*)msgs1, z2, d2, slot2, dims2, (double*)site2,
                     (const double*)msgs2, (const double*)op, max_rank, normalize, (double*)msg_out, sv_out);
  else
    two_site<c64>(z1, d1, slot1, dims1, (c64*)site1, (const c64*)msgs1, z2, d2, slot2, dims2, (c64*)site2,
                  (const c64*)msgs2, (const c64*)op, max_rank, normalize, (c64*)msg_out, sv_out);
  return 0;
}

// version 2 of the two-site gate with `smem_elems` elements of emulated shared memory (-2: the shapes do not fit)
int apply_host_two_site_v2(int dtype, int z1, int d1, int slot1, const int32_t* dims1, void* site1, const void* msgs1, int z2,
                           int d2, int slot2, const int32_t* dims2, void* site2, const void* msgs2, const void* op,
                           int max_rank, int normalize, void* msg_out, double* sv_out, int64_t smem_elems) {
  if (dims1[slot1] != dims2[slot2] || smem_elems <= 0) return -1;
  if (dtype == 0)
    return two_site<double>(z1, d1, slot1, dims1, (double*)site1, (const double*)msgs1, z2, d2, slot2, dims2, (double*)site2,
                            (const double*)msgs2, (const double*)op, max_rank, normalize, (double*)msg_out, sv_out, smem_elems);
  return two_site<c64>(z1, d1, slot1, dims1, (c64*)site1, (const c64*)msgs1, z2, d2, slot2, dims2, (c64*)site2,
                       (const c64*)msgs2, (const c64*)op, max_rank, normalize, (c64*)msg_out, sv_out, smem_elems);
}

// version 3 (Gram path): 0 = applied, 1 = declined (nothing modified), -2 = shapes not supported
int apply_host_two_site_v3(int dtype, int z1, int d1, int slot1, const int32_t* dims1, void* site1, const void* msgs1, int z2,
                           int d2, int slot2, const int32_t* dims2, void* site2, const void* msgs2, const void* op,
                           int max_rank, int normalize, void* msg_out, double* sv_out) {
  if (dims1[slot1] != dims2[slot2]) return -1;
  if (dtype == 0)
    return two_site<double>(z1, d1, slot1, dims1, (double*)site1, (const double*)msgs1, z2, d2, slot2, dims2, (double*)site2,
                            (const double*)msgs2, (const double*)op, max_rank, normalize, (double*)msg_out, sv_out, 0, 1);
  return two_site<c64>(z1, d1, slot1, dims1, (c64*)site1, (const c64*)msgs1, z2, d2, slot2, dims2, (c64*)site2,
                       (const c64*)msgs2, (const c64*)op, max_rank, normalize, (c64*)msg_out, sv_out, 0, 1);
}

int apply_host_one_site(int dtype, int z, int d, const int32_t* dims, void* site, const void* msgs, const void* op,
                        int normalize) {
  if (dtype == 0)
    one_site<double>(z, d, dims, (double*)site, (const double*)msgs, (const double*)op, normalize);
  else
    one_site<c64>(z, d, dims, (c64*)site, (const c64*)msgs, (const c64*)op, normalize);
  return 0;
}

// building blocks, for stage-by-stage tests -------------------------------------------------------------------
// B (m x n, column-major) is orthogonalised in place; V (n x n) receives the accumulated rotations
int apply_host_jacobi(int dtype, void* B, int m, int n, void* V) {
  int flag = 0;
  if (dtype == 0)
    jacobi_cols<double>(host_team(), (double*)B, m, n, (double*)V, &flag);
  else
    jacobi_cols<c64>(host_team(), (c64*)B, m, n, (c64*)V, &flag);
  return 0;
}

// P (rows x cols) -> reflectors + R in place, tau[cols]; then Y (rows x ncols) <- Q Y
int apply_host_qr(int dtype, void* P, int64_t rows, int cols, void* tau, void* Y, int ncols) {
  if (dtype == 0) {
    const int nr = householder_qr<double>(host_team(), (double*)P, rows, cols, (double*)tau);
    if (Y) apply_q<double>(host_team(), (const double*)P, rows, nr, (const double*)tau, (double*)Y, ncols);
  } else {
    const int nr = householder_qr<c64>(host_team(), (c64*)P, rows, cols, (c64*)tau);
    if (Y) apply_q<c64>(host_team(), (const c64*)P, rows, nr, (const c64*)tau, (c64*)Y, ncols);
  }
  return 0;
}

// msg (chi x chi) -> X (chi x chi), Xinv (chi x chi), eigenvalues
int apply_host_gauge(int dtype, const void* msg, int chi, void* X, void* Xinv, double* ev) {
  int flag = 0;
  if (dtype == 0)
    gauge_from_message<double>(host_team(), (const double*)msg, chi, (double*)X, (double*)Xinv, ev, &flag);
  else
    gauge_from_message<c64>(host_team(), (const c64*)msg, chi, (c64*)X, (c64*)Xinv, ev, &flag);
  return 0;
}
}
