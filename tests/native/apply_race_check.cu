// TEST INFRASTRUCTURE: the gate-application device code (csrc/bpx_apply.cuh, csrc/bpx_apply2.cuh, the phases of
// csrc/bpx_apply3.cuh with one lane per warp) run on the host with one
// THREAD PER WARP and a real barrier behind Team::sync(), under ThreadSanitizer.  What the single-lane host harness
// (apply_host.cu) cannot see -- a missing barrier between phases executed by different warps -- shows up here as a data
// race report or as a result that differs from the sequential run.  Schedules with several LANES per warp (threads too,
// warp barriers standing in for __syncwarp() and for the shuffle reductions) cover the intra-warp hazards as well:
// with __syncwarp() disabled TSAN reports races in householder_qr / apply_q on this very program.
//
//   nvcc -O1 -g -std=c++17 -Xcompiler -fsanitize=thread -Xcompiler -pthread -o apply_race_check apply_race_check.cu
//   ./apply_race_check        (exit 0: no race, multi-warp == sequential; TSAN reports make it exit 66)
#include <pthread.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static thread_local pthread_barrier_t* g_barrier = nullptr;
static inline void host_team_barrier() {
  if (g_barrier) pthread_barrier_wait(g_barrier);
}
#define BPX_HOST_TEAM_SYNC() host_team_barrier()
#define BPX_FLAG_SET(p) __atomic_store_n((p), 1, __ATOMIC_RELAXED)

// several LANES per warp as threads: warp-wide sums go through a per-warp scratch line between two warp barriers (the
// shuffle reduction of the device), __syncwarp() is a warp barrier
static int g_lanes = 1;
struct WarpShared {
  pthread_barrier_t bar;
  double slot[64][2];
};
static thread_local WarpShared* g_warp = nullptr;
static inline void host_syncwarp() {
  if (g_warp) pthread_barrier_wait(&g_warp->bar);
}
#define BPX_HOST_LANES g_lanes
#define BPX_HOST_WARP_SUM(team, x) host_warp_sum((team).lane, x)
#define BPX_HOST_SYNCWARP(team) host_syncwarp()
#define BPX_HOST_ATOMIC_ADD(p, v) host_atomic_add((p), (v))
static inline void host_atomic_add(double* p, double v) {
  static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  pthread_mutex_lock(&mu);
  *p += v;
  pthread_mutex_unlock(&mu);
}
static inline double host_warp_sum(int lane, double x) {
  if (!g_warp) return x;
  g_warp->slot[lane][0] = x;
  pthread_barrier_wait(&g_warp->bar);
  double s = 0.0;
  for (int l = 0; l < g_lanes; ++l) s += g_warp->slot[l][0];
  pthread_barrier_wait(&g_warp->bar);
  return s;
}
namespace bpx { struct c64; }
static inline bpx::c64 host_warp_sum(int lane, bpx::c64 x);

#include "../../itensornetworksnext.jl_b200/csrc/bpx_apply3.cuh"
#include "../../itensornetworksnext.jl_b200/csrc/bpx_expect2.cuh"

using namespace bpx;
using namespace bpx::applyk;

static inline bpx::c64 host_warp_sum(int lane, bpx::c64 x) {
  if (!g_warp) return x;
  g_warp->slot[lane][0] = x.re;
  g_warp->slot[lane][1] = x.im;
  pthread_barrier_wait(&g_warp->bar);
  c64 s = make_c64(0.0, 0.0);
  for (int l = 0; l < g_lanes; ++l) s = make_c64(s.re + g_warp->slot[l][0], s.im + g_warp->slot[l][1]);
  pthread_barrier_wait(&g_warp->bar);
  return s;
}

static double urand(uint64_t& s) {
  s = s * 6364136223846793005ull + 1442695040888963407ull;
  return (double)((s >> 11) & ((1ull << 53) - 1)) / (double)(1ull << 53) - 0.5;
}

template <typename T>
static T rnd(uint64_t& s);
template <>
double rnd<double>(uint64_t& s) { return urand(s); }
template <>
c64 rnd<c64>(uint64_t& s) { return make_c64(urand(s), urand(s)); }

template <typename T>
struct Problem {
  GateDesc g;
  std::vector<T> sites, msgs, op;
};

// two vertices of degree z with link dims chi (external) and chi_b (bond), Hermitian PSD messages
template <typename T>
static Problem<T> make_problem(int z, int chi, int chi_b, int d, uint64_t seed) {
  Problem<T> p;
  memset(&p.g, 0, sizeof(p.g));
  p.g.nsides = 2;
  p.g.chi_b = chi_b;
  uint64_t s = seed;
  for (int a = 0; a < 2; ++a) {
    Side& sd = p.g.s[a];
    sd.z = z;
    sd.d = d;
    sd.bond_slot = a == 0 ? 1 % z : 0;
    for (int i = 0; i < z; ++i) {
      sd.dim[i] = i == sd.bond_slot ? chi_b : chi;
      sd.in_msg[i] = (int64_t)p.msgs.size();
      const int c = sd.dim[i];
      std::vector<T> f((size_t)c * c);
      for (auto& x : f) x = rnd<T>(s);
      for (int r = 0; r < c; ++r) f[r + c * r] = Elem<T>::add(f[r + c * r], from_real<T>(1.5));
      for (int r = 0; r < c; ++r)       // M = F^H F
        for (int q = 0; q < c; ++q) {
          T acc = Elem<T>::zero();
          for (int k = 0; k < c; ++k) acc = Elem<T>::fma(Elem<T>::conj(f[k + c * r]), f[k + c * q], acc);
          p.msgs.push_back(acc);
        }
    }
    finish_side(sd, chi_b);
    sd.site_off = (int64_t)p.sites.size();
    for (int64_t i = 0; i < sd.n; ++i) p.sites.push_back(rnd<T>(s));
  }
  p.g.msg12 = (int64_t)p.msgs.size();
  p.g.msg21 = p.g.msg12 + (int64_t)chi_b * chi_b;
  p.msgs.resize(p.msgs.size() + 2 * (size_t)chi_b * chi_b);
  const int m = p.g.s[0].nref * d, n = p.g.s[1].nref * d;
  p.g.k = chi_b < m ? (chi_b < n ? chi_b : n) : (m < n ? m : n);
  for (int i = 0; i < d * d * d * d; ++i) p.op.push_back(rnd<T>(s));
  return p;
}

// run one gate with `nw` warps (threads); version 2 when smem_elems > 0
template <typename T>
static void run(Problem<T> p, int nw, int lanes, int64_t smem_elems, std::vector<T>& sites_out, std::vector<double>& sv) {
  g_lanes = lanes;
  std::vector<WarpShared> warps(nw);
  for (auto& w : warps) pthread_barrier_init(&w.bar, nullptr, lanes);
  sv.assign(p.g.chi_b, 0.0);
  const int64_t total = smem_elems > 0 ? applyk2::layout2_of(p.g, smem_elems).total : layout_of(p.g).total;
  std::vector<T> ws((size_t)total + 2), smem((size_t)(smem_elems > 0 ? smem_elems : 1));
  int flag = 0;
  double ssum = 0.0;
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, nw * lanes);
  auto body = [&](int t) {
    const int w = t / lanes, l = t % lanes;
    g_barrier = nw * lanes > 1 ? &bar : nullptr;
    g_warp = lanes > 1 ? &warps[w] : nullptr;
    Team tm;
    tm.lane = l;
    tm.wid = w;
    tm.nw = nw;
    if (smem_elems > 0)
      applyk2::run_two_site_v2<T>(tm, p.g, p.sites.data(), p.msgs.data(), p.op.data(), ws.data(), sv.data(), 1, &flag, smem.data(),
                                  smem_elems);
    else
      run_gate<T>(tm, p.g, p.sites.data(), p.msgs.data(), p.op.data(), ws.data(), sv.data(), 1, &flag, &ssum);
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nw * lanes; ++t) th.emplace_back(body, t);
  body(0);
  for (auto& t : th) t.join();
  pthread_barrier_destroy(&bar);
  for (auto& w : warps) pthread_barrier_destroy(&w.bar);
  g_lanes = 1;
  sites_out = p.sites;
}

// the gauge-invariant content of a gate's result: the two tensors contracted over their bond.  (The tensors themselves
// carry the phases of the singular vectors, which one-sided Jacobi fixes only up to rounding-sensitive conventions.)
template <typename T>
static std::vector<T> pair_product(const GateDesc& g, const std::vector<T>& sites) {
  const Side& a = g.s[0];
  const Side& b = g.s[1];
  const int64_t ra = a.rows * a.d, rb = b.rows * b.d;
  std::vector<T> ma((size_t)(ra * g.chi_b)), mb((size_t)(rb * g.chi_b)), out((size_t)(ra * rb));
  for (int side = 0; side < 2; ++side) {
    const Side& sd = side ? b : a;
    std::vector<T>& m = side ? mb : ma;
    const int64_t r = sd.rows * sd.d;
    for (int64_t i = 0; i < sd.n; ++i) {
      int64_t row;
      int col;
      split_index(sd, i, row, col);
      m[(size_t)((row * sd.d + col % sd.d) + r * (col / sd.d))] = sites[(size_t)(sd.site_off + i)];
    }
  }
  for (int64_t i = 0; i < ra; ++i)
    for (int64_t j = 0; j < rb; ++j) {
      T acc = Elem<T>::zero();
      for (int k = 0; k < g.chi_b; ++k) acc = Elem<T>::fma(ma[(size_t)(i + ra * k)], mb[(size_t)(j + rb * k)], acc);
      out[(size_t)(i + ra * j)] = acc;
    }
  return out;
}

template <typename T>
static int check(const char* name, int z, int chi, int chi_b, int d, int64_t rb) {
  Problem<T> p = make_problem<T>(z, chi, chi_b, d, 12345 + z * 100 + chi);
  int64_t smem = 0;
  if (rb > 0) {
    for (int a = 0; a < 2; ++a) {
      const int64_t c = p.g.s[a].cols, need = c * c + 2 * c * (c + rb), need2 = 2 * p.g.s[a].rows;
      smem = smem > need ? smem : need;
      smem = smem > need2 ? smem : need2;
    }
    if (applyk2::block_rows(p.g.s[0], smem) == 0 || applyk2::block_rows(p.g.s[1], smem) == 0) {
      printf("%s: shapes do not fit version 2\n", name);
      return 1;
    }
  }
  std::vector<T> ref, got;
  std::vector<double> sv_ref, sv_got;
  run<T>(p, 1, 1, smem, ref, sv_ref);
  const std::vector<T> ref_raw = ref;
  int bad = 0;
  const int sched[5][2] = {{2, 1}, {5, 1}, {8, 1}, {3, 4}, {2, 7}};  // (warps, lanes per warp); lanes = 1 first (raw tensors)
  for (const auto& sc : sched) {
    const int nw = sc[0], lanes = sc[1];
    run<T>(p, nw, lanes, smem, got, sv_got);
    double raw = 0.0;
    if (lanes > 1) {  // a different order of the warp-wide sums: compare what is physical
      for (size_t i = 0; i < got.size(); ++i) raw = fmax(raw, sqrt(Elem<T>::abs2(sub(ref_raw[i], got[i]))));
      got = pair_product<T>(p.g, got);
      if (ref.size() != got.size()) ref = pair_product<T>(p.g, ref);
    }
    double err = 0.0, scale = 0.0;
    for (size_t i = 0; i < ref.size(); ++i) {
      err = fmax(err, sqrt(Elem<T>::abs2(sub(ref[i], got[i]))));
      scale = fmax(scale, sqrt(Elem<T>::abs2(ref[i])));
    }
    double sverr = 0.0;
    for (size_t i = 0; i < sv_ref.size(); ++i) sverr = fmax(sverr, fabs(sv_ref[i] - sv_got[i]));
    // one lane per warp: the arithmetic is the same in every schedule, so the results must agree to rounding of the
    // (order-independent) operations -- in practice bit for bit
    // (with several lanes the order of the warp-wide sums differs from the sequential run: agreement to rounding)
    const double tol = lanes == 1 ? 1e-12 : 1e-9;
    const bool ok = err <= tol * scale && sverr <= tol;
    printf("%-34s warps=%d lanes=%d  max |diff| = %.2e (scale %.2e), sv diff %.2e, raw tensors %.1e  %s\n", name, nw, lanes, err,
           scale, sverr, raw, ok ? "ok" : "MISMATCH");
    bad += !ok;
  }
  return bad;
}

// version 3 (the Gram path, csrc/bpx_apply3.cuh): the phases of one gate on `nw` warps of ONE lane each (its host code plays
// all lane roles on one lane).  The partial sums of the Gram pass are combined per warp, so the result depends on nw in the
// last bits: agreement to rounding, and no race report.
template <typename T>
static int run3(Problem<T> p, int nw, std::vector<T>& sites_out, std::vector<double>& sv) {
  g_lanes = 1;
  sv.assign(p.g.chi_b, 0.0);
  const int64_t need = applyk3::smem_need(p.g, Elem<T>::is_complex);
  if (need == 0) return -1;
  std::vector<T> ws((size_t)applyk3::layout3_of(p.g).total + 2), smem((size_t)need);
  int flag = 0, bad = 0;
  std::vector<int> status(nw, -1);
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, nw);
  auto body = [&](int t) {
    g_barrier = nw > 1 ? &bar : nullptr;
    g_warp = nullptr;
    Team tm;
    tm.lane = 0;
    tm.wid = t;
    tm.nw = nw;
    status[t] = applyk3::run_two_site_v3<T>(tm, p.g, p.sites.data(), p.msgs.data(), p.op.data(), ws.data(), sv.data(), 1, &flag, &bad,
                                            smem.data());
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nw; ++t) th.emplace_back(body, t);
  body(0);
  for (auto& t : th) t.join();
  pthread_barrier_destroy(&bar);
  sites_out = p.sites;
  for (int t = 0; t < nw; ++t)
    if (status[t] != status[0]) return -2;  // every warp must see the same verdict
  return status[0];
}

template <typename T>
static int check3(const char* name, int z, int chi, int chi_b, int d) {
  Problem<T> p = make_problem<T>(z, chi, chi_b, d, 4242 + z * 100 + chi);
  std::vector<T> ref, got;
  std::vector<double> sv_ref, sv_got;
  const int st = run3<T>(p, 1, ref, sv_ref);
  if (st != 0) {
    printf("%-34s sequential run: status %d (the Gram path must take this gate)\n", name, st);
    return 1;
  }
  ref = pair_product<T>(p.g, ref);
  int bad = 0;
  for (int nw : {2, 5, 8}) {
    const int s2 = run3<T>(p, nw, got, sv_got);
    got = pair_product<T>(p.g, got);
    double err = 0.0, scale = 0.0, sverr = 0.0;
    for (size_t i = 0; i < ref.size(); ++i) {
      err = fmax(err, sqrt(Elem<T>::abs2(sub(ref[i], got[i]))));
      scale = fmax(scale, sqrt(Elem<T>::abs2(ref[i])));
    }
    for (size_t i = 0; i < sv_ref.size(); ++i) sverr = fmax(sverr, fabs(sv_ref[i] - sv_got[i]));
    const bool ok = s2 == 0 && err <= 1e-9 * scale && sverr <= 1e-9;
    printf("%-34s warps=%d lanes=1  status %d, max |diff| = %.2e (scale %.2e), sv diff %.2e  %s\n", name, nw, s2, err, scale, sverr,
           ok ? "ok" : "MISMATCH");
    bad += !ok;
  }
  return bad;
}

// two-site expectation kernel (csrc/bpx_expect2.cuh) under the same schedules
template <typename T>
static int check_expect(const char* name, int z, int chi, int chi_b, int d) {
  Problem<T> p = make_problem<T>(z, chi, chi_b, d, 777 + z * 10 + chi);
  expect2::EdgeDesc g;
  memset(&g, 0, sizeof(g));
  g.s[0] = p.g.s[0];
  g.s[1] = p.g.s[1];
  g.chi_b = chi_b;
  const int64_t total = expect2::layout_of(g).total;
  const int sched[6][2] = {{1, 1}, {2, 1}, {5, 1}, {8, 1}, {3, 4}, {2, 7}};
  T ref_num = Elem<T>::zero(), ref_den = Elem<T>::zero();
  int bad = 0;
  for (const auto& sc : sched) {
    const int nw = sc[0], lanes = sc[1];
    g_lanes = lanes;
    std::vector<WarpShared> warps(nw);
    for (auto& w : warps) pthread_barrier_init(&w.bar, nullptr, lanes);
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, nw * lanes);
    std::vector<T> ws((size_t)total + 2);
    double accum[4] = {0, 0, 0, 0};
    T num = Elem<T>::zero(), den = Elem<T>::zero();
    auto body = [&](int t) {
      g_barrier = nw * lanes > 1 ? &bar : nullptr;
      g_warp = lanes > 1 ? &warps[t / lanes] : nullptr;
      Team tm;
      tm.lane = t % lanes;
      tm.wid = t / lanes;
      tm.nw = nw;
      expect2::run_edge<T>(tm, g, p.sites.data(), p.msgs.data(), p.op.data(), ws.data(), &num, &den, accum);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nw * lanes; ++t) th.emplace_back(body, t);
    body(0);
    for (auto& t : th) t.join();
    pthread_barrier_destroy(&bar);
    for (auto& w : warps) pthread_barrier_destroy(&w.bar);
    g_lanes = 1;
    if (nw * lanes == 1) {
      ref_num = num;
      ref_den = den;
      continue;
    }
    const double err = sqrt(Elem<T>::abs2(sub(num, ref_num))) + sqrt(Elem<T>::abs2(sub(den, ref_den)));
    const double scale = sqrt(Elem<T>::abs2(ref_num)) + sqrt(Elem<T>::abs2(ref_den));
    const bool ok = err <= 1e-12 * scale;
    printf("%-34s warps=%d lanes=%d  |diff| = %.2e (scale %.2e)  %s\n", name, nw, lanes, err, scale, ok ? "ok" : "MISMATCH");
    bad += !ok;
  }
  return bad;
}

int main() {
  int bad = 0;
  bad += check_expect<double>("expect2 f64  z=4 chi=3 bond=4 d=2", 4, 3, 4, 2);
  bad += check_expect<c64>("expect2 c128 z=3 chi=4 bond=3 d=2", 3, 4, 3, 2);
  bad += check<double>("v1 f64  z=4 chi=3 bond=4 d=2", 4, 3, 4, 2, 0);
  bad += check<c64>("v1 c128 z=3 chi=4 bond=3 d=2", 3, 4, 3, 2, 0);
  bad += check<double>("v1 f64  z=1 (leaf pair) bond=3", 1, 3, 3, 2, 0);
  bad += check<double>("v2 f64  z=4 chi=3 bond=4 d=2 rb=5", 4, 3, 4, 2, 5);
  bad += check<double>("v2 f64  z=4 chi=4 bond=4 d=2 rb=64", 4, 4, 4, 2, 64);
  bad += check<c64>("v2 c128 z=3 chi=4 bond=3 d=2 rb=3", 3, 4, 3, 2, 3);
  bad += check<c64>("v2 c128 z=3 chi=8 bond=2 d=3 rb=16", 3, 8, 2, 3, 16);
  bad += check<double>("v2 f64  z=1 (leaf pair) bond=3 rb=1", 1, 3, 3, 2, 1);
  bad += check3<double>("v3 f64  z=4 chi=3 bond=4 d=2", 4, 3, 4, 2);
  bad += check3<c64>("v3 c128 z=3 chi=4 bond=3 d=2", 3, 4, 3, 2);
  bad += check3<double>("v3 f64  z=3 chi=5 bond=3 d=3 (odd columns)", 3, 5, 3, 3);
  bad += check3<double>("v3 f64  z=1 (leaf pair) bond=3", 1, 3, 3, 2);
  bad += check3<c64>("v3 c128 z=1 (leaf pair) bond=4", 1, 4, 4, 2);
  bad += check3<double>("v3 f64  z=2 chi=2 bond=5 (rows < cols)", 2, 2, 5, 2);
  bad += check3<c64>("v3 c128 z=4 chi=2 bond=3 d=3", 4, 2, 3, 3);
  printf(bad ? "FAILED: %d mismatches\n" : "all schedules agree\n", bad);
  return bad ? 1 : 0;
}
