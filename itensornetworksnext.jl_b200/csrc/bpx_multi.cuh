// Single-process multi-GPU: ONE context that drives several devices from the calling thread.
//
// The reference is one call tree on one task (beliefpropagation.jl:69-92 -> AlgorithmsInterfaceExtensions.jl:27-32), so a
// drop-in plugin must reach every GPU from one `beliefpropagation()` call: bpx_create_multi(devices, ndev) returns a
// context whose every entry point fans out to one CHILD context per device (the same partitioned contexts the
// one-process-per-GPU mode uses: vertex partition, cut-edge messages stored straight into the owner's message set and the
// residual max-reduced through peer mailboxes INSIDE the sweep kernels, bpx_peer.cuh).  The children live in one address
// space, so their message sets and mailboxes are connected by plain pointers after cudaDeviceEnablePeerAccess -- no IPC
// handles, no process group, no collective library.  All launches of a sweep are asynchronous: the calling thread
// enqueues device 0, 1, ... in turn and the kernels gate each other on the device.
//
// Included at the end of bpx_api.cu (it uses that file's static helpers).
#pragma once

namespace bpx {
namespace multi {

static int fail(bpx_ctx* m, bpx_ctx* child, int rc) {
  if (rc && child) m->err = child->err;
  return rc;
}

template <typename F>
static int each(bpx_ctx* m, F f) {
  for (bpx_ctx* c : m->children) {
    cudaSetDevice(c->device);
    const int rc = f(c);
    if (rc) return fail(m, c, rc);
  }
  return BPX_OK;
}

static bpx_ctx* owner_of(bpx_ctx* m, int64_t v) { return m->children[m->multi_owner.empty() ? 0 : m->multi_owner[v]]; }

// balanced contiguous blocks of the vertex order; cost ~ flops of a vertex's updates (z * d * prod(dims) * max dim)
static std::vector<int32_t> default_owner(bpx_ctx* c0, int ndev) {
  const int64_t nv = c0->nv;
  std::vector<double> cost(nv);
  double total = 0.0;
  for (int64_t v = 0; v < nv; ++v) {
    const bpx::VDesc& d = c0->h_vdesc[v];
    int mx = 1;
    for (int l = 0; l < d.z; ++l) mx = std::max(mx, d.dim[l]);
    cost[v] = (double)std::max(1, d.z) * (double)d.n * mx + 1.0;
    total += cost[v];
  }
  std::vector<int32_t> owner(nv, 0);
  double acc = 0.0;
  for (int64_t v = 0; v < nv; ++v) {
    owner[v] = (int32_t)std::min<double>(ndev - 1, std::floor(acc / total * ndev));
    acc += cost[v];
  }
  return owner;
}

static int connect_direct(bpx_ctx* a, bpx_ctx* b, int rank_b) {
  bpx::Peer p;
  p.rank = rank_b;
  p.msg[0] = b->d_msg[0];
  p.msg[1] = b->d_msg[1];
  p.mailbox = b->d_mailbox;
  p.ipc = false;
  a->peers.push_back(p);
  return BPX_OK;
}

// (re-)partition the children and wire their message sets / mailboxes together
static int partition(bpx_ctx* m, const int32_t* owner_or_null) {
  const int n = (int)m->children.size();
  bpx_ctx* c0 = m->children[0];
  std::vector<int32_t> owner = owner_or_null ? std::vector<int32_t>(owner_or_null, owner_or_null + c0->nv) : default_owner(c0, n);
  for (int64_t v = 0; v < c0->nv; ++v)
    if (owner[v] < 0 || owner[v] >= n) {
      set_error(m, "bpx_set_owner: owner[%lld] = %d out of range (0..%d)", (long long)v, owner[v], n - 1);
      return BPX_ERR_INVALID;
    }
  m->multi_owner = owner;
  for (bpx_ctx* c : m->children) c->halo_in_runs.clear();
  if (n == 1) return BPX_OK;
  int rank = 0;
  for (bpx_ctx* c : m->children) {
    cudaSetDevice(c->device);
    const int rc = bpx_set_partition(c, rank++, n, owner.data());
    if (rc) return fail(m, c, rc);
  }
  for (int a = 0; a < n; ++a) {
    bpx_ctx* ca = m->children[a];
    cudaSetDevice(ca->device);
    ca->halo_in_runs.clear();
    for (int64_t e = 0; e < c0->ne; ++e)
      if (owner[c0->dst[e]] == a && owner[c0->src[e]] != a) {
        const int64_t b0 = c0->msg_off[e], b1 = c0->msg_off[e + 1];
        if (!ca->halo_in_runs.empty() && ca->halo_in_runs.back().second == b0) ca->halo_in_runs.back().second = b1;
        else ca->halo_in_runs.emplace_back(b0, b1);
      }
    for (int b = 0; b < n; ++b)
      if (b != a) connect_direct(ca, m->children[b], b);
    const int rc = bpx::halo_finalize(ca);
    if (rc) return fail(m, ca, rc);
  }
  return BPX_OK;
}

static int create(const int* devices, int ndev, bpx_ctx** out) {
  if (!out || !devices || ndev < 1 || ndev > 64) {
    set_error(nullptr, "bpx_create_multi: bad arguments (1 <= ndev <= 64)");
    return BPX_ERR_INVALID;
  }
  *out = nullptr;
  bpx_ctx* m = new (std::nothrow) bpx_ctx();
  if (!m) {
    set_error(nullptr, "bpx_create_multi: out of host memory");
    return BPX_ERR_ALLOC;
  }
  m->device = devices[0];
  for (int i = 0; i < ndev; ++i) {
    bpx_ctx* c = nullptr;
    const int rc = bpx_create(devices[i], &c);
    if (rc) {
      for (bpx_ctx* k : m->children) bpx_destroy(k);
      delete m;
      return rc;  // (the create error string is already set)
    }
    c->is_child = true;
    c->no_pad = true;  // (the parent's host-side merges use the children's own layouts)
    m->children.push_back(c);
  }
  // peer access between every pair of distinct devices (the data path of a sweep is st.global / ld.acquire.sys on peers)
  for (int a = 0; a < ndev; ++a)
    for (int b = 0; b < ndev; ++b) {
      if (devices[a] == devices[b]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
      if (!can) {
        set_error(nullptr, "bpx_create_multi: device %d cannot access device %d (no NVLink / P2P path)", devices[a], devices[b]);
        for (bpx_ctx* k : m->children) bpx_destroy(k);
        delete m;
        return BPX_ERR_UNSUPPORTED;
      }
      cudaSetDevice(devices[a]);
      const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        set_error(nullptr, "bpx_create_multi: cudaDeviceEnablePeerAccess(%d -> %d): %s", devices[a], devices[b], cudaGetErrorString(e));
        cudaGetLastError();
        for (bpx_ctx* k : m->children) bpx_destroy(k);
        delete m;
        return BPX_ERR_CUDA;
      }
      cudaGetLastError();
    }
  *out = m;
  return BPX_OK;
}

static int destroy(bpx_ctx* m) {
  // every device idle first: a child's kernels may still be storing into another child's buffers
  for (bpx_ctx* c : m->children) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
  }
  for (bpx_ctx* c : m->children) bpx_destroy(c);
  delete m;
  return BPX_OK;
}

static int set_dims(bpx_ctx* m, int dtype, int mode, const int32_t* phys_dim, const int32_t* link_dim) {
  int rc = each(m, [&](bpx_ctx* c) -> int { return bpx_set_dims(c, dtype, mode, phys_dim, link_dim); });
  if (rc) return rc;
  bpx_ctx* c0 = m->children[0];
  m->dims_set = true;
  m->dtype = c0->dtype;
  m->mode = c0->mode;
  m->esize = c0->esize;
  m->nv = c0->nv;
  m->ne = c0->ne;
  m->n_und = c0->n_und;
  return partition(m, nullptr);
}

static int get_messages(bpx_ctx* m, void* packed) {
  // every message from the child that owns its source vertex (the only one that computes it)
  int rc = each(m, [&](bpx_ctx* c) -> int { return halo_gate(c); });
  if (rc) return rc;
  rc = each(m, [&](bpx_ctx* c) -> int {
    if (m->children.size() == 1) return bpx_get_messages(c, packed);
    for (auto& r : c->owned_runs) {
      const size_t o = (size_t)r.first * c->esize, len = (size_t)(r.second - r.first) * c->esize;
      BPX_CUDA(c, cudaMemcpyAsync((char*)packed + o, (const char*)c->d_msg[c->cur] + o, len, cudaMemcpyDeviceToHost, c->stream));
    }
    return BPX_OK;
  });
  if (rc) return rc;
  return each(m, [&](bpx_ctx* c) -> int {
    BPX_CUDA(c, cudaStreamSynchronize(c->stream));
    return BPX_OK;
  });
}

static int check_halo_error(bpx_ctx* m, const char* who) {
  return each(m, [&](bpx_ctx* c) -> int {
    if (!c->d_halo_error) return BPX_OK;
    int flag = 0;
    BPX_CUDA(c, cudaMemcpy(&flag, c->d_halo_error, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) {
      set_error(c, "%s: device %d timed out waiting for a peer device's sweep post", who, c->device);
      return BPX_ERR_CUDA;
    }
    return BPX_OK;
  });
}

static int sweep(bpx_ctx* m, int max_sweeps, double tol, int normalize, double* residual_out, int* sweeps_done) {
  if (m->children.size() == 1) return fail(m, m->children[0], bpx_sweep(m->children[0], max_sweeps, tol, normalize, residual_out, sweeps_done));
  REQUIRE(m, max_sweeps >= 0, "bpx_sweep: max_sweeps < 0");
  int done = 0, rc;
  double res = INFINITY;
  if ((rc = each(m, [&](bpx_ctx* c) -> int { int r = halo_gate(c); return r ? r : residual_ring_clear(c); }))) return rc;
  bpx_ctx* c0 = m->children[0];
  for (int it = 0; it < max_sweeps; ++it) {
    // one launch (group) per device, enqueued from this thread; the kernels wait for each other's posts on the device
    if ((rc = each(m, [&](bpx_ctx* c) -> int { return sweep_once(c, normalize); }))) return rc;
    ++done;
    if (tol > 0.0) {
      // the GLOBAL residual: device 0's gate folds every device's maximum (one NVLink hop); a single host sync per sweep
      cudaSetDevice(c0->device);
      if ((rc = halo_gate(c0)) || (rc = residual_read(c0, c0->history_len - 1, &res))) return fail(m, c0, rc);
      if (res < tol) break;
    }
  }
  if ((rc = each(m, [&](bpx_ctx* c) -> int { return halo_gate(c); }))) return rc;
  if (done > 0 && !(tol > 0.0)) {
    cudaSetDevice(c0->device);
    if ((rc = residual_read(c0, c0->history_len - 1, &res))) return fail(m, c0, rc);
  }
  if ((rc = each(m, [&](bpx_ctx* c) -> int {
         BPX_CUDA(c, cudaStreamSynchronize(c->stream));
         return BPX_OK;
       })))
    return rc;
  if ((rc = check_halo_error(m, "bpx_sweep"))) return rc;
  if (residual_out) *residual_out = res;
  if (sweeps_done) *sweeps_done = done;
  return BPX_OK;
}

static int sweep_host(bpx_ctx* m, const void* packed_in, void* packed_out, int normalize, double* residual_out) {
  if (m->children.size() == 1) return fail(m, m->children[0], bpx_sweep_host(m->children[0], packed_in, packed_out, normalize, residual_out));
  // every device moves the messages it owns: uploads, sweep and downloads are enqueued on all devices before the first
  // wait (a device's gate needs the posts of the others' sweeps)
  int rc = each(m, [&](bpx_ctx* c) -> int { return sweep_host_staged_enqueue(c, packed_in, packed_out, normalize); });
  if (rc) return rc;
  double res = INFINITY;
  rc = each(m, [&](bpx_ctx* c) -> int {
    double r = INFINITY;
    const int k = sweep_host_staged_finish(c, &r);
    if (c == m->children[0]) res = r;
    return k;
  });
  if (rc) return rc;
  if ((rc = check_halo_error(m, "bpx_sweep_host"))) return rc;
  if (residual_out) *residual_out = res;
  return BPX_OK;
}

// per-vertex results: every child fills in the vertices it owns (0 elsewhere)
template <typename F>
static int merge_vertex(bpx_ctx* m, void* out, F call) {
  const size_t es = (size_t)m->esize;
  std::vector<char> tmp((size_t)std::max<int64_t>(1, m->nv) * es);
  int rank = 0;
  for (bpx_ctx* c : m->children) {
    cudaSetDevice(c->device);
    const int rc = call(c, (void*)tmp.data());
    if (rc) return fail(m, c, rc);
    for (int64_t v = 0; v < m->nv; ++v)
      if (m->children.size() == 1 || m->multi_owner[v] == rank) memcpy((char*)out + v * es, tmp.data() + v * es, es);
    ++rank;
  }
  return BPX_OK;
}

static int edge_scalars(bpx_ctx* m, void* out) {
  // both messages of an undirected edge are valid on the device that owns the source of its first orientation
  const size_t es = (size_t)m->esize;
  bpx_ctx* c0 = m->children[0];
  std::vector<char> tmp((size_t)std::max<int64_t>(1, m->n_und) * es);
  int rank = 0;
  for (bpx_ctx* c : m->children) {
    cudaSetDevice(c->device);
    const int rc = bpx_edge_scalars(c, tmp.data());
    if (rc) return fail(m, c, rc);
    int64_t i = 0;
    for (int64_t e = 0; e < c0->ne; ++e) {
      if (e >= c0->rev[e]) continue;
      if (m->children.size() == 1 || m->multi_owner[c0->src[e]] == rank) memcpy((char*)out + i * es, tmp.data() + i * es, es);
      ++i;
    }
    ++rank;
  }
  return BPX_OK;
}

static int counters(bpx_ctx* m, int64_t out[3], int reset) {
  int64_t acc[3] = {0, 0, 0};
  for (bpx_ctx* c : m->children) {
    int64_t t[3];
    bpx_counters(c, t, reset);
    acc[0] += t[0];
    acc[1] += t[1];
    if (c == m->children[0]) acc[2] = t[2];
  }
  if (out) memcpy(out, acc, sizeof(acc));
  return BPX_OK;
}

static int set_stream(bpx_ctx* m, void* cuda_stream) {
  if (m->children.size() == 1) return fail(m, m->children[0], bpx_set_stream(m->children[0], cuda_stream));
  REQUIRE(m, cuda_stream == nullptr, "bpx_set_stream: a multi-device context runs one internal stream per device");
  return BPX_OK;
}

// gates / two-site expectation values: every item goes to the device that owns ALL its vertices; an item across a cut edge
// is refused like on per-rank contexts (include/bpx.h)
static int apply_two(bpx_ctx* m, int64_t n, const int64_t* edges, const void* ops, int max_rank, int normalize, double* sv_out) {
  bpx_ctx* c0 = m->children[0];
  if (m->children.size() == 1) return fail(m, c0, bpx_apply_two_site_gates(c0, n, edges, ops, max_rank, normalize, sv_out));
  REQUIRE(m, n >= 0 && (n == 0 || (edges && ops)), "bpx_apply_two_site_gates: bad arguments");
  const size_t es = (size_t)m->esize;
  const size_t nc = m->children.size();
  std::vector<std::vector<int64_t>> ed(nc);
  std::vector<std::vector<char>> op(nc);
  std::vector<std::vector<int64_t>> sv_at(nc);  // where each gate's singular values go in the caller's array
  int64_t sv_off = 0;
  size_t op_off = 0;
  for (int64_t g = 0; g < n; ++g) {
    const int64_t e = edges[g];
    REQUIRE(m, e >= 0 && e < c0->ne, "bpx_apply_two_site_gates: edges[%lld] out of range", (long long)g);
    const int32_t u = c0->src[e], v = c0->dst[e];
    if (m->multi_owner[u] != m->multi_owner[v]) {
      set_error(m, "bpx_apply_two_site_gates: gate %lld acts across a cut edge (devices %d and %d)", (long long)g, m->multi_owner[u], m->multi_owner[v]);
      return BPX_ERR_UNSUPPORTED;
    }
    const int k = m->multi_owner[u];
    const size_t d1 = (size_t)c0->phys_dim[u], d2 = (size_t)c0->phys_dim[v], nel = d1 * d2 * d1 * d2;
    ed[k].push_back(e);
    op[k].insert(op[k].end(), (const char*)ops + op_off * es, (const char*)ops + (op_off + nel) * es);
    sv_at[k].push_back(sv_off);
    op_off += nel;
    sv_off += c0->link_dim[e];
  }
  for (size_t k = 0; k < nc; ++k) {
    if (ed[k].empty()) continue;
    bpx_ctx* c = m->children[k];
    cudaSetDevice(c->device);
    int64_t tot = 0;
    for (int64_t e : ed[k]) tot += c0->link_dim[e];
    std::vector<double> sv((size_t)std::max<int64_t>(1, tot));
    const int rc = bpx_apply_two_site_gates(c, (int64_t)ed[k].size(), ed[k].data(), op[k].data(), max_rank, normalize, sv_out ? sv.data() : nullptr);
    if (rc) return fail(m, c, rc);
    if (sv_out) {
      int64_t o = 0;
      for (size_t i = 0; i < ed[k].size(); ++i) {
        const int64_t len = c0->link_dim[ed[k][i]];
        memcpy(sv_out + sv_at[k][i], sv.data() + o, (size_t)len * sizeof(double));
        o += len;
      }
    }
  }
  return BPX_OK;
}

static int apply_one(bpx_ctx* m, int64_t n, const int64_t* vertices, const void* ops, int normalize) {
  bpx_ctx* c0 = m->children[0];
  if (m->children.size() == 1) return fail(m, c0, bpx_apply_one_site_gates(c0, n, vertices, ops, normalize));
  REQUIRE(m, n >= 0 && (n == 0 || (vertices && ops)), "bpx_apply_one_site_gates: bad arguments");
  const size_t es = (size_t)m->esize, nc = m->children.size();
  std::vector<std::vector<int64_t>> vs(nc);
  std::vector<std::vector<char>> op(nc);
  size_t op_off = 0;
  for (int64_t g = 0; g < n; ++g) {
    const int64_t v = vertices[g];
    REQUIRE(m, v >= 0 && v < c0->nv, "bpx_apply_one_site_gates: vertices[%lld] out of range", (long long)g);
    const int k = m->multi_owner[v];
    const size_t nel = (size_t)c0->phys_dim[v] * c0->phys_dim[v];
    vs[k].push_back(v);
    op[k].insert(op[k].end(), (const char*)ops + op_off * es, (const char*)ops + (op_off + nel) * es);
    op_off += nel;
  }
  for (size_t k = 0; k < nc; ++k) {
    if (vs[k].empty()) continue;
    bpx_ctx* c = m->children[k];
    cudaSetDevice(c->device);
    const int rc = bpx_apply_one_site_gates(c, (int64_t)vs[k].size(), vs[k].data(), op[k].data(), normalize);
    if (rc) return fail(m, c, rc);
  }
  return BPX_OK;
}

static int edge_expect(bpx_ctx* m, int64_t n, const int64_t* edges, const void* ops, void* num_out, void* den_out) {
  bpx_ctx* c0 = m->children[0];
  if (m->children.size() == 1) return fail(m, c0, bpx_edge_expect(c0, n, edges, ops, num_out, den_out));
  REQUIRE(m, n >= 0 && (n == 0 || (edges && ops && num_out && den_out)), "bpx_edge_expect: bad arguments");
  const size_t es = (size_t)m->esize, nc = m->children.size();
  std::vector<std::vector<int64_t>> ed(nc), at(nc);
  std::vector<std::vector<char>> op(nc);
  size_t op_off = 0;
  for (int64_t g = 0; g < n; ++g) {
    const int64_t e = edges[g];
    REQUIRE(m, e >= 0 && e < c0->ne, "bpx_edge_expect: edges[%lld] out of range", (long long)g);
    const int32_t u = c0->src[e], v = c0->dst[e];
    if (m->multi_owner[u] != m->multi_owner[v]) {
      set_error(m, "bpx_edge_expect: edge %lld is a cut edge (devices %d and %d)", (long long)g, m->multi_owner[u], m->multi_owner[v]);
      return BPX_ERR_UNSUPPORTED;
    }
    const int k = m->multi_owner[u];
    const size_t d1 = (size_t)c0->phys_dim[u], d2 = (size_t)c0->phys_dim[v], nel = d1 * d2 * d1 * d2;
    ed[k].push_back(e);
    at[k].push_back(g);
    op[k].insert(op[k].end(), (const char*)ops + op_off * es, (const char*)ops + (op_off + nel) * es);
    op_off += nel;
  }
  for (size_t k = 0; k < nc; ++k) {
    if (ed[k].empty()) continue;
    bpx_ctx* c = m->children[k];
    cudaSetDevice(c->device);
    std::vector<char> num(ed[k].size() * es), den(ed[k].size() * es);
    const int rc = bpx_edge_expect(c, (int64_t)ed[k].size(), ed[k].data(), op[k].data(), num.data(), den.data());
    if (rc) return fail(m, c, rc);
    for (size_t i = 0; i < ed[k].size(); ++i) {
      memcpy((char*)num_out + (size_t)at[k][i] * es, num.data() + i * es, es);
      memcpy((char*)den_out + (size_t)at[k][i] * es, den.data() + i * es, es);
    }
  }
  return BPX_OK;
}

}  // namespace multi
}  // namespace bpx

extern "C" int bpx_create_multi(const int* devices, int ndev, bpx_ctx** out) { return bpx::multi::create(devices, ndev, out); }

extern "C" int bpx_num_devices(const bpx_ctx* ctx) {
  if (ctx && ctx->pad_active) ctx = ctx->children[0];  // zero-padded problem: the child is the (multi-device) context
  return !ctx ? -1 : (ctx->children.empty() ? 1 : (int)ctx->children.size());
}

extern "C" int bpx_set_owner(bpx_ctx* ctx, const int32_t* owner) {
  if (!ctx) return BPX_ERR_INVALID;
  if (ctx->pad_active) ctx = ctx->children[0];
  REQUIRE(ctx, !ctx->children.empty(), "bpx_set_owner: not a multi-device context (use bpx_set_partition on per-rank contexts)");
  REQUIRE(ctx, ctx->dims_set, "bpx_set_owner: call bpx_set_dims first");
  return bpx::multi::partition(ctx, owner);
}

extern "C" int bpx_get_owner(const bpx_ctx* ctx, int32_t* owner_out) {
  if (!ctx || !owner_out) return BPX_ERR_INVALID;
  if (ctx->pad_active) ctx = ctx->children[0];
  if (ctx->children.empty()) {
    for (int64_t v = 0; v < ctx->nv; ++v) owner_out[v] = ctx->owner.empty() ? 0 : ctx->owner[v];
    return BPX_OK;
  }
  for (int64_t v = 0; v < ctx->nv; ++v) owner_out[v] = ctx->multi_owner.empty() ? 0 : ctx->multi_owner[v];
  return BPX_OK;
}
