// Internal zero-padding of link dimensions 9..15 to 16 (the reference's update is shape agnostic).
//
// The tensor-pipe kernel for degree-4 vertices (bpx_sliced2.cuh) is written for link dimension 16 exactly; a Float64 PEPS
// with chi = 10 or 12 would otherwise run on the generic kernel (~0.6 TFLOP/s).  When a problem has degree-4, d = 2,
// Float64 vertices whose four link dimensions all lie in 9..16 (and not all are 16), the context the caller holds becomes a
// thin PARENT: it keeps the caller's dimensions and packed layouts, and owns ONE child context in which every link of
// dimension 9..15 has dimension 16.  Site tensors and messages are embedded with zeros on the way in and sliced on the way
// out; everything else (sweeps, residuals, beliefs, gates) runs in the child unchanged -- all BP quantities of a
// zero-padded network equal those of the original one exactly (padded tensor entries are zero, so every padded message
// entry stays exactly zero, sums, inner products and norms are unchanged).  Cost: (16/chi)^4 more memory and flops for the
// padded tensors, against a 10 x faster kernel.  BPX_NO_PAD=1, a forced kernel policy and complex / single-layer problems
// do not pad.  A context made by bpx_create_multi pads the same way: its child is then a multi-device context over the same
// device list (the per-device contexts themselves never pad).
//
// Included by bpx_api.cu after bpx_multi.cuh (uses that file's static helpers and the public entry points).
#pragma once

namespace bpx {
namespace pad {

// does this problem want padding?  idim = internal link dims (per directed edge)
static bpx_ctx* lay(bpx_ctx* c) { return c->children.empty() ? c : c->children[0]; }  // who holds graph / host layouts

static bool wanted(bpx_ctx* ctx, int dtype, int mode, const int32_t* phys, const int32_t* link, std::vector<int32_t>& idim) {
  const bpx_ctx* g = lay(ctx);  // a multi-device parent keeps the graph in its children
  if (ctx->no_pad || !g->graph_set || getenv("BPX_NO_PAD") || g->kernel_policy != BPX_KERNEL_AUTO) return false;
  if (dtype != BPX_F64 || mode != BPX_MODE_NORM || g->ne == 0 || !phys || !link) return false;
  for (int64_t e = 0; e < g->ne; ++e)
    if (link[e] < 1 || link[e] != link[g->rev[e]]) return false;  // (reported by the plain path)
  idim.assign(link, link + g->ne);
  bool any = false;
  for (int64_t v = 0; v < g->nv; ++v) {
    if (g->deg[v] != 4 || phys[v] != 2) continue;
    bool in_range = true, all16 = true;
    for (int32_t e : g->out_edge[v]) {
      if (link[e] < 9 || link[e] > 16) in_range = false;
      if (link[e] != 16) all16 = false;
    }
    if (!in_range || all16) continue;
    any = true;
  }
  // pad EVERY link of dimension 9..15: the boundary vertices then have uniform dimension 16 as well (the tuned on-chip
  // kernel instead of ragged slices), and the padded problem is exactly the shape BASELINE config 5 exercises
  if (any)
    for (int64_t e = 0; e < g->ne; ++e)
      if (link[e] >= 9 && link[e] <= 15) idim[e] = 16;
  return any;
}

static int replay_graph(bpx_ctx* p, bpx_ctx* c) {
  std::vector<int64_t> s64(p->src.begin(), p->src.end()), d64(p->dst.begin(), p->dst.end());
  return bpx_set_graph(c, p->nv, p->ne, s64.data(), d64.data(), p->slot.data());
}

// back to what the caller created: a plain context, or a multi-device parent with fresh per-device children
static void teardown(bpx_ctx* p) {
  if (!p->pad_active) return;
  for (bpx_ctx* c : p->children) bpx_destroy(c);
  p->children.clear();
  p->pad_active = false;
  p->dims_set = false;
  for (int dev : p->pad_devices) {
    bpx_ctx* c = nullptr;
    if (bpx_create(dev, &c)) continue;
    c->is_child = true;
    c->no_pad = true;
    if (p->graph_set) replay_graph(p, c);
    p->children.push_back(c);
  }
  p->pad_devices.clear();
}

// the caller's dims and packed layouts live in the parent; the child (a plain context, or a multi-device context over the
// same device list) gets the padded problem
static int wrap(bpx_ctx* p, int dtype, int mode, const int32_t* phys, const int32_t* link, const std::vector<int32_t>& idim) {
  const bool was_multi = !p->children.empty();
  std::vector<int> devs;
  std::vector<bpx_ctx*> old = p->children;
  if (was_multi) {  // take the graph over from the children (identical in all of them)
    const bpx_ctx* g = old[0];
    for (bpx_ctx* c : old) devs.push_back(c->device);
    p->nv = g->nv;
    p->ne = g->ne;
    p->src = g->src;
    p->dst = g->dst;
    p->slot = g->slot;
    p->rev = g->rev;
    p->deg = g->deg;
    p->out_edge = g->out_edge;
    p->graph_set = true;
  }
  const int64_t nv = p->nv, ne = p->ne;
  bpx_ctx* c = nullptr;
  int rc = was_multi ? bpx_create_multi(devs.data(), (int)devs.size(), &c) : bpx_create(p->device, &c);
  if (rc) {
    set_error(p, "bpx_set_dims: creating the padded context failed: %s", bpx_last_error(nullptr));
    return rc;
  }
  c->no_pad = true;
  rc = replay_graph(p, c);
  if (!rc && !was_multi && p->stream != p->own_stream) rc = bpx_set_stream(c, (void*)p->stream);
  if (!rc) rc = bpx_set_dims(c, dtype, mode, phys, idim.data());
  if (rc) {
    p->err = c->err;
    bpx_destroy(c);
    return rc;
  }
  if (was_multi) {
    for (bpx_ctx* k : old) bpx_destroy(k);
    p->children.clear();
    p->multi_owner.clear();
  } else {
    free_problem(p);  // the parent holds no device memory
  }
  p->pad_devices = devs;
  p->dtype = dtype;
  p->mode = mode;
  p->esize = 8;
  p->link_dim.assign(link, link + ne);
  p->phys_dim.assign(phys, phys + nv);
  p->site_off.assign(nv + 1, 0);
  p->msg_off.assign(ne + 1, 0);
  for (int64_t v = 0; v < nv; ++v) {
    int64_t n = phys[v];
    for (int32_t e : p->out_edge[v]) n *= link[e];
    p->site_off[v + 1] = p->site_off[v] + n;
  }
  for (int64_t e = 0; e < ne; ++e) p->msg_off[e + 1] = p->msg_off[e] + (int64_t)link[e] * link[e];
  p->n_und = lay(c)->n_und;
  p->children.push_back(c);
  p->pad_active = true;
  p->dims_set = true;
  return BPX_OK;
}

// ---- embedding / slicing of column-major blocks ------------------------------------------------------------------
// small (sdim) inside big (bdim), both column-major with nd dims; to_big: big must be zero-initialised by the caller
static void copy_block(double* big, const int* bdim, double* small_, const int* sdim, int nd, bool to_big) {
  int64_t ns = 1;
  for (int k = 0; k < nd; ++k) ns *= sdim[k];
  if (ns == 0) return;
  const int run = sdim[0];
  int idx[BPX_MAX_DEGREE + 2] = {0};
  for (int64_t s0 = 0; s0 < ns; s0 += run) {
    int64_t boff = 0, stride = 1;
    for (int k = 0; k < nd; ++k) {
      boff += idx[k] * stride;
      stride *= bdim[k];
    }
    if (to_big) memcpy(big + boff, small_ + s0, (size_t)run * sizeof(double));
    else memcpy(small_ + s0, big + boff, (size_t)run * sizeof(double));
    for (int k = 1; k < nd; ++k) {  // next run: advance dims 1.. (dim 0 is the run)
      if (++idx[k] < sdim[k]) break;
      idx[k] = 0;
    }
  }
}

static int site_dims(bpx_ctx* p, int64_t v, int* udim, int* idim) {
  bpx_ctx* c = p->children[0];
  int nd = 0;
  udim[nd] = idim[nd] = p->phys_dim[v];
  ++nd;
  for (int32_t e : p->out_edge[v]) {
    udim[nd] = p->link_dim[e];
    idim[nd] = lay(c)->link_dim[e];
    ++nd;
  }
  return nd;
}

static int set_site_tensor(bpx_ctx* p, int64_t v, const void* data) {
  REQUIRE(p, v >= 0 && v < p->nv && data, "bpx_set_site_tensor: bad arguments");
  bpx_ctx* c = p->children[0];
  int udim[BPX_MAX_DEGREE + 2], idim[BPX_MAX_DEGREE + 2];
  const int nd = site_dims(p, v, udim, idim);
  std::vector<double> big((size_t)(lay(c)->site_off[v + 1] - lay(c)->site_off[v]), 0.0);
  copy_block(big.data(), idim, const_cast<double*>((const double*)data), udim, nd, true);
  return multi::fail(p, c, bpx_set_site_tensor(c, v, big.data()));
}

static int set_site_tensors(bpx_ctx* p, const void* packed) {
  REQUIRE(p, packed || p->site_off[p->nv] == 0, "bpx_set_site_tensors: NULL data");
  for (int64_t v = 0; v < p->nv; ++v) {
    const int rc = set_site_tensor(p, v, (const double*)packed + p->site_off[v]);
    if (rc) return rc;
  }
  return BPX_OK;
}

static int get_site_tensor(bpx_ctx* p, int64_t v, void* data) {
  REQUIRE(p, v >= 0 && v < p->nv && data, "bpx_get_site_tensor: bad arguments");
  bpx_ctx* c = p->children[0];
  int udim[BPX_MAX_DEGREE + 2], idim[BPX_MAX_DEGREE + 2];
  const int nd = site_dims(p, v, udim, idim);
  std::vector<double> big((size_t)(lay(c)->site_off[v + 1] - lay(c)->site_off[v]));
  const int rc = bpx_get_site_tensor(c, v, big.data());
  if (rc) return multi::fail(p, c, rc);
  copy_block(big.data(), idim, (double*)data, udim, nd, false);
  return BPX_OK;
}

// whole message sets: user packed <-> internal packed (host)
static void embed_messages(bpx_ctx* p, const double* user, std::vector<double>& big) {
  bpx_ctx* c = lay(p->children[0]);
  big.assign((size_t)c->msg_off[c->ne], 0.0);
  for (int64_t e = 0; e < p->ne; ++e) {
    const int ud[2] = {p->link_dim[e], p->link_dim[e]}, id[2] = {c->link_dim[e], c->link_dim[e]};
    copy_block(big.data() + c->msg_off[e], id, const_cast<double*>(user + p->msg_off[e]), ud, 2, true);
  }
}
static void slice_messages(bpx_ctx* p, std::vector<double>& big, double* user) {
  bpx_ctx* c = lay(p->children[0]);
  for (int64_t e = 0; e < p->ne; ++e) {
    const int ud[2] = {p->link_dim[e], p->link_dim[e]}, id[2] = {c->link_dim[e], c->link_dim[e]};
    copy_block(big.data() + c->msg_off[e], id, user + p->msg_off[e], ud, 2, false);
  }
}

static int set_messages(bpx_ctx* p, const void* packed) {
  REQUIRE(p, packed || p->msg_off[p->ne] == 0, "bpx_set_messages: NULL data");
  std::vector<double> big;
  embed_messages(p, (const double*)packed, big);
  return multi::fail(p, p->children[0], bpx_set_messages(p->children[0], big.data()));
}

static int get_messages(bpx_ctx* p, void* packed) {
  REQUIRE(p, packed || p->msg_off[p->ne] == 0, "bpx_get_messages: NULL buffer");
  bpx_ctx* c = p->children[0];
  std::vector<double> big((size_t)lay(c)->msg_off[lay(c)->ne]);
  const int rc = bpx_get_messages(c, big.data());
  if (rc) return multi::fail(p, c, rc);
  slice_messages(p, big, (double*)packed);
  return BPX_OK;
}

static int get_message(bpx_ctx* p, int64_t e, void* data) {
  REQUIRE(p, e >= 0 && e < p->ne && data, "bpx_get_message: bad arguments");
  bpx_ctx* c = p->children[0];
  std::vector<double> big((size_t)(lay(c)->msg_off[e + 1] - lay(c)->msg_off[e]));
  const int rc = bpx_get_message(c, e, big.data());
  if (rc) return multi::fail(p, c, rc);
  const int ud[2] = {p->link_dim[e], p->link_dim[e]}, id[2] = {lay(c)->link_dim[e], lay(c)->link_dim[e]};
  copy_block(big.data(), id, (double*)data, ud, 2, false);
  return BPX_OK;
}

// host iterate in, one sweep, host iterate + residual out (staged: the caller's buffers have the unpadded layout)
static int sweep_host(bpx_ctx* p, const void* packed_in, void* packed_out, int normalize, double* residual_out) {
  REQUIRE(p, packed_in && packed_out, "bpx_sweep_host: NULL buffer");
  bpx_ctx* c = p->children[0];
  int rc = set_messages(p, packed_in);
  if (rc) return rc;
  double res = 0.0;
  int done = 0;
  if ((rc = bpx_sweep(c, 1, 0.0, normalize, &res, &done))) return multi::fail(p, c, rc);
  if ((rc = get_messages(p, packed_out))) return rc;
  if (residual_out) *residual_out = res;
  return BPX_OK;
}

static int iterate_diff(bpx_ctx* p, const void* other_packed, double* out) {
  REQUIRE(p, other_packed && out, "bpx_iterate_diff: NULL argument");
  std::vector<double> big;
  embed_messages(p, (const double*)other_packed, big);
  return multi::fail(p, p->children[0], bpx_iterate_diff(p->children[0], big.data(), out));
}

// The kept rank is capped by the CALLER's link dimension (in the child the link has 16 entries and would keep up to 16
// singular values of the gated bond): gates are grouped by that dimension, one child call per group.  Singular values come
// back packed by the caller's link dims.
static int apply_two(bpx_ctx* p, int64_t n, const int64_t* edges, const void* ops, int max_rank, int normalize, double* sv_out) {
  bpx_ctx* c = p->children[0];
  REQUIRE(p, n >= 0 && max_rank >= 0 && (n == 0 || (edges && ops)), "bpx_apply_two_site_gates: bad arguments");
  if (n == 0) return BPX_OK;
  std::vector<char> used((size_t)p->nv, 0);
  std::vector<int64_t> op_off((size_t)n + 1, 0), sv_off((size_t)n + 1, 0);
  std::map<int, std::vector<int64_t>> groups;  // caller's link dim -> gates
  for (int64_t g = 0; g < n; ++g) {
    const int64_t e = edges[g];
    REQUIRE(p, e >= 0 && e < p->ne, "bpx_apply_two_site_gates: gate %lld: edge %lld out of range", (long long)g, (long long)e);
    const int32_t v1 = p->src[e], v2 = p->dst[e];
    REQUIRE(p, !used[v1] && !used[v2], "bpx_apply_two_site_gates: gate %lld shares a vertex with an earlier gate of the batch", (long long)g);
    used[v1] = used[v2] = 1;
    const int64_t dd = (int64_t)p->phys_dim[v1] * p->phys_dim[v2];
    op_off[g + 1] = op_off[g] + dd * dd;
    sv_off[g + 1] = sv_off[g] + p->link_dim[e];
    groups[p->link_dim[e]].push_back(g);
  }
  for (auto& kv : groups) {
    const int chi_u = kv.first;
    const std::vector<int64_t>& gs = kv.second;
    std::vector<int64_t> ed;
    std::vector<double> op, sv;
    int64_t svn = 0;
    for (int64_t g : gs) {
      ed.push_back(edges[g]);
      op.insert(op.end(), (const double*)ops + op_off[g], (const double*)ops + op_off[g + 1]);
      svn += lay(c)->link_dim[edges[g]];
    }
    sv.assign((size_t)svn, 0.0);
    const int k = max_rank > 0 ? std::min(max_rank, chi_u) : chi_u;
    const int rc = bpx_apply_two_site_gates(c, (int64_t)ed.size(), ed.data(), op.data(), k, normalize, sv_out ? sv.data() : nullptr);
    if (rc) return multi::fail(p, c, rc);
    if (sv_out) {
      int64_t oi = 0;
      for (int64_t g : gs) {
        memcpy(sv_out + sv_off[g], sv.data() + oi, (size_t)chi_u * sizeof(double));
        oi += lay(c)->link_dim[edges[g]];
      }
    }
  }
  return BPX_OK;
}

// ---- synthetic data: the child's device-side generators, then zero everything outside the caller's dims ------------
__global__ void mask_sites(const VDesc* vd, int64_t nv, const int32_t* udims /* [nv][BPX_MAX_DEGREE] */, double* sites) {
  for (int64_t v = blockIdx.y; v < nv; v += gridDim.y) {
    const VDesc d = vd[v];
    if (!d.owned) continue;
    const int32_t* ud = udims + v * BPX_MAX_DEGREE;
    bool padded = false;
    for (int k = 0; k < d.z; ++k) padded |= ud[k] != d.dim[k];
    if (!padded) continue;
    double* a = sites + d.site_off;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * blockDim.x) {
      int64_t rest = i / d.d;
      bool out = false;
      for (int k = 0; k < d.z; ++k) {
        out |= (int)(rest % d.dim[k]) >= ud[k];
        rest /= d.dim[k];
      }
      if (out) a[i] = 0.0;
    }
  }
}
__global__ void mask_messages(const int64_t* msg_off, const int32_t* idim, const int32_t* udim, int64_t ne, double* m0, double* m1) {
  const int64_t e = blockIdx.x;
  if (e >= ne || idim[e] == udim[e]) return;
  const int n = idim[e], u = udim[e];
  for (int i = threadIdx.x; i < n * n; i += blockDim.x)
    if (i % n >= u || i / n >= u) m0[msg_off[e] + i] = m1[msg_off[e] + i] = 0.0;
}

static int fill_synthetic(bpx_ctx* p, uint64_t seed) {
  bpx_ctx* c = p->children[0];
  int rc = bpx_fill_synthetic(c, seed);
  if (rc) return multi::fail(p, c, rc);
  std::vector<int32_t> ud((size_t)p->nv * BPX_MAX_DEGREE, 0);
  for (int64_t v = 0; v < p->nv; ++v)
    for (int k = 0; k < p->deg[v]; ++k) ud[(size_t)v * BPX_MAX_DEGREE + k] = p->link_dim[p->out_edge[v][k]];
  std::vector<bpx_ctx*> leaves = c->children.empty() ? std::vector<bpx_ctx*>{c} : c->children;
  for (bpx_ctx* l : leaves) {  // every device masks the tensors it holds and its copies of the message sets
    BPX_CUDA(p, cudaSetDevice(l->device));
    int32_t *d_ud = nullptr, *d_ul = nullptr, *d_il = nullptr;
    if ((rc = upload(l, &d_ud, ud)) || (rc = upload(l, &d_ul, p->link_dim)) || (rc = upload(l, &d_il, l->link_dim))) {
      cudaFree(d_ud);
      cudaFree(d_ul);
      return multi::fail(p, l, rc);
    }
    dim3 grid(32, (unsigned)std::min<int64_t>(std::max<int64_t>(p->nv, 1), 4096));
    mask_sites<<<grid, 256, 0, l->stream>>>(l->d_vdesc, l->nv, d_ud, (double*)l->d_sites);
    if (p->ne > 0)
      mask_messages<<<(unsigned)p->ne, 128, 0, l->stream>>>(l->d_msg_off, d_il, d_ul, p->ne, (double*)l->d_msg[0], (double*)l->d_msg[1]);
    const cudaError_t ce = cudaGetLastError(), ce2 = cudaStreamSynchronize(l->stream);
    cudaFree(d_ud);
    cudaFree(d_ul);
    cudaFree(d_il);
    l->sites_dirty = true;
    BPX_CUDA(p, ce);
    BPX_CUDA(p, ce2);
  }
  return BPX_OK;
}

}  // namespace pad
}  // namespace bpx
