// libbpx.so: context, bucketing, launch logic and the extern "C" surface declared in include/bpx.h.
// Built for sm_100a only (see build.py).  No CPU fallback: compute entry points need the device.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges cost one pointer test unless a profiler is attached

#include "bpx_common.cuh"
#include "bpx_generic.cuh"
#include "bpx_ctx.h"
#include "bpx_fast.cuh"
#include "bpx_halo.cuh"
#include "bpx_apply.cuh"
#include "bpx_apply2.cuh"
#include "bpx_apply3.cuh"
#include "bpx_expect2.cuh"

using namespace bpx;

static thread_local std::string g_create_error;

void bpx::set_error(bpx_ctx* ctx, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx)
    ctx->err = buf;
  else
    g_create_error = buf;
}

#define REQUIRE(ctx, cond, ...)           \
  do {                                    \
    if (!(cond)) {                        \
      set_error(ctx, __VA_ARGS__);        \
      return BPX_ERR_INVALID;             \
    }                                     \
  } while (0)

static int sweep_once(bpx_ctx* ctx, int normalize);
static int residual_read(bpx_ctx* ctx, int idx, double* out);
static int residual_ring_clear(bpx_ctx* ctx);
static int sweep_host_staged_enqueue(bpx_ctx* ctx, const void* packed_in, void* packed_out, int normalize);
static int sweep_host_staged_finish(bpx_ctx* ctx, double* residual_out);
#include "bpx_multi.cuh"  // single-process multi-GPU: a parent context fans every entry point out to one child per device
#define MULTI(ctx, expr)                                   \
  do {                                                     \
    if ((ctx) && !(ctx)->children.empty()) return (expr);  \
  } while (0)
#define MULTI0(ctx) ((ctx)->children[0])

template <typename T>
static int dev_alloc(bpx_ctx* ctx, T** p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
  if (e != cudaSuccess) {
    set_error(ctx, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    cudaGetLastError();
    return BPX_ERR_ALLOC;
  }
  return BPX_OK;
}

template <typename T>
static int upload(bpx_ctx* ctx, T** dptr, const std::vector<T>& h) {
  int rc = dev_alloc(ctx, dptr, h.size());
  if (rc) return rc;
  if (!h.empty()) BPX_CUDA(ctx, cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return BPX_OK;
}

// Context-cached work space: returns a buffer of at least `bytes` for role `slot`, growing it when needed.
static const size_t WS_KEEP_BYTES = (size_t)2 << 30;  // (the gate kernel wants 1.1 GiB at cfg5 shape: 296 CTAs x 3.8 MB)
template <typename T>
static int ws_get(bpx_ctx* ctx, int slot, size_t bytes, T** out) {
  if (bytes == 0) bytes = 8;
  if (ctx->ws_bytes[slot] < bytes) {
    if (ctx->ws[slot]) {
      cudaStreamSynchronize(ctx->stream);
      cudaFree(ctx->ws[slot]);
      ctx->ws[slot] = nullptr;
      ctx->ws_bytes[slot] = 0;
    }
    const size_t want = bytes <= WS_KEEP_BYTES ? std::max(bytes + bytes / 4, (size_t)4096) : bytes;
    cudaError_t e = cudaMalloc(&ctx->ws[slot], want);
    size_t got = want;
    if (e != cudaSuccess && want > bytes) {
      cudaGetLastError();
      e = cudaMalloc(&ctx->ws[slot], bytes);
      got = bytes;
    }
    if (e != cudaSuccess) {
      ctx->ws[slot] = nullptr;
      set_error(ctx, "cudaMalloc(%zu bytes of work space) failed: %s", bytes, cudaGetErrorString(e));
      cudaGetLastError();
      return BPX_ERR_ALLOC;
    }
    ctx->ws_bytes[slot] = got;
  }
  *out = (T*)ctx->ws[slot];
  return BPX_OK;
}
// end of a call: give back the buffers that are too large to keep (the stream has been synchronised)
static void ws_trim(bpx_ctx* ctx) {
  for (int i = 0; i < bpx_ctx::WS_COUNT; ++i)
    if (ctx->ws_bytes[i] > WS_KEEP_BYTES) {
      cudaFree(ctx->ws[i]);
      ctx->ws[i] = nullptr;
      ctx->ws_bytes[i] = 0;
    }
}
static void ws_release(bpx_ctx* ctx) {
  for (int i = 0; i < bpx_ctx::WS_COUNT; ++i) {
    if (ctx->ws[i]) cudaFree(ctx->ws[i]);
    ctx->ws[i] = nullptr;
    ctx->ws_bytes[i] = 0;
  }
  if (ctx->h_logsum) cudaFreeHost(ctx->h_logsum);
  ctx->h_logsum = nullptr;
}
template <typename T>
static int ws_upload(bpx_ctx* ctx, int slot, const std::vector<T>& h, T** dptr) {
  int rc = ws_get(ctx, slot, h.size() * sizeof(T), dptr);
  if (rc) return rc;
  // pageable source: the copy has returned from the host buffer's point of view when the call returns
  if (!h.empty()) BPX_CUDA(ctx, cudaMemcpyAsync(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return BPX_OK;
}

static void free_problem(bpx_ctx* c) {
  halo_release(c);
  c->rank = 0;
  c->nranks = 1;
  c->owner.clear();
  auto F = [](auto*& p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  F(c->d_vdesc);
  F(c->d_src);
  F(c->d_slot);
  F(c->d_rev);
  F(c->d_msg_off);
  F(c->d_sites);
  F(c->d_msg[0]);
  F(c->d_msg[1]);
  F(c->d_msg_snapshot);
  F(c->d_residual);
  F(c->d_reskeys);
  F(c->d_reskeys_local);
  F(c->d_generic_edges);
  F(c->d_scratch);
  F(c->d_und_edge);
  F(c->d_first_out_edge);
  F(c->d_owned_edges);
  F(c->d_owned_vertices);
  F(c->d_all_edges);
  F(c->d_fast_scratch);
  F(c->d_onchip_items);
  F(c->d_sliced_items);
  F(c->d_sliced2_items);
  F(c->d_sliced2_group_ptr);
  F(c->d_sliced2_partials);
  F(c->d_sliced2_part1);
  F(c->d_sliced2_gsync);
  c->n_sliced2_groups = c->n_sliced2_items = 0;
  F(c->d_onchip16_items);
  F(c->d_onchip16c_items);
  F(c->d_onchip8c_items);
  F(c->d_img8c);
  F(c->d_img16c);
  F(c->d_sites_swz);
  for (auto& b : c->buckets) {
    F(b.d_vertices);
    F(b.d_edges);
    F(b.d_vx_site);
    F(b.d_vx_moff);
    for (auto& ev : b.timing) {
      cudaEventDestroy(ev.first);
      cudaEventDestroy(ev.second);
    }
  }
  c->buckets.clear();
  c->dims_set = false;
}

#include "bpx_pad.cuh"  // internal zero-padding of link dims 9..15 (thin parent context + one padded child)
#define PAD(ctx, expr)                                \
  do {                                                \
    if ((ctx) && (ctx)->pad_active) return (expr);    \
  } while (0)

extern "C" int bpx_version(void) { return BPX_VERSION; }

extern "C" int bpx_create(int device, bpx_ctx** out) {
  if (!out) {
    set_error(nullptr, "bpx_create: out is NULL");
    return BPX_ERR_INVALID;
  }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error(nullptr, "bpx_create: no CUDA device (%s); libbpx has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    cudaGetLastError();
    return BPX_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    set_error(nullptr, "bpx_create: device %d out of range (0..%d)", device, ndev - 1);
    return BPX_ERR_INVALID;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    set_error(nullptr, "bpx_create: cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    return BPX_ERR_CUDA;
  }
  if (prop.major != 10) {
    set_error(nullptr, "bpx_create: device %d is sm_%d%d; libbpx is built for sm_100a only", device, prop.major,
              prop.minor);
    return BPX_ERR_UNSUPPORTED;
  }
  bpx_ctx* c = new (std::nothrow) bpx_ctx();
  if (!c) {
    set_error(nullptr, "bpx_create: out of host memory");
    return BPX_ERR_ALLOC;
  }
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    set_error(nullptr, "bpx_create: %s", cudaGetErrorString(e));
    delete c;
    return BPX_ERR_CUDA;
  }
  c->stream = c->own_stream;
  *out = c;
  return BPX_OK;
}

extern "C" int bpx_destroy(bpx_ctx* ctx) {
  MULTI(ctx, bpx::multi::destroy(ctx));
  if (!ctx) return BPX_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  halo_release(ctx);
  free_problem(ctx);
  ws_release(ctx);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->apply_stream) cudaStreamDestroy(ctx->apply_stream);
  for (auto& g : ctx->io_graphs) cudaGraphExecDestroy(g.exec);
  if (ctx->copy_stream) {
    cudaStreamDestroy(ctx->copy_stream);
    cudaEventDestroy(ctx->ev_io_start);
    cudaEventDestroy(ctx->ev_io_done);
    cudaFree(ctx->d_io_progress);
    cudaFreeHost(ctx->h_io_progress);
  }
  delete ctx;
  return BPX_OK;
}

extern "C" const char* bpx_last_error(const bpx_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

// ---- problem description -----------------------------------------------------------------------
extern "C" int bpx_set_graph(bpx_ctx* ctx, int64_t nv, int64_t ne, const int64_t* src, const int64_t* dst,
                             const int32_t* slot) {
  if (ctx && ctx->pad_active) bpx::pad::teardown(ctx);  // a new graph: back to a plain context
  MULTI(ctx, bpx::multi::each(ctx, [&](bpx_ctx* c) -> int { return bpx_set_graph(c, nv, ne, src, dst, slot); }));
  if (!ctx) return BPX_ERR_INVALID;
  REQUIRE(ctx, nv >= 0 && ne >= 0 && (ne == 0 || (src && dst && slot)), "bpx_set_graph: bad arguments");
  REQUIRE(ctx, nv < (1ll << 31) && ne < (1ll << 31), "bpx_set_graph: more than 2^31 vertices/edges");
  BPX_CUDA(ctx, cudaSetDevice(ctx->device));
  free_problem(ctx);
  ctx->graph_set = false;
  ctx->nv = nv;
  ctx->ne = ne;
  ctx->src.assign(ne, 0);
  ctx->dst.assign(ne, 0);
  ctx->slot.assign(ne, 0);
  ctx->rev.assign(ne, -1);
  ctx->deg.assign(nv, 0);
  std::map<std::pair<int64_t, int64_t>, int32_t> index;
  for (int64_t e = 0; e < ne; ++e) {
    REQUIRE(ctx, src[e] >= 0 && src[e] < nv && dst[e] >= 0 && dst[e] < nv, "bpx_set_graph: edge %lld endpoint out of range",
            (long long)e);
    REQUIRE(ctx, src[e] != dst[e], "bpx_set_graph: self loop on edge %lld", (long long)e);
    REQUIRE(ctx, index.emplace(std::make_pair(src[e], dst[e]), (int32_t)e).second, "bpx_set_graph: duplicate edge %lld",
            (long long)e);
    ctx->src[e] = (int32_t)src[e];
    ctx->dst[e] = (int32_t)dst[e];
    ctx->slot[e] = slot[e];
    ctx->deg[src[e]]++;
  }
  ctx->out_edge.assign(nv, std::vector<int32_t>());
  for (int64_t v = 0; v < nv; ++v) {
    REQUIRE(ctx, ctx->deg[v] <= BPX_MAX_DEGREE, "bpx_set_graph: vertex %lld has degree %d > BPX_MAX_DEGREE", (long long)v,
            ctx->deg[v]);
    ctx->out_edge[v].assign(ctx->deg[v], -1);
  }
  for (int64_t e = 0; e < ne; ++e) {
    auto it = index.find(std::make_pair(dst[e], src[e]));
    REQUIRE(ctx, it != index.end(), "bpx_set_graph: edge %lld has no reverse edge", (long long)e);
    ctx->rev[e] = it->second;
    const int32_t u = ctx->src[e], s = ctx->slot[e];
    REQUIRE(ctx, s >= 0 && s < ctx->deg[u], "bpx_set_graph: slot[%lld] = %d out of range", (long long)e, s);
    REQUIRE(ctx, ctx->out_edge[u][s] < 0, "bpx_set_graph: slot %d used twice at vertex %d", s, u);
    ctx->out_edge[u][s] = (int32_t)e;
  }
  ctx->graph_set = true;
  return BPX_OK;
}

static int pick_kernel(bpx_ctx* ctx, const Bucket& b);

extern "C" int bpx_set_dims(bpx_ctx* ctx, int dtype, int mode, const int32_t* phys_dim, const int32_t* link_dim) {
  if (ctx && ctx->pad_active) bpx::pad::teardown(ctx);  // dims are re-declared: decide again
  if (ctx && (dtype == BPX_F64 || dtype == BPX_C64) && (mode == BPX_MODE_NORM || mode == BPX_MODE_SINGLE)) {
    std::vector<int32_t> idim;  // link dims 9..15 of degree-4 Float64 vertices: zero-padded to 16 in a child context
    if (bpx::pad::wanted(ctx, dtype, mode, phys_dim, link_dim, idim)) return bpx::pad::wrap(ctx, dtype, mode, phys_dim, link_dim, idim);
  }
  MULTI(ctx, bpx::multi::set_dims(ctx, dtype, mode, phys_dim, link_dim));
  if (!ctx) return BPX_ERR_INVALID;
  REQUIRE(ctx, ctx->graph_set, "bpx_set_dims: call bpx_set_graph first");
  REQUIRE(ctx, dtype == BPX_F64 || dtype == BPX_C64, "bpx_set_dims: unknown dtype %d", dtype);
  REQUIRE(ctx, mode == BPX_MODE_NORM || mode == BPX_MODE_SINGLE, "bpx_set_dims: unknown mode %d", mode);
  REQUIRE(ctx, ctx->ne == 0 || link_dim, "bpx_set_dims: link_dim is NULL");
  REQUIRE(ctx, mode == BPX_MODE_SINGLE || ctx->nv == 0 || phys_dim, "bpx_set_dims: phys_dim is NULL");
  BPX_CUDA(ctx, cudaSetDevice(ctx->device));
  free_problem(ctx);
  ctx->dtype = dtype;
  ctx->mode = mode;
  ctx->esize = dtype == BPX_F64 ? 8 : 16;
  const int64_t nv = ctx->nv, ne = ctx->ne;
  ctx->link_dim.assign(link_dim, link_dim + ne);
  for (int64_t e = 0; e < ne; ++e) {
    REQUIRE(ctx, link_dim[e] >= 1, "bpx_set_dims: link_dim[%lld] < 1", (long long)e);
    REQUIRE(ctx, link_dim[e] == link_dim[ctx->rev[e]], "bpx_set_dims: link_dim differs between edge %lld and its reverse",
            (long long)e);
  }
  ctx->phys_dim.assign(nv, 1);
  if (mode == BPX_MODE_NORM)
    for (int64_t v = 0; v < nv; ++v) {
      REQUIRE(ctx, phys_dim[v] >= 1, "bpx_set_dims: phys_dim[%lld] < 1", (long long)v);
      ctx->phys_dim[v] = phys_dim[v];
    }
  // packed layouts
  ctx->site_off.assign(nv + 1, 0);
  ctx->msg_off.assign(ne + 1, 0);
  ctx->h_vdesc.assign(nv, VDesc());
  std::vector<VDesc>& vdesc = ctx->h_vdesc;
  int64_t max_n = 1;
  int max_out = 1;
  for (int64_t v = 0; v < nv; ++v) {
    VDesc& d = vdesc[v];
    memset(&d, 0, sizeof(d));
    d.z = ctx->deg[v];
    d.d = ctx->phys_dim[v];
    int64_t n = d.d;
    for (int i = 0; i < d.z; ++i) {
      const int32_t e = ctx->out_edge[v][i];
      d.dim[i] = ctx->link_dim[e];
      d.out_edge[i] = e;
      d.in_edge[i] = ctx->rev[e];
      n *= d.dim[i];
      REQUIRE(ctx, n < (1ll << 40), "bpx_set_dims: site tensor of vertex %lld too large", (long long)v);
    }
    d.n = n;
    d.site_off = ctx->site_off[v];
    ctx->site_off[v + 1] = ctx->site_off[v] + n;
    max_n = std::max(max_n, n);
  }
  for (int64_t e = 0; e < ne; ++e) {
    const int64_t chi = ctx->link_dim[e];
    const int64_t n = mode == BPX_MODE_NORM ? chi * chi : chi;
    ctx->msg_off[e + 1] = ctx->msg_off[e] + n;
    max_out = std::max<int64_t>(max_out, n);
  }
  ctx->max_site_elems = max_n;
  ctx->max_msg_elems = max_out;

  int rc;
  if ((rc = upload(ctx, &ctx->d_src, ctx->src))) return rc;
  if ((rc = upload(ctx, &ctx->d_slot, ctx->slot))) return rc;
  if ((rc = upload(ctx, &ctx->d_rev, ctx->rev))) return rc;
  if ((rc = upload(ctx, &ctx->d_msg_off, ctx->msg_off))) return rc;
  const size_t es = ctx->esize;
  for (int k = 0; k < 2; ++k) {
    if ((rc = dev_alloc(ctx, (char**)&ctx->d_msg[k], (size_t)ctx->msg_off[ne] * es))) return rc;
    BPX_CUDA(ctx, cudaMemset(ctx->d_msg[k], 0, std::max<size_t>(1, (size_t)ctx->msg_off[ne] * es)));
  }
  ctx->cur = 0;
  if ((rc = dev_alloc(ctx, &ctx->d_residual, (size_t)ne))) return rc;
  ctx->history_cap = 4096;
  if ((rc = dev_alloc(ctx, &ctx->d_reskeys, (size_t)ctx->history_cap + 1))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_reskeys_local, (size_t)ctx->history_cap + 1))) return rc;
  BPX_CUDA(ctx, cudaMemset(ctx->d_reskeys, 0, ((size_t)ctx->history_cap + 1) * sizeof(unsigned long long)));
  BPX_CUDA(ctx, cudaMemset(ctx->d_reskeys_local, 0, ((size_t)ctx->history_cap + 1) * sizeof(unsigned long long)));
  ctx->history_len = 0;
  std::vector<int32_t> und;
  for (int64_t e = 0; e < ne; ++e)
    if (e < ctx->rev[e]) und.push_back((int32_t)e);
  ctx->n_und = (int64_t)und.size();
  if ((rc = upload(ctx, &ctx->d_und_edge, und))) return rc;

  // generic-kernel geometry: shared memory if two tensor copies (+ one output) fit, else global scratch
  const int64_t smem_need = (2 * max_n + max_out) * (int64_t)es;
  if (smem_need <= (int64_t)ctx->max_smem_optin - 2048) {
    ctx->gen_smem_elems = (int)max_n;
    ctx->gen_smem_bytes = (int)smem_need;
    ctx->gen_grid = ctx->num_sms * 8;
  } else {
    ctx->gen_smem_elems = 0;
    ctx->gen_smem_bytes = (int)(max_out * es);
    ctx->gen_grid = ctx->num_sms * 4;
    if ((rc = dev_alloc(ctx, (char**)&ctx->d_scratch, (size_t)ctx->gen_grid * 2 * max_n * es))) return rc;
  }

  // buckets: vertices sharing (degree, phys dim, link dims)
  std::map<std::vector<int32_t>, int> bucket_of;
  for (int64_t v = 0; v < nv; ++v) {
    std::vector<int32_t> key;
    key.push_back(ctx->deg[v]);
    key.push_back(ctx->phys_dim[v]);
    for (int i = 0; i < ctx->deg[v]; ++i) key.push_back(vdesc[v].dim[i]);
    auto it = bucket_of.find(key);
    if (it == bucket_of.end()) {
      it = bucket_of.emplace(key, (int)ctx->buckets.size()).first;
      Bucket b;
      b.z = ctx->deg[v];
      b.d = ctx->phys_dim[v];
      b.chi = b.z > 0 ? vdesc[v].dim[0] : 0;
      for (int i = 1; i < b.z; ++i)
        if (vdesc[v].dim[i] != b.chi) b.chi = 0;
      for (int i = 0; i < b.z; ++i) b.max_dim = std::max<int>(b.max_dim, vdesc[v].dim[i]);
      ctx->buckets.push_back(b);
    }
    ctx->buckets[it->second].vertices.push_back((int32_t)v);
  }
  ctx->owner.clear();
  ctx->rank = 0;
  ctx->nranks = 1;
  ctx->dev_site_off.clear();
  ctx->dev_site_total = 0;
  ctx->dims_set = true;
  if ((rc = relayout_sites(ctx))) return rc;
  return rebuild_work_lists(ctx);
}

// Site tensors live on the rank that owns the vertex.  (Re)build the compact device layout for the current owner
// map, move tensors that are already resident, refresh the device descriptors.
int bpx::relayout_sites(bpx_ctx* ctx) {
  const int64_t nv = ctx->nv;
  const size_t es = ctx->esize;
  std::vector<int64_t> old_off = ctx->dev_site_off;
  void* old = ctx->d_sites;
  ctx->dev_site_off.assign(nv + 1, -1);
  int64_t total = 0;
  for (int64_t v = 0; v < nv; ++v) {
    const bool mine = ctx->owner.empty() || ctx->owner[v] == ctx->rank;
    if (mine) {
      ctx->dev_site_off[v] = total;
      total += ctx->site_off[v + 1] - ctx->site_off[v];
    }
    ctx->h_vdesc[v].owned = mine ? 1 : 0;
    ctx->h_vdesc[v].site_off = mine ? ctx->dev_site_off[v] : 0;
  }
  ctx->dev_site_total = total;
  ctx->d_sites = nullptr;
  int rc = dev_alloc(ctx, (char**)&ctx->d_sites, (size_t)total * es);
  if (rc) {
    ctx->d_sites = old;
    return rc;
  }
  if (old && !old_off.empty()) {
    // keep what is already resident (bpx_set_partition after an upload): copy runs of consecutive vertices
    int64_t v = 0;
    while (v < nv) {
      if (ctx->dev_site_off[v] < 0 || old_off[v] < 0) {
        ++v;
        continue;
      }
      int64_t w = v;
      while (w + 1 < nv && ctx->dev_site_off[w + 1] >= 0 && old_off[w + 1] >= 0 &&
             old_off[w + 1] - old_off[v] == ctx->dev_site_off[w + 1] - ctx->dev_site_off[v])
        ++w;
      const size_t bytes = (size_t)(ctx->dev_site_off[w] - ctx->dev_site_off[v] + ctx->site_off[w + 1] - ctx->site_off[w]) * es;
      BPX_CUDA(ctx, cudaMemcpyAsync((char*)ctx->d_sites + (size_t)ctx->dev_site_off[v] * es, (char*)old + (size_t)old_off[v] * es, bytes,
                                    cudaMemcpyDeviceToDevice, ctx->stream));
      v = w + 1;
    }
    BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (old) cudaFree(old);
  if (ctx->d_vdesc) {
    cudaFree(ctx->d_vdesc);
    ctx->d_vdesc = nullptr;
  }
  if ((rc = upload(ctx, &ctx->d_vdesc, ctx->h_vdesc))) return rc;
  ctx->sites_dirty = true;
  return BPX_OK;
}

// (re)derive per-bucket vertex/edge lists restricted to the vertices this rank owns
int bpx::rebuild_work_lists(bpx_ctx* ctx) {
  ctx->work_epoch++;  // captured steps refer to the old lists / buffers
  auto F = [](auto*& p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  std::vector<int32_t> owned_edges;
  for (auto& b : ctx->buckets) {
    F(b.d_vertices);
    F(b.d_edges);
    b.my_vertices.clear();
    b.my_edges.clear();
    for (int32_t v : b.vertices) {
      if (!ctx->owner.empty() && ctx->owner[v] != ctx->rank) continue;
      b.my_vertices.push_back(v);
      for (int32_t e : ctx->out_edge[v]) b.my_edges.push_back(e);
    }
    int rc;
    if ((rc = upload(ctx, &b.d_vertices, b.my_vertices))) return rc;
    if ((rc = upload(ctx, &b.d_edges, b.my_edges))) return rc;
    b.kernel = pick_kernel(ctx, b);
    owned_edges.insert(owned_edges.end(), b.my_edges.begin(), b.my_edges.end());
  }
  std::sort(owned_edges.begin(), owned_edges.end());
  {
    std::vector<int32_t> ge;
    for (auto& b : ctx->buckets)
      if (b.kernel == BPX_KERNEL_GENERIC) ge.insert(ge.end(), b.my_edges.begin(), b.my_edges.end());
    std::sort(ge.begin(), ge.end());
    F(ctx->d_generic_edges);
    ctx->n_generic_edges = (int64_t)ge.size();
    int rcg = upload(ctx, &ctx->d_generic_edges, ge);
    if (rcg) return rcg;
  }
  {
    std::vector<int32_t> ov;
    for (int64_t v = 0; v < ctx->nv; ++v)
      if (ctx->owner.empty() || ctx->owner[v] == ctx->rank) ov.push_back((int32_t)v);
    F(ctx->d_owned_vertices);
    ctx->n_owned_vertices = (int64_t)ov.size();
    int rcv = upload(ctx, &ctx->d_owned_vertices, ov);
    if (rcv) return rcv;
  }
  F(ctx->d_owned_edges);
  ctx->n_owned_edges = (int64_t)owned_edges.size();
  int rc = upload(ctx, &ctx->d_owned_edges, owned_edges);
  if (rc) return rc;
  ctx->owned_runs.clear();
  ctx->upload_end.assign(ctx->ne, 0);
  ctx->owned_elems = 0;
  for (int32_t e : owned_edges) {  // sorted by edge id
    const int64_t b = ctx->msg_off[e], en = ctx->msg_off[e + 1];
    if (en == b) continue;
    if (!ctx->owned_runs.empty() && ctx->owned_runs.back().second == b) ctx->owned_runs.back().second = en;
    else ctx->owned_runs.emplace_back(b, en);
    ctx->owned_elems += en - b;
    ctx->upload_end[e] = ctx->owned_elems;
  }
  return fast_prepare(ctx);
}

static int pick_kernel(bpx_ctx* ctx, const Bucket& b) {
  const int want = ctx->kernel_policy;
  if (want == BPX_KERNEL_GENERIC) return BPX_KERNEL_GENERIC;
  const int best = fast_kernel_for(ctx, b);  // BPX_KERNEL_GENERIC if no specialised kernel applies
  if (want == BPX_KERNEL_AUTO) return best;
  return fast_kernel_supported(ctx, b, want) ? want : BPX_KERNEL_GENERIC;
}

extern "C" int64_t bpx_num_vertices(const bpx_ctx* ctx) { return !ctx ? -1 : (ctx->children.empty() ? ctx->nv : ctx->children[0]->nv); }
extern "C" int64_t bpx_num_edges(const bpx_ctx* ctx) { return !ctx ? -1 : (ctx->children.empty() ? ctx->ne : ctx->children[0]->ne); }
extern "C" int64_t bpx_rev(const bpx_ctx* ctx, int64_t e) {
  if (ctx && !ctx->children.empty()) return bpx_rev(ctx->children[0], e);
  return (ctx && ctx->graph_set && e >= 0 && e < ctx->ne) ? ctx->rev[e] : -1;
}
extern "C" int64_t bpx_site_offset(const bpx_ctx* ctx, int64_t v) {
  if (ctx && ctx->pad_active) return (v >= 0 && v <= ctx->nv) ? ctx->site_off[v] : -1;  // the caller's layout
  if (ctx && !ctx->children.empty()) return bpx_site_offset(ctx->children[0], v);
  return (ctx && ctx->dims_set && v >= 0 && v <= ctx->nv) ? ctx->site_off[v] : -1;
}
extern "C" int64_t bpx_site_device_offset(const bpx_ctx* ctx, int64_t v) {
  if (ctx && !ctx->children.empty()) return (v < 0 || v >= ctx->nv) ? -1 : bpx_site_device_offset(ctx->children[ctx->multi_owner.empty() ? 0 : ctx->multi_owner[v]], v);
  return (ctx && ctx->dims_set && v >= 0 && v < ctx->nv) ? ctx->dev_site_off[v] : -1;
}
extern "C" int64_t bpx_message_offset(const bpx_ctx* ctx, int64_t e) {
  if (ctx && ctx->pad_active) return (e >= 0 && e <= ctx->ne) ? ctx->msg_off[e] : -1;  // the caller's layout
  if (ctx && !ctx->children.empty()) return bpx_message_offset(ctx->children[0], e);
  return (ctx && ctx->dims_set && e >= 0 && e <= ctx->ne) ? ctx->msg_off[e] : -1;
}

#define NEED_DIMS(ctx, name)                                     \
  do {                                                           \
    if (!ctx) return BPX_ERR_INVALID;                            \
    REQUIRE(ctx, ctx->dims_set, name ": call bpx_set_dims first"); \
    BPX_CUDA(ctx, cudaSetDevice(ctx->device));                   \
  } while (0)

extern "C" int bpx_set_site_tensors(bpx_ctx* ctx, const void* packed) {
  PAD(ctx, bpx::pad::set_site_tensors(ctx, packed));
  MULTI(ctx, bpx::multi::each(ctx, [&](bpx_ctx* c) -> int { return bpx_set_site_tensors(c, packed); }));
  NEED_DIMS(ctx, "bpx_set_site_tensors");
  REQUIRE(ctx, packed || ctx->site_off[ctx->nv] == 0, "bpx_set_site_tensors: NULL data");
  // a rank only needs the tensors it owns, but uploading all keeps offsets identical everywhere
  ctx->sites_dirty = true;
  // runs of consecutive owned vertices are contiguous on both sides (a single rank: one copy)
  const int64_t nv = ctx->nv;
  int64_t v = 0;
  while (v < nv) {
    if (ctx->dev_site_off[v] < 0) {
      ++v;
      continue;
    }
    int64_t w = v;
    while (w + 1 < nv && ctx->dev_site_off[w + 1] >= 0) ++w;
    const size_t bytes = (size_t)(ctx->site_off[w + 1] - ctx->site_off[v]) * ctx->esize;
    BPX_CUDA(ctx, cudaMemcpyAsync((char*)ctx->d_sites + (size_t)ctx->dev_site_off[v] * ctx->esize,
                                  (const char*)packed + (size_t)ctx->site_off[v] * ctx->esize, bytes, cudaMemcpyHostToDevice, ctx->stream));
    v = w + 1;
  }
  BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPX_OK;
}

extern "C" int bpx_set_site_tensor(bpx_ctx* ctx, int64_t v, const void* data) {
  PAD(ctx, bpx::pad::set_site_tensor(ctx, v, data));
  MULTI(ctx, (v < 0 || v >= ctx->nv) ? (int)BPX_ERR_INVALID : bpx::multi::fail(ctx, bpx::multi::owner_of(ctx, v), bpx_set_site_tensor(bpx::multi::owner_of(ctx, v), v, data)));
  NEED_DIMS(ctx, "bpx_set_site_tensor");
  REQUIRE(ctx, v >= 0 && v < ctx->nv && data, "bpx_set_site_tensor: bad arguments");
  if (ctx->dev_site_off[v] < 0) return BPX_OK;  // not resident on this rank
  const size_t off = (size_t)ctx->dev_site_off[v] * ctx->esize, n = (size_t)(ctx->site_off[v + 1] - ctx->site_off[v]) * ctx->esize;
  ctx->sites_dirty = true;
  BPX_CUDA(ctx, cudaMemcpyAsync((char*)ctx->d_sites + off, data, n, cudaMemcpyHostToDevice, ctx->stream));
  BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPX_OK;
}

// Connected ranks of a partitioned run store halo messages straight into each other's message sets during a sweep.  A call
// that REWRITES this rank's sets (bpx_set_messages, bpx_fill_synthetic) is therefore collective: it ends with a cross-rank
// barrier on the stream, so that no rank starts sweeping -- and storing into this rank's halo slots -- before every rank
// has finished rewriting (seen on two B200s with a cold second rank: the late rank's upload overwrote the first halo
// messages of the early one, a wrong third sweep).  One process driving several devices (bpx_create_multi) serialises
// these calls on the host and needs no barrier.
static int collective_fence(bpx_ctx* ctx) {
  if (ctx->nranks > 1 && ctx->halo_connected && !ctx->is_child) return bpx_peer_barrier(ctx);
  return BPX_OK;
}

extern "C" int bpx_set_messages(bpx_ctx* ctx, const void* packed) {
  PAD(ctx, bpx::pad::set_messages(ctx, packed));
  MULTI(ctx, bpx::multi::each(ctx, [&](bpx_ctx* c) -> int { return bpx_set_messages(c, packed); }));
  NEED_DIMS(ctx, "bpx_set_messages");
  REQUIRE(ctx, packed || ctx->msg_off[ctx->ne] == 0, "bpx_set_messages: NULL data");
  const size_t n = (size_t)ctx->msg_off[ctx->ne] * ctx->esize;
  ctx->cur = 0;  // all ranks of a partitioned run restart on the same parity
  BPX_CUDA(ctx, cudaMemcpyAsync(ctx->d_msg[ctx->cur], packed, n, cudaMemcpyHostToDevice, ctx->stream));
  // partitioned runs: both sets must hold every message, so that edges a rank does not own keep their value across
  // ping-pong (a single rank rewrites every message each sweep, and the sequential path re-syncs the sets itself)
  if (ctx->nranks > 1)
    BPX_CUDA(ctx, cudaMemcpyAsync(ctx->d_msg[ctx->cur ^ 1], ctx->d_msg[ctx->cur], n, cudaMemcpyDeviceToDevice, ctx->stream));
  const int rc = collective_fence(ctx);
  if (rc) return rc;
  BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPX_OK;
}

// One reference-facing step with HOST buffers: upload the iterate, one synchronous sweep, download the new iterate
// and the fused residual -- the per-sweep call of the Julia plugin (AI.step! + StopWhenConverged), one host sync.
// device-accessible alias of a pinned (page-locked, mapped) host pointer, or NULL
static void* mapped_alias(const void* host) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

// can this rank's sweep run with streamed host I/O?  (every launch is a kernel family that implements HostIO)
static bool sweep_streams_host_io(bpx_ctx* ctx) {
  int n = 0;
  for (int bi = 0; bi < (int)ctx->buckets.size(); ++bi) {
    const Bucket& b = ctx->buckets[bi];
    if (b.my_edges.empty()) continue;
    if (b.kernel != BPX_KERNEL_ONCHIP && b.kernel != BPX_KERNEL_SLICED) return false;
    if (b.leader == bi) ++n;
  }
  return n >= 1;
}

static void io_graphs_clear(bpx_ctx* ctx) {
  for (auto& g : ctx->io_graphs) cudaGraphExecDestroy(g.exec);
  ctx->io_graphs.clear();
}

// Enqueue one streamed step (also the body that is captured into a CUDA graph): progress word and residual slot
// reset, the sweep kernel (gated on the progress word, storing into the host alias), the chunked upload on the copy
// stream, and the residual key into pinned host memory.
// cuStreamWriteValue32 through the runtime's driver entry point (libbpx links cudart only): a stream memory op is a
// much cheaper "chunk has landed" signal than an 8-byte copy behind every chunk.  NULL if unavailable.
typedef int (*StreamWriteValue32Fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
static StreamWriteValue32Fn stream_write_value32() {
  static StreamWriteValue32Fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (!getenv("BPX_IO_NO_MEMOP") && cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &st) == cudaSuccess &&
        st == cudaDriverEntryPointSuccess)
      fn = (StreamWriteValue32Fn)p;
    cudaGetLastError();
  }
  return fn;
}

// chunks of the owned runs for the streamed upload: (element begin, element end, cumulative uploaded elements)
struct IoChunk {
  int64_t b, e, cum;
};
static std::vector<IoChunk> io_chunks(bpx_ctx* ctx, int64_t chunk_elems) {
  std::vector<IoChunk> out;
  int64_t cum = 0;
  for (auto& r : ctx->owned_runs)
    for (int64_t b = r.first; b < r.second;) {
      int64_t e = std::min(r.second, b + chunk_elems);
      if (r.second - e < chunk_elems / 4) e = r.second;  // no tiny tail chunks
      if (e < r.second) e &= ~(int64_t)1;                // whole 16-byte units
      cum += e - b;
      out.push_back({b, e, cum});
      b = e;
    }
  return out;
}

// Enqueue one streamed step (also the body that is captured into a CUDA graph): the sweep kernel (gated on the
// progress word, storing into the host alias, handing the residual key to the host) and the chunked upload on the
// copy stream.
static int enqueue_streamed_step(bpx_ctx* ctx, const void* packed_in, void* out_alias, int* err_alias, int normalize,
                                 const std::vector<IoChunk>& chunks) {
  const bool single = ctx->nranks == 1;
  if (single) {
    ctx->cur = 0;
    ctx->history_len = 0;  // the step's residual lands in history slot 0 (a plain store by the kernel's last CTA)
    ctx->ring_dirty = false;
  }
  // no memsets: the progress word, the CTA ticket and the step's private residual slot are reset by the previous
  // step's last CTA (and zero-initialised)
  BPX_CUDA(ctx, cudaEventRecord(ctx->ev_io_start, ctx->stream));
  BPX_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_io_start, 0));
  ctx->io_args.progress = ctx->d_io_progress;
  ctx->io_args.host_out = (double*)out_alias;
  ctx->io_args.error_flag = err_alias;
  ctx->io_args.ticket = reinterpret_cast<unsigned int*>(ctx->d_io_progress + 1);
  if (single) {  // partitioned runs keep the residual ring / mailbox protocol (the global maximum needs the peers' posts)
    ctx->io_args.host_key = reinterpret_cast<unsigned long long*>(err_alias) - 1;  // pinned slot [32] (the error flag is [33])
    ctx->io_args.local_key = reinterpret_cast<unsigned long long*>(ctx->d_io_progress + 2);
    ctx->io_args.ring_key = ctx->d_reskeys;
    ctx->slot_override = ctx->io_args.local_key;
  }
  void* const dst_set = ctx->d_msg[ctx->cur];
  int rc = sweep_once(ctx, normalize);
  ctx->slot_override = nullptr;
  ctx->io_args = HostIO{};
  if (rc) return rc;
  StreamWriteValue32Fn wv = ctx->owned_elems < (1ll << 32) ? stream_write_value32() : nullptr;
  for (size_t c = 0; c < chunks.size(); ++c) {
    const size_t o = (size_t)chunks[c].b * ctx->esize, len = (size_t)(chunks[c].e - chunks[c].b) * ctx->esize;
    BPX_CUDA(ctx, cudaMemcpyAsync((char*)dst_set + o, (const char*)packed_in + o, len, cudaMemcpyHostToDevice, ctx->copy_stream));
    if (getenv("BPX_IO_TEST_STALL")) continue;  // (tests: the upload never reports progress -> the kernel gives up)
    if (wv) {  // low word of the (zeroed) 64-bit progress counter
      if (wv(ctx->copy_stream, (unsigned long long)(uintptr_t)ctx->d_io_progress, (unsigned int)chunks[c].cum, 0) != 0) {
        set_error(ctx, "bpx_sweep_host: cuStreamWriteValue32 failed");
        return BPX_ERR_CUDA;
      }
    } else {
      ctx->h_io_progress[c] = chunks[c].cum;
      BPX_CUDA(ctx, cudaMemcpyAsync(ctx->d_io_progress, ctx->h_io_progress + c, sizeof(long long), cudaMemcpyHostToDevice, ctx->copy_stream));
    }
  }
  BPX_CUDA(ctx, cudaEventRecord(ctx->ev_io_done, ctx->copy_stream));
  BPX_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_io_done, 0));
  return BPX_OK;  // (single rank: the residual key arrives in pinned host memory from the kernel's last CTA)
}

// staged host step, first half: everything is enqueued, nothing waits (a multi-device context enqueues all its devices
// before the first of them waits for its peers)
static int sweep_host_staged_enqueue(bpx_ctx* ctx, const void* packed_in, void* packed_out, int normalize) {
  int rc;
  if (ctx->nranks == 1) ctx->cur = 0;
  if ((rc = halo_gate(ctx))) return rc;
  for (auto& r : ctx->owned_runs) {
    const size_t o = (size_t)r.first * ctx->esize, len = (size_t)(r.second - r.first) * ctx->esize;
    BPX_CUDA(ctx, cudaMemcpyAsync((char*)ctx->d_msg[ctx->cur] + o, (const char*)packed_in + o, len, cudaMemcpyHostToDevice, ctx->stream));
  }
  for (auto& r : ctx->halo_in_runs) {  // (children of a multi-device context: the peers' cut-edge messages, from the host too)
    const size_t o = (size_t)r.first * ctx->esize, len = (size_t)(r.second - r.first) * ctx->esize;
    BPX_CUDA(ctx, cudaMemcpyAsync((char*)ctx->d_msg[ctx->cur] + o, (const char*)packed_in + o, len, cudaMemcpyHostToDevice, ctx->stream));
  }
  if ((rc = sweep_once(ctx, normalize))) return rc;
  for (auto& r : ctx->owned_runs) {
    const size_t o = (size_t)r.first * ctx->esize, len = (size_t)(r.second - r.first) * ctx->esize;
    BPX_CUDA(ctx, cudaMemcpyAsync((char*)packed_out + o, (const char*)ctx->d_msg[ctx->cur] + o, len, cudaMemcpyDeviceToHost, ctx->stream));
  }
  return BPX_OK;
}
// ... second half: the (global) residual; synchronises the stream
static int sweep_host_staged_finish(bpx_ctx* ctx, double* residual_out) {
  int rc;
  if ((rc = halo_gate(ctx))) return rc;
  return residual_read(ctx, ctx->history_len - 1, residual_out);
}

extern "C" int bpx_sweep_host(bpx_ctx* ctx, const void* packed_in, void* packed_out, int normalize, double* residual_out) {
  PAD(ctx, bpx::pad::sweep_host(ctx, packed_in, packed_out, normalize, residual_out));
  MULTI(ctx, bpx::multi::sweep_host(ctx, packed_in, packed_out, normalize, residual_out));
  NEED_DIMS(ctx, "bpx_sweep_host");
  REQUIRE(ctx, (packed_in && packed_out) || ctx->msg_off[ctx->ne] == 0, "bpx_sweep_host: NULL buffer");
  REQUIRE(ctx, ctx->nranks == 1 || ctx->halo_connected, "bpx_sweep_host: partitioned context without connected peers");
  const size_t n = (size_t)ctx->owned_elems * ctx->esize;
  const bool single = ctx->nranks == 1;
  int rc;
  double res = INFINITY;
  void* out_alias =
      (n > 0 && !ctx->io_stream_disabled && packed_in != packed_out && sweep_streams_host_io(ctx) && mapped_alias(packed_in))
          ? mapped_alias(packed_out)
          : nullptr;
  const int cur0 = ctx->cur;
  if (out_alias) {
    // ---- streamed: the kernel starts at once; the upload arrives in chunks behind a progress word the items wait
    // for, and the epilogues store the new messages straight into the caller's buffer.  On a single rank the whole
    // step is one CUDA-graph launch (cached per buffer pair): the ~20 stream calls it replaces cost more than the
    // sweep.  Partitioned contexts (sweep ids in the kernel arguments) enqueue the same operations on the streams. ----
    // every chunk costs ~10 us of copy-engine latency (measured): about 1 MiB per chunk, at most 16
    int n_chunks = (int)std::max<size_t>(1, std::min<size_t>(16, (n + (512 << 10)) >> 20));
    if (const char* env = getenv("BPX_IO_CHUNKS")) n_chunks = std::max(1, std::min(32, atoi(env)));
    if (!ctx->copy_stream) {
      BPX_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
      BPX_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_io_start, cudaEventDisableTiming));
      BPX_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_io_done, cudaEventDisableTiming));
      BPX_CUDA(ctx, cudaMalloc((void**)&ctx->d_io_progress, 4 * sizeof(long long)));  // [0] progress, [1] CTA ticket, [2] residual key
      BPX_CUDA(ctx, cudaMemset(ctx->d_io_progress, 0, 4 * sizeof(long long)));
      BPX_CUDA(ctx, cudaHostAlloc((void**)&ctx->h_io_progress, 34 * sizeof(long long), cudaHostAllocMapped));
      ctx->h_io_progress[33] = 0;
    }
    if ((rc = fast_refresh_sites(ctx))) return rc;  // never part of the captured step
    int* err_alias = (int*)mapped_alias(ctx->h_io_progress + 33);
    std::vector<IoChunk> chunks = io_chunks(ctx, std::max<int64_t>(2, (ctx->owned_elems + n_chunks - 1) / n_chunks));
    if (chunks.size() > 32) chunks = io_chunks(ctx, ctx->owned_elems);  // (many runs: one chunk per run at most)
    if (getenv("BPX_IO_DEBUG"))
      fprintf(stderr, "[bpx io] owned_elems %lld chunks %zu first cum %lld last cum %lld esize %d alias %p\n", (long long)ctx->owned_elems,
              chunks.size(), (long long)chunks.front().cum, (long long)chunks.back().cum, ctx->esize, out_alias);
    const bool use_graph = single && chunks.size() <= 32 && !ctx->profiling && !ctx->io_graph_disabled && !getenv("BPX_IO_NO_GRAPH");
    if (use_graph) {
      bpx_ctx::IoGraph* g = nullptr;
      for (auto& c : ctx->io_graphs)
        if (c.in == packed_in && c.out == packed_out && c.normalize == normalize && c.chunks == (int)chunks.size() && c.epoch == ctx->work_epoch)
          g = &c;
      const int64_t sweeps0 = ctx->n_sweeps, updates0 = ctx->n_updates, launches0 = ctx->n_launches;
      if (!g) {
        if (ctx->io_graphs.size() >= 8) io_graphs_clear(ctx);
        cudaGraph_t graph = nullptr;
        BPX_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        rc = enqueue_streamed_step(ctx, packed_in, out_alias, err_alias, normalize, chunks);
        cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
        if (rc) {
          if (graph) cudaGraphDestroy(graph);
          return rc;
        }
        BPX_CUDA(ctx, ce);
        cudaGraphExec_t exec = nullptr;
        ce = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        BPX_CUDA(ctx, ce);
        ctx->io_graphs.push_back({packed_in, packed_out, normalize, (int)chunks.size(), ctx->work_epoch, ctx->n_launches - launches0, exec});
        g = &ctx->io_graphs.back();
      }
      // host-side state of "one sweep from set 0 into set 1, recorded in residual slot 0"
      ctx->cur = 1;
      ctx->history_len = 1;
      ctx->n_sweeps = sweeps0 + 1;
      ctx->n_updates = updates0 + ctx->n_owned_edges;
      ctx->n_launches = launches0 + g->launches;
      for (size_t c = 0; c < chunks.size(); ++c) ctx->h_io_progress[c] = chunks[c].cum;  // (read by the copy engine at run time)
      BPX_CUDA(ctx, cudaGraphLaunch(g->exec, ctx->stream));
    } else {
      if ((rc = enqueue_streamed_step(ctx, packed_in, out_alias, err_alias, normalize, chunks))) return rc;
    }
    if (single) {
      BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      ctx->ring_dirty = true;  // slot 0 is re-used by every streamed step: other sweep entry points clear the ring first
      unsigned long long key;
      memcpy(&key, ctx->h_io_progress + 32, sizeof(key));
      res = residual_from_key(key);
    } else {
      if ((rc = halo_gate(ctx))) return rc;  // the global residual needs every rank's post of this sweep
      if ((rc = residual_read(ctx, ctx->history_len - 1, &res))) return rc;  // synchronises the stream
    }
    if (ctx->h_io_progress[33] != 0) {
      // The kernel gave up waiting for the upload: something serialises kernel and copies (a profiler replaying
      // kernels, a runtime that does not overlap the graph's branches).  The upload itself has completed by now; the
      // results of this attempt are discarded and the step is repeated staged -- as every later step of this context.
      ctx->h_io_progress[33] = 0;
      BPX_CUDA(ctx, cudaMemset(ctx->d_io_progress, 0, 4 * sizeof(long long)));  // the self-resetting words are stale
      ctx->io_stream_disabled = true;
      if (single) ctx->ring_dirty = true;
      else {
        set_error(ctx, "bpx_sweep_host: timed out waiting for the streamed upload on a partitioned context");
        return BPX_ERR_CUDA;  // (the peers have seen this sweep's pushes: it cannot be repeated locally)
      }
      ctx->cur = cur0;
      out_alias = nullptr;
    }
  }
  if (!out_alias) {
    // ---- staged: upload the owned messages, sweep, download the owned messages ----
    if ((rc = sweep_host_staged_enqueue(ctx, packed_in, packed_out, normalize))) return rc;
    if ((rc = sweep_host_staged_finish(ctx, &res))) return rc;
  }
  if (residual_out) *residual_out = res;
  return BPX_OK;
}

extern "C" int bpx_host_register(bpx_ctx* ctx, void* ptr, size_t bytes) {
  MULTI(ctx, bpx::multi::fail(ctx, MULTI0(ctx), bpx_host_register(MULTI0(ctx), ptr, bytes)));
  if (!ctx) return BPX_ERR_INVALID;
  REQUIRE(ctx, ptr && bytes > 0, "bpx_host_register: bad arguments");
  cudaSetDevice(ctx->device);
  BPX_CUDA(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
  return BPX_OK;
}

extern "C" int bpx_host_unregister(bpx_ctx* ctx, void* ptr) {
  MULTI(ctx, bpx::multi::fail(ctx, MULTI0(ctx), bpx_host_unregister(MULTI0(ctx), ptr)));
  if (!ctx) return BPX_ERR_INVALID;
  REQUIRE(ctx, ptr != nullptr, "bpx_host_unregister: NULL pointer");
  cudaSetDevice(ctx->device);
  io_graphs_clear(ctx);  // captured steps may refer to the buffer
  BPX_CUDA(ctx, cudaHostUnregister(ptr));
  return BPX_OK;
}

extern "C" int bpx_get_messages(bpx_ctx* ctx, void* packed) {
  PAD(ctx, bpx::pad::get_messages(ctx, packed));
  MULTI(ctx, bpx::multi::get_messages(ctx, packed));
  NEED_DIMS(ctx, "bpx_get_messages");
  { int rc_ = halo_gate(ctx); if (rc_) return rc_; }
  REQUIRE(ctx, packed || ctx->msg_off[ctx->ne] == 0, "bpx_get_messages: NULL data");
  BPX_CUDA(ctx, cudaMemcpyAsync(packed, ctx->d_msg[ctx->cur], (size_t)ctx->msg_off[ctx->ne] * ctx->esize, cudaMemcpyDeviceToHost,
                                ctx->stream));
  BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPX_OK;
}

extern "C" int bpx_get_message(bpx_ctx* ctx, int64_t e, void* data) {
  PAD(ctx, bpx::pad::get_message(ctx, e, data));
  MULTI(ctx, (e < 0 || e >= ctx->ne) ? (int)BPX_ERR_INVALID : bpx::multi::fail(ctx, bpx::multi::owner_of(ctx, MULTI0(ctx)->src[e]), bpx_get_message(bpx::multi::owner_of(ctx, MULTI0(ctx)->src[e]), e, data)));
  NEED_DIMS(ctx, "bpx_get_message");
  { int rc_ = halo_gate(ctx); if (rc_) return rc_; }
  REQUIRE(ctx, e >= 0 && e < ctx->ne && data, "bpx_get_message: bad arguments");
  const size_t off = (size_t)ctx->msg_off[e] * ctx->esize, n = (size_t)(ctx->msg_off[e + 1] - ctx->msg_off[e]) * ctx->esize;
  BPX_CUDA(ctx, cudaMemcpyAsync(data, (char*)ctx->d_msg[ctx->cur] + off, n, cudaMemcpyDeviceToHost, ctx->stream));
  BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPX_OK;
}

// ---- launches -----------------------------------------------------------------------------------
static GenericArgs generic_args(bpx_ctx* ctx, const void* msg_in, void* msg_out, const int32_t* work, int64_t n_work,
                                int normalize) {
  GenericArgs g;
  memset(&g, 0, sizeof(g));
  g.vdesc = ctx->d_vdesc;
  g.src = ctx->d_src;
  g.slot = ctx->d_slot;
  g.msg_off = ctx->d_msg_off;
  g.sites = ctx->d_sites;
  g.msg_in = msg_in;
  g.msg_out = msg_out;
  g.residual = ctx->d_residual;
  g.work = work;
  g.n_work = n_work;
  g.scratch = ctx->d_scratch;
  g.scratch_elems = ctx->max_site_elems;
  g.smem_elems = ctx->gen_smem_elems;
  g.normalize = normalize;
  g.mode = ctx->mode;
  return g;
}

template <typename K>
static int set_smem(bpx_ctx* ctx, K kernel, int bytes) {
  if (bytes > 48 * 1024) BPX_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return BPX_OK;
}

int bpx::launch_generic_update(bpx_ctx* ctx, const void* msg_in, void* msg_out, const int32_t* d_work, int64_t n_work,
                               int normalize, unsigned long long* resmax) {
  if (n_work == 0) return BPX_OK;
  GenericArgs g = generic_args(ctx, msg_in, msg_out, d_work, n_work, normalize);
  g.resmax = resmax;
  g.stop_key = resmax ? ctx->stop_key : 0ull;
  const int grid = (int)std::min<int64_t>(n_work, ctx->gen_grid);
  int rc;
  if (ctx->dtype == BPX_F64) {
    if ((rc = set_smem(ctx, bp_update_generic<double>, ctx->gen_smem_bytes))) return rc;
    bp_update_generic<double><<<grid, 256, ctx->gen_smem_bytes, ctx->stream>>>(g);
  } else {
    if ((rc = set_smem(ctx, bp_update_generic<c64>, ctx->gen_smem_bytes))) return rc;
    bp_update_generic<c64><<<grid, 256, ctx->gen_smem_bytes, ctx->stream>>>(g);
  }
  ctx->n_launches++;
  BPX_CUDA(ctx, cudaGetLastError());
  return BPX_OK;
}

// ---- per-sweep residual keys --------------------------------------------------------------------------------
// Sweep number h (since the ring was last cleared) records into slot h % history_cap; kernels fold their per-edge
// terms with atomicMax (residual_record), so a sweep needs no reduction kernel unless generic buckets took part.
static int residual_ring_clear(bpx_ctx* ctx) {
  const size_t bytes = ((size_t)ctx->history_cap + 1) * sizeof(unsigned long long);
  BPX_CUDA(ctx, cudaMemsetAsync(ctx->d_reskeys, 0, bytes, ctx->stream));
  BPX_CUDA(ctx, cudaMemsetAsync(ctx->d_reskeys_local, 0, bytes, ctx->stream));
  ctx->history_len = 0;
  ctx->ring_dirty = false;
  return BPX_OK;
}

static int residual_begin_sweep(bpx_ctx* ctx) {
  if (ctx->history_len >= ctx->history_cap || ctx->ring_dirty) {  // ring full (or slots re-used by streamed steps): start over
    int rc = halo_gate(ctx);
    if (rc) return rc;
    if ((rc = residual_ring_clear(ctx))) return rc;
  }
  ctx->cur_slot = (ctx->nranks > 1 ? ctx->d_reskeys_local : ctx->d_reskeys) + ctx->history_len;
  if (ctx->slot_override) ctx->cur_slot = ctx->slot_override;  // streamed steps record into a private, self-resetting slot
  return BPX_OK;
}

static int launch_residual_max(bpx_ctx* ctx, const int32_t* list, int64_t n, unsigned long long* slot) {
  if (n <= 0) return BPX_OK;
  bp_residual_max<<<1, 1024, 0, ctx->stream>>>(ctx->d_residual, list, n, slot);
  ctx->n_launches++;
  BPX_CUDA(ctx, cudaGetLastError());
  return BPX_OK;
}

// residual of recorded sweep `idx` (host value; synchronises)
static int residual_read(bpx_ctx* ctx, int idx, double* out) {
  unsigned long long key = 0;
  BPX_CUDA(ctx, cudaMemcpyAsync(&key, ctx->d_reskeys + idx, sizeof(key), cudaMemcpyDeviceToHost, ctx->stream));
  BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out = residual_from_key(key);
  return BPX_OK;
}

// one synchronous sweep: every owned directed edge, bucket by bucket, from d_msg[cur] into d_msg[cur^1]
struct NvtxScope {  // RAII range: every return path of the function closes it
  explicit NvtxScope(const char* name) { nvtxRangePushA(name); }
  ~NvtxScope() { nvtxRangePop(); }
};

static int sweep_once(bpx_ctx* ctx, int normalize) {
  NvtxScope nvtx_sweep("bpx:sweep");
  const void* in = ctx->d_msg[ctx->cur];
  void* out = ctx->d_msg[ctx->cur ^ 1];
  int rc;
  // Partitioned runs.  If every launch of this rank's sweep is a specialised kernel, the exchange is fused into them:
  // the FIRST launch waits for the peers' posts (prologue), every launch stores its cut-edge messages into the peers
  // from its epilogue, and the last CTA of the LAST launch posts: no gate / halo / post launches.  Otherwise the
  // exchange runs as separate small kernels.
  int n_fast = 0, n_generic = 0;
  std::vector<std::pair<int64_t, int>> order;  // (-edges of the launch group, leader bucket): largest launch first, so
  //                                              that a streamed upload overlaps with it; generic launch last
  for (int bi = 0; bi < (int)ctx->buckets.size(); ++bi) {
    const Bucket& b = ctx->buckets[bi];
    if (b.my_edges.empty()) continue;
    if (plain_family(b.kernel)) ++n_generic;
    else if (b.leader == bi) ++n_fast;
    if (b.leader == bi) {
      int64_t edges = 0;
      for (const Bucket& o : ctx->buckets)
        if (o.leader == bi && o.kernel == b.kernel) edges += (int64_t)o.my_edges.size();
      order.emplace_back(b.kernel == BPX_KERNEL_GENERIC ? 1 : -edges, bi);
    }
  }
  std::stable_sort(order.begin(), order.end(), [](const std::pair<int64_t, int>& a, const std::pair<int64_t, int>& b) { return a.first < b.first; });
  const bool fused = ctx->nranks > 1 && ctx->halo_connected && n_fast >= 1 && n_generic == 0 && !ctx->sites_dirty;
  ctx->peer_args = PeerArgs{};
  if (!fused && (rc = halo_gate(ctx))) return rc;
  if ((rc = fast_refresh_sites(ctx))) return rc;
  if ((rc = residual_begin_sweep(ctx))) return rc;
  if (fused) {
    PeerArgs& pa = ctx->peer_args;
    pa.nranks = ctx->nranks;
    pa.rank = ctx->rank;
    pa.my_mailbox = reinterpret_cast<Mailbox*>(ctx->d_mailbox);
    pa.peer_mailbox = reinterpret_cast<Mailbox* const*>(ctx->d_peer_mailbox);
    pa.wait_id = 0;                 // set per launch below: the first launch gates, the last one posts
    pa.post_id = 0;
    pa.wait_mask = ctx->recv_mask;  // only the ranks that feed this rank are awaited inside the kernel;
    pa.prev_global_key = nullptr;   // the global residual is folded by the explicit gate when the host asks for it
    pa.local_key = ctx->cur_slot;
    pa.peer_out = reinterpret_cast<double* const*>(ctx->d_peer_msg) + (size_t)(ctx->cur ^ 1) * ctx->nranks;
    pa.ticket = ctx->d_ticket;
    pa.error_flag = ctx->d_halo_error;
  }
  const unsigned long long fused_wait = fused && ctx->gate_pending ? ctx->sweep_id : 0;
  const unsigned long long fused_post = fused ? ++ctx->sweep_id : 0;
  unsigned int* const io_ticket = ctx->io_args.ticket;
  for (size_t li = 0; li < order.size(); ++li) {
    const int bi = order[li].second;
    Bucket& b = ctx->buckets[bi];  // (buckets merged into a launch group run in their leader's launch)
    if (fused) {
      ctx->peer_args.wait_id = li == 0 ? fused_wait : 0;
      ctx->peer_args.post_id = li + 1 == order.size() ? fused_post : 0;
    }
    ctx->io_args.ticket = li + 1 == order.size() ? io_ticket : nullptr;  // streamed host I/O: the last launch finishes the step
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->profiling && b.timing.size() < 8192) {
      BPX_CUDA(ctx, cudaEventCreate(&ev0));
      BPX_CUDA(ctx, cudaEventCreate(&ev1));
      b.timing.emplace_back(ev0, ev1);
      BPX_CUDA(ctx, cudaEventRecord(ev0, ctx->stream));
    }
    {
      // NVTX range per bucket launch: "bpx:bucket z=<degree> chi=<dim> d=<phys> kernel=<family>"
      static const char* const fam[] = {"auto", "generic", "onchip", "sliced", "vertex"};
      char label[96];
      snprintf(label, sizeof(label), "bpx:bucket z=%d chi=%d d=%d kernel=%s edges=%lld", b.z, b.chi, b.d, fam[b.kernel & 7 ? (b.kernel <= 4 ? b.kernel : 0) : 0],
               (long long)b.my_edges.size());
      nvtxRangePushA(label);
    }
    if (b.kernel == BPX_KERNEL_GENERIC)  // ONE launch for all generic buckets (the kernel takes any mix of edges)
      rc = launch_generic_update(ctx, in, out, ctx->d_generic_edges, ctx->n_generic_edges, normalize, ctx->cur_slot);
    else if (b.kernel == BPX_KERNEL_VERTEX)
      rc = launch_vertex_update(ctx, b, in, out, normalize);
    else
      rc = launch_fast_update(ctx, b, in, out, normalize);
    nvtxRangePop();
    if (rc) return rc;
    if (ev1) BPX_CUDA(ctx, cudaEventRecord(ev1, ctx->stream));
  }
  if (fused) {
    ctx->gate_pending = true;  // the NEXT launch (or an explicit halo_gate) waits for this sweep's posts
    ctx->gate_hist_idx = ctx->history_len;
    ctx->peer_args = PeerArgs{};
  } else if (ctx->nranks > 1) {
    if ((rc = halo_push(ctx, out))) return rc;
    // local key -> mailbox of every rank; the gate folds them (into d_reskeys[slot]) before the next sweep
    ctx->gate_hist_idx = ctx->history_len;
    if ((rc = halo_post_residual(ctx))) return rc;
  }
  ctx->history_len++;
  ctx->cur ^= 1;
  ctx->n_updates += ctx->n_owned_edges;
  ctx->n_sweeps++;
  return BPX_OK;
}

extern "C" int bpx_sweep(bpx_ctx* ctx, int max_sweeps, double tol, int normalize, double* residual_out, int* sweeps_done) {
  MULTI(ctx, bpx::multi::sweep(ctx, max_sweeps, tol, normalize, residual_out, sweeps_done));
  NEED_DIMS(ctx, "bpx_sweep");
  REQUIRE(ctx, max_sweeps >= 0, "bpx_sweep: max_sweeps < 0");
  int done = 0, rc;
  double res = INFINITY;
  if ((rc = halo_gate(ctx))) return rc;
  if ((rc = residual_ring_clear(ctx))) return rc;
  if (tol > 0.0 && ctx->nranks == 1 && !getenv("BPX_HOST_STOP")) {
    // StopWhenConverged ON THE DEVICE (AlgorithmsInterfaceExtensions.jl:84-119): sweeps are enqueued in batches ahead of
    // the host; every launch of a sweep first looks at the previous sweep's final residual key and turns into a no-op
    // once it is below the tolerance (sweep_already_converged), so the iterate is exactly the one after the first
    // converged sweep however many sweeps were enqueued.  The host synchronises once per BATCH (4, 8, 16, 16, ...) to
    // read the keys and stop enqueuing -- not once per sweep.
    const unsigned long long tol_key = residual_key(tol);
    const int cur0 = ctx->cur;
    const int64_t sweeps0 = ctx->n_sweeps, updates0 = ctx->n_updates;
    int batch = 4;
    bool converged = false;
    std::vector<unsigned long long> keys;
    while (done < max_sweeps && !converged) {
      if (ctx->history_len >= ctx->history_cap && (rc = residual_ring_clear(ctx))) return rc;  // (host is in sync here)
      const int first = ctx->history_len;
      const int nb = std::min(std::min(batch, max_sweeps - done), ctx->history_cap - first);
      for (int b = 0; b < nb; ++b) {
        ctx->stop_key = ctx->history_len > first || first > 0 ? tol_key : 0ull;  // slot[-1] exists and belongs to this call
        rc = sweep_once(ctx, normalize);
        ctx->stop_key = 0ull;
        if (rc) return rc;
      }
      keys.resize(nb);
      BPX_CUDA(ctx, cudaMemcpyAsync(keys.data(), ctx->d_reskeys + first, nb * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
      BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      int ran = nb;
      for (int b = 0; b < nb; ++b) {
        res = residual_from_key(keys[b]);
        if (res < tol) {
          converged = true;
          ran = b + 1;  // sweeps b+1 .. nb-1 of the batch were no-ops
          break;
        }
      }
      done += ran;
      ctx->history_len = first + ran;
      batch = std::min(16, batch * 2);
    }
    // host-side bookkeeping of the sweeps that really ran
    ctx->cur = cur0 ^ (done & 1);
    ctx->n_sweeps = sweeps0 + done;
    ctx->n_updates = updates0 + (int64_t)done * ctx->n_owned_edges;
    if (residual_out) *residual_out = res;
    if (sweeps_done) *sweeps_done = done;
    return BPX_OK;
  }
  for (int it = 0; it < max_sweeps; ++it) {
    if ((rc = sweep_once(ctx, normalize))) return rc;
    ++done;
    if (tol > 0.0) {
      // StopWhenConverged: stop after the first sweep whose (global) residual is below tol; partitioned contexts need
      // every rank's post of the sweep, so the host checks after each one
      if ((rc = halo_gate(ctx))) return rc;
      if ((rc = residual_read(ctx, ctx->history_len - 1, &res))) return rc;
      if (res < tol) break;
    }
  }
  if ((rc = halo_gate(ctx))) return rc;
  if (done > 0 && !(tol > 0.0)) {
    if ((rc = residual_read(ctx, ctx->history_len - 1, &res))) return rc;
  }
  BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->d_halo_error) {
    int flag = 0;
    BPX_CUDA(ctx, cudaMemcpy(&flag, ctx->d_halo_error, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) {
      set_error(ctx, "bpx_sweep: timed out waiting for a peer rank's sweep post");
      return BPX_ERR_CUDA;
    }
  }
  if (residual_out) *residual_out = res;
  if (sweeps_done) *sweeps_done = done;
  return BPX_OK;
}

extern "C" int bpx_sweep_async(bpx_ctx* ctx, int n_sweeps, int normalize) {
  MULTI(ctx, bpx::multi::each(ctx, [&](bpx_ctx* c) -> int { return bpx_sweep_async(c, n_sweeps, normalize); }));
  NEED_DIMS(ctx, "bpx_sweep_async");
  REQUIRE(ctx, n_sweeps >= 0, "bpx_sweep_async: n_sweeps < 0");
  for (int it = 0; it < n_sweeps; ++it) {
    int rc = sweep_once(ctx, normalize);
    if (rc) return rc;
  }
  return BPX_OK;
}

extern "C" int bpx_set_profiling(bpx_ctx* ctx, int enable) {
  MULTI(ctx, bpx::multi::each(ctx, [&](bpx_ctx* c) -> int { return bpx_set_profiling(c, enable); }));
  NEED_DIMS(ctx, "bpx_set_profiling");
  BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto& b : ctx->buckets) {
    for (auto& ev : b.timing) {
      cudaEventDestroy(ev.first);
      cudaEventDestroy(ev.second);
    }
    b.timing.clear();
    b.timed_ms = 0.0;
    b.timed_launches = 0;
  }
  ctx->profiling = enable != 0;
  return BPX_OK;
}

extern "C" int bpx_bucket_time(bpx_ctx* ctx, int bucket, double* total_ms, int64_t* launches) {
  MULTI(ctx, bpx::multi::fail(ctx, MULTI0(ctx), bpx_bucket_time(MULTI0(ctx), bucket, total_ms, launches)));
  NEED_DIMS(ctx, "bpx_bucket_time");
  REQUIRE(ctx, bucket >= 0 && bucket < (int)ctx->buckets.size(), "bpx_bucket_time: bucket out of range");
  BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  Bucket& b = ctx->buckets[bucket];
  for (auto& ev : b.timing) {
    float ms = 0.f;
    BPX_CUDA(ctx, cudaEventElapsedTime(&ms, ev.first, ev.second));
    b.timed_ms += ms;
    b.timed_launches++;
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  b.timing.clear();
  if (total_ms) *total_ms = b.timed_ms;
  if (launches) *launches = b.timed_launches;
  return BPX_OK;
}

extern "C" int bpx_sweep_sequence(bpx_ctx* ctx, const int64_t* edge_seq, int64_t n_seq, int max_sweeps, double tol,
                                  int normalize, double* residual_out, int* sweeps_done) {
  MULTI(ctx, bpx::multi::fail(ctx, MULTI0(ctx), bpx_sweep_sequence(MULTI0(ctx), edge_seq, n_seq, max_sweeps, tol, normalize, residual_out, sweeps_done)));
  NEED_DIMS(ctx, "bpx_sweep_sequence");
  REQUIRE(ctx, ctx->nranks == 1, "bpx_sweep_sequence: the sequential schedule is single-GPU only");
  REQUIRE(ctx, max_sweeps >= 0 && n_seq >= 0 && (n_seq == 0 || edge_seq), "bpx_sweep_sequence: bad arguments");
  const int64_t ne = ctx->ne;
  // split the sequence into maximal runs of consecutive updates whose write set is disjoint from the
  // run's read set: inside a run the in-place updates are independent and go out as one launch
  std::vector<int32_t> flat;
  std::vector<int64_t> batch_ptr(1, 0);
  {
    std::vector<int> r_stamp(ne, -1), w_stamp(ne, -1);
    int batch = 0;
    for (int64_t i = 0; i < n_seq; ++i) {
      const int64_t e = edge_seq[i];
      REQUIRE(ctx, e >= 0 && e < ne, "bpx_sweep_sequence: edge_seq[%lld] out of range", (long long)i);
      const int32_t u = ctx->src[e];
      bool conflict = (r_stamp[e] == batch) || (w_stamp[e] == batch);
      for (int32_t f : ctx->out_edge[u])
        if (f != e && w_stamp[ctx->rev[f]] == batch) conflict = true;
      if (conflict) {
        batch_ptr.push_back((int64_t)flat.size());
        ++batch;
      }
      w_stamp[e] = batch;
      for (int32_t f : ctx->out_edge[u])
        if (f != e) r_stamp[ctx->rev[f]] = batch;
      flat.push_back((int32_t)e);
    }
    batch_ptr.push_back((int64_t)flat.size());
  }
  int32_t* d_flat = nullptr;
  int rc = upload(ctx, &d_flat, flat);
  if (rc) return rc;
  const size_t msg_bytes = (size_t)ctx->msg_off[ne] * ctx->esize;
  if (!ctx->d_msg_snapshot && (rc = dev_alloc(ctx, (char**)&ctx->d_msg_snapshot, msg_bytes))) {
    cudaFree(d_flat);
    return rc;
  }
  int done = 0;
  double res = INFINITY;
  rc = residual_ring_clear(ctx);
  void* m = ctx->d_msg[ctx->cur];
  for (int it = 0; it < max_sweeps && rc == BPX_OK; ++it) {
    if (ctx->history_len >= ctx->history_cap && (rc = residual_ring_clear(ctx))) break;
    cudaMemcpyAsync(ctx->d_msg_snapshot, m, msg_bytes, cudaMemcpyDeviceToDevice, ctx->stream);
    for (size_t b = 0; b + 1 < batch_ptr.size() && rc == BPX_OK; ++b) {
      const int64_t n = batch_ptr[b + 1] - batch_ptr[b];
      // edges of one run may belong to different buckets: the generic kernel takes any mix
      rc = launch_generic_update(ctx, m, m, d_flat + batch_ptr[b], n, normalize, nullptr);
    }
    if (rc) break;
    // iterate_diff against the previous sweep over ALL edges (beliefpropagation.jl:261-267)
    if (ne > 0) {
      const int threads = 256;
      const int64_t blocks = (ne * 32 + threads - 1) / threads;
      if (ctx->dtype == BPX_F64)
        bp_edge_residual<double><<<(unsigned)blocks, threads, 0, ctx->stream>>>((const double*)ctx->d_msg_snapshot, (const double*)m,
                                                                               ctx->d_msg_off, ne, ctx->d_residual);
      else
        bp_edge_residual<c64><<<(unsigned)blocks, threads, 0, ctx->stream>>>((const c64*)ctx->d_msg_snapshot, (const c64*)m,
                                                                            ctx->d_msg_off, ne, ctx->d_residual);
      ctx->n_launches++;
    }
    rc = launch_residual_max(ctx, nullptr, ne, ctx->d_reskeys + ctx->history_len);
    if (rc) break;
    ++done;
    ctx->n_updates += n_seq;
    ctx->n_sweeps++;
    ctx->history_len++;
    if (tol > 0.0) {
      if ((rc = residual_read(ctx, ctx->history_len - 1, &res))) break;
      if (res < tol) break;
    }
  }
  if (rc == BPX_OK && done > 0 && !(tol > 0.0)) rc = residual_read(ctx, ctx->history_len - 1, &res);
  cudaError_t ce = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_flat);
  if (rc) return rc;
  BPX_CUDA(ctx, ce);
  // keep both sets identical so a later synchronous sweep starts from the same iterate
  BPX_CUDA(ctx, cudaMemcpyAsync(ctx->d_msg[ctx->cur ^ 1], m, msg_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (residual_out) *residual_out = res;
  if (sweeps_done) *sweeps_done = done;
  return BPX_OK;
}

extern "C" int bpx_residual_history(bpx_ctx* ctx, double* out, int n, int* n_out) {
  MULTI(ctx, bpx::multi::fail(ctx, MULTI0(ctx), bpx_residual_history(MULTI0(ctx), out, n, n_out)));
  NEED_DIMS(ctx, "bpx_residual_history");
  { int rc_ = halo_gate(ctx); if (rc_) return rc_; }
  const int k = std::max(0, std::min(n, ctx->history_len));
  if (k > 0) {
    REQUIRE(ctx, out, "bpx_residual_history: out is NULL");
    std::vector<unsigned long long> keys(k);
    BPX_CUDA(ctx, cudaMemcpyAsync(keys.data(), ctx->d_reskeys, k * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < k; ++i) out[i] = residual_from_key(keys[i]);
  }
  if (n_out) *n_out = k;
  return BPX_OK;
}

extern "C" int bpx_last_residual(bpx_ctx* ctx, double* out) {
  MULTI(ctx, bpx::multi::fail(ctx, MULTI0(ctx), bpx_last_residual(MULTI0(ctx), out)));
  NEED_DIMS(ctx, "bpx_last_residual");
  { int rc_ = halo_gate(ctx); if (rc_) return rc_; }
  REQUIRE(ctx, out, "bpx_last_residual: out is NULL");
  if (ctx->history_len == 0) {
    *out = INFINITY;
    return BPX_OK;
  }
  return residual_read(ctx, ctx->history_len - 1, out);
}

extern "C" int bpx_iterate_diff(bpx_ctx* ctx, const void* other_packed, double* out) {
  PAD(ctx, bpx::pad::iterate_diff(ctx, other_packed, out));
  MULTI(ctx, bpx::multi::fail(ctx, MULTI0(ctx), bpx_iterate_diff(MULTI0(ctx), other_packed, out)));
  NEED_DIMS(ctx, "bpx_iterate_diff");
  { int rc_ = halo_gate(ctx); if (rc_) return rc_; }
  REQUIRE(ctx, out && (other_packed || ctx->ne == 0), "bpx_iterate_diff: bad arguments");
  const int64_t ne = ctx->ne;
  if (ne == 0) {
    *out = -INFINITY;
    return BPX_OK;
  }
  const size_t msg_bytes = (size_t)ctx->msg_off[ne] * ctx->esize;
  int rc;
  if (!ctx->d_msg_snapshot && (rc = dev_alloc(ctx, (char**)&ctx->d_msg_snapshot, msg_bytes))) return rc;
  BPX_CUDA(ctx, cudaMemcpyAsync(ctx->d_msg_snapshot, other_packed, msg_bytes, cudaMemcpyHostToDevice, ctx->stream));
  const int threads = 256;
  const int64_t blocks = (ne * 32 + threads - 1) / threads;
  if (ctx->dtype == BPX_F64)
    bp_edge_residual<double><<<(unsigned)blocks, threads, 0, ctx->stream>>>((const double*)ctx->d_msg[ctx->cur],
                                                                           (const double*)ctx->d_msg_snapshot, ctx->d_msg_off, ne,
                                                                           ctx->d_residual);
  else
    bp_edge_residual<c64><<<(unsigned)blocks, threads, 0, ctx->stream>>>((const c64*)ctx->d_msg[ctx->cur],
                                                                        (const c64*)ctx->d_msg_snapshot, ctx->d_msg_off, ne,
                                                                        ctx->d_residual);
  ctx->n_launches++;
  BPX_CUDA(ctx, cudaGetLastError());
  BPX_CUDA(ctx, cudaMemsetAsync(ctx->d_reskeys + ctx->history_cap, 0, sizeof(unsigned long long), ctx->stream));
  if ((rc = launch_residual_max(ctx, nullptr, ne, ctx->d_reskeys + ctx->history_cap))) return rc;
  return residual_read(ctx, ctx->history_cap, out);
}

// ---- beliefs --------------------------------------------------------------------------------------
// vertex_scalar(s) (messagecache.jl:139-151) on the UPDATE kernels: one sweep of every bucket's kernel with normalize = 0
// from the current message set into the idle one -- no exchange, no residual, no flip -- and one dot product per vertex
// (bp_vertex_belief).  Every bucket family, single-layer networks included, serves beliefs at sweep speed this way; only
// isolated vertices (no out-edge) are left to the generic scalar kernel.
static int belief_sweep(bpx_ctx* ctx, char* d_out /* nv elements, zero-initialised */) {
  NvtxScope nvtx_beliefs("bpx:beliefs");
  int rc;
  if ((rc = fast_refresh_sites(ctx))) return rc;
  const void* in = ctx->d_msg[ctx->cur];
  void* tmp = ctx->d_msg[ctx->cur ^ 1];
  const PeerArgs peer0 = ctx->peer_args;
  const HostIO io0 = ctx->io_args;
  unsigned long long* const slot0 = ctx->cur_slot;
  const unsigned long long stop0 = ctx->stop_key;
  ctx->peer_args = PeerArgs{};
  ctx->io_args = HostIO{};
  ctx->stop_key = 0ull;
  ctx->cur_slot = ctx->d_reskeys + ctx->history_cap;  // spare entry behind the ring: the residual of this pass is discarded
  rc = BPX_OK;
  for (int bi = 0; bi < (int)ctx->buckets.size() && rc == BPX_OK; ++bi) {
    Bucket& b = ctx->buckets[bi];
    if (b.my_edges.empty() || b.leader != bi) continue;
    if (b.kernel == BPX_KERNEL_GENERIC) rc = launch_generic_update(ctx, in, tmp, ctx->d_generic_edges, ctx->n_generic_edges, 0, ctx->cur_slot);
    else if (b.kernel == BPX_KERNEL_VERTEX) rc = launch_vertex_update(ctx, b, in, tmp, 0);
    else rc = launch_fast_update(ctx, b, in, tmp, 0);
  }
  ctx->peer_args = peer0;
  ctx->io_args = io0;
  ctx->cur_slot = slot0;
  ctx->stop_key = stop0;
  if (rc) return rc;
  if (!ctx->d_first_out_edge) {
    std::vector<int32_t> first(ctx->nv, -1);
    for (int64_t v = 0; v < ctx->nv; ++v)
      if (!ctx->out_edge[v].empty()) first[v] = ctx->out_edge[v][0];
    if ((rc = upload(ctx, &ctx->d_first_out_edge, first))) return rc;
  }
  const int64_t n_work = ctx->nranks > 1 ? ctx->n_owned_vertices : ctx->nv;
  if (n_work > 0) {
    const int threads = 256;
    const int64_t blocks = (n_work * 32 + threads - 1) / threads;
    const int32_t* list = ctx->nranks > 1 ? ctx->d_owned_vertices : nullptr;
    if (ctx->dtype == BPX_F64)
      bp_vertex_belief<double><<<(unsigned)blocks, threads, 0, ctx->stream>>>((const double*)tmp, (const double*)in, ctx->d_msg_off, ctx->d_rev,
                                                                             ctx->d_first_out_edge, list, n_work, (double*)d_out);
    else
      bp_vertex_belief<c64><<<(unsigned)blocks, threads, 0, ctx->stream>>>((const c64*)tmp, (const c64*)in, ctx->d_msg_off, ctx->d_rev,
                                                                          ctx->d_first_out_edge, list, n_work, (c64*)d_out);
    ctx->n_launches++;
    BPX_CUDA(ctx, cudaGetLastError());
  }
  return BPX_OK;
}

// Enqueues the vertex scalars (or expectation-value numerators) into the context's WS_SCALARS buffer; nothing leaves the device.
static int vertex_scalars_enqueue(bpx_ctx* ctx, const void* ops_packed, char** d_out_p) {
  const int64_t nv = ctx->nv;
  char* d_out = nullptr;
  char* d_ops = nullptr;
  int64_t* d_op_off = nullptr;
  int rc = ws_get(ctx, bpx_ctx::WS_SCALARS, (size_t)nv * ctx->esize, &d_out);
  if (rc) return rc;
  *d_out_p = d_out;
  if (ops_packed) {
    std::vector<int64_t> op_off(nv + 1, 0);
    for (int64_t v = 0; v < nv; ++v) op_off[v + 1] = op_off[v] + (int64_t)ctx->phys_dim[v] * ctx->phys_dim[v];
    if ((rc = ws_upload(ctx, bpx_ctx::WS_OP_OFF, op_off, &d_op_off)) || (rc = ws_get(ctx, bpx_ctx::WS_OPS, (size_t)op_off[nv] * ctx->esize, &d_ops)))
      return rc;
    BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // op_off is a local
    BPX_CUDA(ctx, cudaMemcpyAsync(d_ops, ops_packed, (size_t)op_off[nv] * ctx->esize, cudaMemcpyHostToDevice, ctx->stream));
  }
  // partitioned contexts hold (and know the in-messages of) their own vertices only: the others are reported as 0
  BPX_CUDA(ctx, cudaMemsetAsync(d_out, 0, (size_t)nv * ctx->esize, ctx->stream));
  int32_t* d_isolated = nullptr;
  int64_t n_isolated = 0;
  const bool on_update_kernels = !ops_packed && !getenv("BPX_BELIEFS_GENERIC");
  if (on_update_kernels) {
    // plain vertex scalars: at sweep speed on the buckets' own (tensor-pipe) kernels; the generic scalar kernel below then
    // only visits vertices without a link (their scalar is the plain norm of the tensor)
    if ((rc = belief_sweep(ctx, d_out))) return rc;
    std::vector<int32_t> iso;
    for (int64_t v = 0; v < nv; ++v)
      if (ctx->out_edge[v].empty() && (ctx->owner.empty() || ctx->owner[v] == ctx->rank)) iso.push_back((int32_t)v);
    n_isolated = (int64_t)iso.size();
    if (n_isolated > 0) {
      if ((rc = ws_upload(ctx, bpx_ctx::WS_LIST, iso, &d_isolated))) return rc;
      BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // iso is a local
    }
  }
  GenericArgs g = on_update_kernels
                      ? generic_args(ctx, ctx->d_msg[ctx->cur], nullptr, d_isolated, n_isolated, 0)
                      : generic_args(ctx, ctx->d_msg[ctx->cur], nullptr, ctx->nranks > 1 ? ctx->d_owned_vertices : nullptr,
                                     ctx->nranks > 1 ? ctx->n_owned_vertices : nv, 0);
  g.ops = d_ops;
  g.op_off = d_op_off;
  g.scalars_out = d_out;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(g.n_work, ctx->gen_grid));
  // the scalar kernel needs no output staging, but sharing the update kernel's geometry keeps it simple
  if (g.n_work > 0) {
    if (ctx->dtype == BPX_F64) {
      rc = set_smem(ctx, bp_vertex_scalar_generic<double>, ctx->gen_smem_bytes);
      if (!rc) bp_vertex_scalar_generic<double><<<grid, 256, ctx->gen_smem_bytes, ctx->stream>>>(g);
    } else {
      rc = set_smem(ctx, bp_vertex_scalar_generic<c64>, ctx->gen_smem_bytes);
      if (!rc) bp_vertex_scalar_generic<c64><<<grid, 256, ctx->gen_smem_bytes, ctx->stream>>>(g);
    }
    ctx->n_launches++;
    if (rc) return rc;
    BPX_CUDA(ctx, cudaGetLastError());
  }
  return BPX_OK;
}

static int vertex_scalars_impl(bpx_ctx* ctx, const void* ops_packed, void* out) {
  const int64_t nv = ctx->nv;
  if (nv == 0) return BPX_OK;
  REQUIRE(ctx, out, "vertex scalars: out is NULL");
  char* d_out = nullptr;
  int rc = vertex_scalars_enqueue(ctx, ops_packed, &d_out);
  if (rc) {
    cudaStreamSynchronize(ctx->stream);
    return rc;
  }
  const cudaError_t ce = cudaMemcpyAsync(out, d_out, (size_t)nv * ctx->esize, cudaMemcpyDeviceToHost, ctx->stream);
  const cudaError_t ce2 = cudaStreamSynchronize(ctx->stream);
  BPX_CUDA(ctx, ce);
  BPX_CUDA(ctx, ce2);
  return BPX_OK;
}

extern "C" int bpx_vertex_scalars(bpx_ctx* ctx, void* out) {
  MULTI(ctx, bpx::multi::merge_vertex(ctx, out, [&](bpx_ctx* c, void* o) -> int { return bpx_vertex_scalars(c, o); }));
  NEED_DIMS(ctx, "bpx_vertex_scalars");
  { int rc_ = halo_gate(ctx); if (rc_) return rc_; }
  return vertex_scalars_impl(ctx, nullptr, out);
}

extern "C" int bpx_vertex_expect_numerators(bpx_ctx* ctx, const void* ops_packed, void* out) {
  MULTI(ctx, bpx::multi::merge_vertex(ctx, out, [&](bpx_ctx* c, void* o) -> int { return bpx_vertex_expect_numerators(c, ops_packed, o); }));
  NEED_DIMS(ctx, "bpx_vertex_expect_numerators");
  { int rc_ = halo_gate(ctx); if (rc_) return rc_; }
  REQUIRE(ctx, ctx->mode == BPX_MODE_NORM, "bpx_vertex_expect_numerators: NORM mode only");
  REQUIRE(ctx, ops_packed || ctx->nv == 0, "bpx_vertex_expect_numerators: ops is NULL");
  return vertex_scalars_impl(ctx, ops_packed, out);
}

// edge scalars of all undirected edges into WS_EDGE_SCALARS (device only)
static int edge_scalars_enqueue(bpx_ctx* ctx, char** d_out_p) {
  const int64_t n = ctx->n_und;
  char* d_out = nullptr;
  int rc = ws_get(ctx, bpx_ctx::WS_EDGE_SCALARS, (size_t)n * ctx->esize, &d_out);
  if (rc) return rc;
  *d_out_p = d_out;
  if (n == 0) return BPX_OK;
  const int threads = 256;
  const int64_t blocks = (n * 32 + threads - 1) / threads;
  if (ctx->dtype == BPX_F64)
    bp_edge_scalar<double><<<(unsigned)blocks, threads, 0, ctx->stream>>>((const double*)ctx->d_msg[ctx->cur], ctx->d_msg_off,
                                                                         ctx->d_und_edge, ctx->d_rev, n, (double*)d_out);
  else
    bp_edge_scalar<c64><<<(unsigned)blocks, threads, 0, ctx->stream>>>((const c64*)ctx->d_msg[ctx->cur], ctx->d_msg_off,
                                                                      ctx->d_und_edge, ctx->d_rev, n, (c64*)d_out);
  ctx->n_launches++;
  BPX_CUDA(ctx, cudaGetLastError());
  return BPX_OK;
}

extern "C" int bpx_edge_scalars(bpx_ctx* ctx, void* out) {
  MULTI(ctx, bpx::multi::edge_scalars(ctx, out));
  NEED_DIMS(ctx, "bpx_edge_scalars");
  { int rc_ = halo_gate(ctx); if (rc_) return rc_; }
  const int64_t n = ctx->n_und;
  if (n == 0) return BPX_OK;
  REQUIRE(ctx, out, "bpx_edge_scalars: out is NULL");
  char* d_out = nullptr;
  int rc = edge_scalars_enqueue(ctx, &d_out);
  if (rc) {
    cudaStreamSynchronize(ctx->stream);
    return rc;
  }
  const cudaError_t ce = cudaMemcpyAsync(out, d_out, (size_t)n * ctx->esize, cudaMemcpyDeviceToHost, ctx->stream);
  const cudaError_t ce2 = cudaStreamSynchronize(ctx->stream);
  BPX_CUDA(ctx, ce);
  BPX_CUDA(ctx, ce2);
  return BPX_OK;
}

// ---- bethe_free_energy fully on the device (messagecache.jl:185-201) ---------------------------------------------------
// One CTA reduces log|t| and arg(t) over the vertex terms this rank owns and over the undirected edges whose first
// orientation starts at an owned vertex, in a fixed order (deterministic); 7 doubles come back:
//   [0] sum log|num|  [1] sum arg(num)  [2] sum log|den|  [3] sum arg(den)  [4] any real(num) < 0  [5] any real(den) < 0
//   [6] any den == 0
template <typename T>
__global__ void bp_bethe_logsum(const T* __restrict__ vs, const int32_t* __restrict__ vlist, int64_t n_v,
                                const T* __restrict__ es, const int32_t* __restrict__ und_edge, const int32_t* __restrict__ src,
                                const int32_t* __restrict__ owner, int rank, int64_t n_e, double* __restrict__ out) {
  using E = Elem<T>;
  __shared__ double red[7][32];
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = threadIdx.x; i < n_v; i += blockDim.x) {
    const T t = vs[vlist ? vlist[i] : i];
    const double re = E::real(t), im = E::imag(t);
    acc[0] += log(hypot(re, im));
    acc[1] += atan2(im, re);
    if (re < 0) acc[4] = 1.0;
  }
  for (int64_t i = threadIdx.x; i < n_e; i += blockDim.x) {
    if (owner && owner[src[und_edge[i]]] != rank) continue;
    const T t = es[i];
    const double re = E::real(t), im = E::imag(t);
    acc[2] += log(hypot(re, im));
    acc[3] += atan2(im, re);
    if (re < 0) acc[5] = 1.0;
    if (re == 0 && im == 0) acc[6] = 1.0;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = 0; k < 7; ++k) {
    double x = acc[k];
    for (int o = 16; o > 0; o >>= 1) {
      const double y = __shfl_down_sync(0xffffffffu, x, o);
      x = k < 4 ? x + y : fmax(x, y);
    }
    if (lane == 0) red[k][w] = x;
  }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    for (int k = 0; k < 7; ++k) {
      double x = lane < nw ? red[k][lane] : 0.0;
      for (int o = 16; o > 0; o >>= 1) {
        const double y = __shfl_down_sync(0xffffffffu, x, o);
        x = k < 4 ? x + y : fmax(x, y);
      }
      if (lane == 0) out[k] = x;
    }
  }
}

extern "C" int bpx_bethe_free_energy_parts(bpx_ctx* ctx, double parts[7]) {
  REQUIRE(ctx, parts, "bpx_bethe_free_energy_parts: parts is NULL");
  if (ctx && !ctx->children.empty()) {
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (bpx_ctx* c : ctx->children) {
      cudaSetDevice(c->device);
      double p[7];
      const int rc = bpx_bethe_free_energy_parts(c, p);
      if (rc) return bpx::multi::fail(ctx, c, rc);
      for (int k = 0; k < 4; ++k) acc[k] += p[k];
      for (int k = 4; k < 7; ++k) acc[k] = std::max(acc[k], p[k]);
    }
    memcpy(parts, acc, sizeof(acc));
    return BPX_OK;
  }
  NEED_DIMS(ctx, "bpx_bethe_free_energy");
  { int rc_ = halo_gate(ctx); if (rc_) return rc_; }
  NvtxScope nvtx("bpx:bethe_free_energy");
  char *d_vs = nullptr, *d_es = nullptr;
  double* d_out = nullptr;
  int rc = vertex_scalars_enqueue(ctx, nullptr, &d_vs);
  if (!rc) rc = edge_scalars_enqueue(ctx, &d_es);
  if (!rc) rc = ws_get(ctx, bpx_ctx::WS_LOGSUM, 8 * sizeof(double), &d_out);
  if (!rc && !ctx->h_logsum && cudaMallocHost(&ctx->h_logsum, 8 * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    set_error(ctx, "bpx_bethe_free_energy: cudaMallocHost failed");
    rc = BPX_ERR_ALLOC;
  }
  if (rc) {
    cudaStreamSynchronize(ctx->stream);
    return rc;
  }
  const bool part = ctx->nranks > 1;
  int32_t* d_owner = nullptr;
  if (part) {
    if ((rc = ws_upload(ctx, bpx_ctx::WS_OUT, ctx->owner, &d_owner))) return rc;
  }
  const int64_t n_v = part ? ctx->n_owned_vertices : ctx->nv;
  const int32_t* vlist = part ? ctx->d_owned_vertices : nullptr;
  if (ctx->dtype == BPX_F64)
    bp_bethe_logsum<double><<<1, 1024, 0, ctx->stream>>>((const double*)d_vs, vlist, n_v, (const double*)d_es, ctx->d_und_edge, ctx->d_src,
                                                         d_owner, ctx->rank, ctx->n_und, d_out);
  else
    bp_bethe_logsum<c64><<<1, 1024, 0, ctx->stream>>>((const c64*)d_vs, vlist, n_v, (const c64*)d_es, ctx->d_und_edge, ctx->d_src,
                                                      d_owner, ctx->rank, ctx->n_und, d_out);
  ctx->n_launches++;
  cudaError_t ce = cudaGetLastError();
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(ctx->h_logsum, d_out, 7 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  const cudaError_t ce2 = cudaStreamSynchronize(ctx->stream);
  BPX_CUDA(ctx, ce);
  BPX_CUDA(ctx, ce2);
  memcpy(parts, ctx->h_logsum, 7 * sizeof(double));
  return BPX_OK;
}

extern "C" int bpx_bethe_free_energy(bpx_ctx* ctx, double out[2], int* promoted) {
  REQUIRE(ctx, out, "bpx_bethe_free_energy: out is NULL");
  double p[7];
  const int rc = bpx_bethe_free_energy_parts(ctx, p);
  if (rc) return rc;
  const bpx_ctx* c = ctx->children.empty() ? ctx : ctx->children[0];
  // messagecache.jl:189-194: real terms are promoted to complex only if one of them is negative; complex terms always
  // carry their phase.  log(complex(x)) = log|x| + i arg(x)
  const bool cplx = c->dtype != BPX_F64;
  const bool pn = cplx || p[4] != 0.0, pd = cplx || p[5] != 0.0;
  if (promoted) *promoted = (pn || pd) ? 1 : 0;
  if (p[6] != 0.0) {  // :196-198
    out[0] = -INFINITY;
    out[1] = 0.0;
    return BPX_OK;
  }
  out[0] = p[0] - p[2];
  out[1] = (pn ? p[1] : 0.0) - (pd ? p[3] : 0.0);
  return BPX_OK;
}

// ---- introspection ------------------------------------------------------------------------------
extern "C" int bpx_num_buckets(const bpx_ctx* ctx) {
  if (ctx && !ctx->children.empty()) return bpx_num_buckets(ctx->children[0]);
  return (ctx && ctx->dims_set) ? (int)ctx->buckets.size() : -1;
}

extern "C" int bpx_bucket_info(const bpx_ctx* ctx, int bucket, int64_t info[8]) {
  if (ctx && !ctx->children.empty()) {  // a child's view, with the vertex / edge counts summed over the devices
    int rc = bpx_bucket_info(ctx->children[0], bucket, info);
    for (size_t k = 1; k < ctx->children.size() && rc == BPX_OK; ++k) {
      int64_t t[8];
      rc = bpx_bucket_info(ctx->children[k], bucket, t);
      info[3] += t[3];
      info[4] += t[4];
    }
    return rc;
  }
  if (!ctx || !ctx->dims_set || bucket < 0 || bucket >= (int)ctx->buckets.size() || !info) return BPX_ERR_INVALID;
  const Bucket& b = ctx->buckets[bucket];
  info[0] = b.z;
  info[1] = b.chi;
  info[2] = b.d;
  info[3] = (int64_t)b.my_vertices.size();
  info[4] = (int64_t)b.my_edges.size();
  info[5] = b.kernel;
  info[6] = b.leader;
  info[7] = 0;
  return BPX_OK;
}

extern "C" int bpx_set_kernel_policy(bpx_ctx* ctx, int kernel) {
  MULTI(ctx, bpx::multi::each(ctx, [&](bpx_ctx* c) -> int { return bpx_set_kernel_policy(c, kernel); }));
  if (!ctx) return BPX_ERR_INVALID;
  REQUIRE(ctx, kernel >= BPX_KERNEL_AUTO && kernel <= BPX_KERNEL_VERTEX, "bpx_set_kernel_policy: unknown kernel %d", kernel);
  ctx->kernel_policy = kernel;
  if (ctx->dims_set) {
    BPX_CUDA(ctx, cudaSetDevice(ctx->device));
    return rebuild_work_lists(ctx);  // re-picks every bucket's kernel, rebuilds launch groups and edge lists
  }
  return BPX_OK;
}

extern "C" int bpx_counters(bpx_ctx* ctx, int64_t out[3], int reset) {
  MULTI(ctx, bpx::multi::counters(ctx, out, reset));
  if (!ctx) return BPX_ERR_INVALID;
  if (out) {
    out[0] = ctx->n_launches;
    out[1] = ctx->n_updates;
    out[2] = ctx->n_sweeps;
  }
  if (reset) ctx->n_launches = ctx->n_updates = ctx->n_sweeps = 0;
  return BPX_OK;
}

extern "C" int bpx_set_stream(bpx_ctx* ctx, void* cuda_stream) {
  MULTI(ctx, bpx::multi::set_stream(ctx, cuda_stream));
  if (!ctx) return BPX_ERR_INVALID;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return BPX_OK;
}

// ---- the consumer of the messages: BP simple-update gate application (bpx_apply.cuh) ---------------------------
extern "C" int bpx_get_site_tensor(bpx_ctx* ctx, int64_t v, void* data) {
  PAD(ctx, bpx::pad::get_site_tensor(ctx, v, data));
  MULTI(ctx, (v < 0 || v >= ctx->nv) ? (int)BPX_ERR_INVALID : bpx::multi::fail(ctx, bpx::multi::owner_of(ctx, v), bpx_get_site_tensor(bpx::multi::owner_of(ctx, v), v, data)));
  NEED_DIMS(ctx, "bpx_get_site_tensor");
  REQUIRE(ctx, v >= 0 && v < ctx->nv && data, "bpx_get_site_tensor: bad arguments");
  REQUIRE(ctx, ctx->dev_site_off[v] >= 0, "bpx_get_site_tensor: vertex %lld is not resident on this rank", (long long)v);
  const size_t off = (size_t)ctx->dev_site_off[v] * ctx->esize, n = (size_t)(ctx->site_off[v + 1] - ctx->site_off[v]) * ctx->esize;
  BPX_CUDA(ctx, cudaMemcpyAsync(data, (char*)ctx->d_sites + off, n, cudaMemcpyDeviceToHost, ctx->stream));
  BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPX_OK;
}

static void apply_fill_side(bpx_ctx* ctx, applyk::Side& s, int64_t v, int bond_slot, int chi_b) {
  const VDesc& vd = ctx->h_vdesc[v];
  s.site_off = ctx->dev_site_off[v];
  s.z = vd.z;
  s.d = vd.d;
  s.bond_slot = bond_slot;
  for (int i = 0; i < vd.z; ++i) {
    s.dim[i] = vd.dim[i];
    s.in_msg[i] = ctx->msg_off[vd.in_edge[i]];
  }
  applyk::finish_side(s, chi_b);
}

// Version 3 of the two-site gates (bpx_apply3.cuh, the Gram path): three kernels per chunk of the batch -- (gate, side)
// CTAs absorb the messages and form the Gram matrices, one small CTA per gate solves the bond problem, (gate, side) CTAs
// write A' = A W -- with a work space per gate of the chunk.  Gates it declines (status 1: rank-deficient / indefinite
// message, ill-conditioned Gram matrix) come back in `rest` for the step-by-step versions, untouched.
// *taken = false: the batch has a shape the kernels do not take (nothing ran).
template <typename T>
static void apply_launch_sides(const applyk3::ApplyArgs3& a3, int bytes, int grid, cudaStream_t st) {
  applyk3::bp_apply3_sides<T><<<(int)std::min<int64_t>(2 * (a3.g1 - a3.g0), grid), applyk::NT, bytes, st>>>(a3);
}
template <typename T>
static void apply_launch_bond(const applyk3::ApplyArgs3& a3, int bytes, int grid, int bond_ctas, cudaStream_t st) {
  const int gb = (int)std::min<int64_t>(a3.g1 - a3.g0, grid);
  constexpr int HI = Elem<T>::is_complex ? 3 : 5;
  if (bond_ctas >= HI)
    applyk3::bp_apply3_bond<T, HI><<<gb, applyk3::NT_BOND, bytes, st>>>(a3);
  else
    applyk3::bp_apply3_bond<T, HI - 1><<<gb, applyk3::NT_BOND, bytes, st>>>(a3);
}
template <typename T>
static void apply_launch_final(const applyk3::ApplyArgs3& a3, int bytes, int grid, cudaStream_t st) {
  applyk3::bp_apply3_final<T><<<(int)std::min<int64_t>(2 * (a3.g1 - a3.g0), grid), applyk::NT, bytes, st>>>(a3);
}
template <typename T>
static int apply_launch_v3(bpx_ctx* ctx, const applyk3::ApplyArgs3& a3, const int bytes[3], const int grid[3], int bond_ctas,
                           cudaEvent_t* ev /* 4 events (debug timing) or NULL */) {
  if (ev) cudaEventRecord(ev[0], ctx->stream);
  apply_launch_sides<T>(a3, bytes[0], grid[0], ctx->stream);
  if (ev) cudaEventRecord(ev[1], ctx->stream);
  apply_launch_bond<T>(a3, bytes[1], grid[1], bond_ctas, ctx->stream);
  if (ev) cudaEventRecord(ev[2], ctx->stream);
  apply_launch_final<T>(a3, bytes[2], grid[2], ctx->stream);
  if (ev) cudaEventRecord(ev[3], ctx->stream);
  ctx->n_launches += 3;
  BPX_CUDA(ctx, cudaGetLastError());
  return BPX_OK;
}
// Long batches: the bond kernel (pure latency chains, no DRAM traffic) of chunk i runs BESIDE the side kernel (bound by the
// memory system) of chunk i + 1 and the final kernel of chunk i - 1: half-sized chunks, the dense kernels with one CTA per SM
// on the context's stream (S0 S1 F0 S2 F1 ..), the bond kernel with two CTAs per SM on a second stream (B0 B1 ..), events
// S_i -> B_i -> F_i; two work-space buffers suffice (S_{i+2} follows F_i in stream order).
template <typename T>
static int apply_overlapped_v3(bpx_ctx* ctx, applyk3::ApplyArgs3 a3, const int bytes[3], int bond_ctas, int64_t ng, int64_t chunk,
                               char* ws[2]) {
  if (!ctx->apply_stream) BPX_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->apply_stream, cudaStreamNonBlocking));
  const int64_t nchunks = (ng + chunk - 1) / chunk;
  std::vector<cudaEvent_t> ev((size_t)(2 * nchunks + 1));
  for (auto& e : ev) BPX_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  const int gd = ctx->num_sms, gbond = 2 * ctx->num_sms;
  BPX_CUDA(ctx, cudaEventRecord(ev[2 * nchunks], ctx->stream));  // uploads, the zeroed status
  BPX_CUDA(ctx, cudaStreamWaitEvent(ctx->apply_stream, ev[2 * nchunks], 0));
  auto args = [&](int64_t i) {
    applyk3::ApplyArgs3 a = a3;
    a.g0 = i * chunk;
    a.g1 = std::min(ng, a.g0 + chunk);
    a.base.ws = ws[i & 1];
    return a;
  };
  auto sides = [&](int64_t i) {
    apply_launch_sides<T>(args(i), bytes[0], gd, ctx->stream);
    cudaEventRecord(ev[2 * i], ctx->stream);
    cudaStreamWaitEvent(ctx->apply_stream, ev[2 * i], 0);
    apply_launch_bond<T>(args(i), bytes[1], gbond, bond_ctas, ctx->apply_stream);
    cudaEventRecord(ev[2 * i + 1], ctx->apply_stream);
  };
  auto fin = [&](int64_t i) {
    cudaStreamWaitEvent(ctx->stream, ev[2 * i + 1], 0);
    apply_launch_final<T>(args(i), bytes[2], gd, ctx->stream);
  };
  sides(0);
  for (int64_t i = 1; i < nchunks; ++i) {
    sides(i);
    fin(i - 1);
  }
  fin(nchunks - 1);
  ctx->n_launches += 3 * nchunks;
  const cudaError_t ce = cudaGetLastError();
  for (auto& e : ev) cudaEventDestroy(e);  // (released when the recorded work has completed)
  BPX_CUDA(ctx, ce);
  return BPX_OK;
}
template <typename T>
static int apply_prepare_v3(bpx_ctx* ctx, const int bytes[3], int bond_ctas, int per_sm[3]) {
  int rc;
  constexpr int HI = Elem<T>::is_complex ? 3 : 5;
  if ((rc = set_smem(ctx, applyk3::bp_apply3_sides<T>, bytes[0]))) return rc;
  if ((rc = set_smem(ctx, applyk3::bp_apply3_final<T>, bytes[2]))) return rc;
  BPX_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[0], applyk3::bp_apply3_sides<T>, applyk::NT, bytes[0]));
  BPX_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[2], applyk3::bp_apply3_final<T>, applyk::NT, bytes[2]));
  if (bond_ctas >= HI) {
    if ((rc = set_smem(ctx, applyk3::bp_apply3_bond<T, HI>, bytes[1]))) return rc;
    BPX_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[1], applyk3::bp_apply3_bond<T, HI>, applyk3::NT_BOND, bytes[1]));
  } else {
    if ((rc = set_smem(ctx, applyk3::bp_apply3_bond<T, HI - 1>, bytes[1]))) return rc;
    BPX_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[1], applyk3::bp_apply3_bond<T, HI - 1>, applyk3::NT_BOND, bytes[1]));
  }
  return BPX_OK;
}

static int apply_run_v3(bpx_ctx* ctx, const std::vector<applyk::GateDesc>& gates, const applyk::GateDesc* d_gates, const char* d_ops,
                        int normalize, double* sv_dev, int64_t sv_stride, std::vector<applyk::GateDesc>& rest, bool* taken) {
  *taken = false;
  const int64_t ng = (int64_t)gates.size();
  const bool cplx = ctx->dtype != BPX_F64;
  applyk3::SmemNeed3 need = {0, 0, 0};
  int64_t ws_stride = 0, tt_stride = 0;
  for (int64_t g = 0; g < ng; ++g) {
    const applyk3::SmemNeed3 nd = applyk3::smem_need3(gates[g], cplx);
    if (nd.all() == 0) return BPX_OK;
    need.sides = std::max(need.sides, nd.sides);
    need.bond = std::max(need.bond, nd.bond);
    need.fin = std::max(need.fin, nd.fin);
    ws_stride = std::max(ws_stride, applyk3::layout3_of(gates[g], false).total);
    tt_stride = std::max(tt_stride, (applyk3::tt_elems(gates[g]) + 1) & ~(int64_t)1);
  }
  const int bytes[3] = {(int)(need.sides * ctx->esize), (int)(need.bond * ctx->esize), (int)(need.fin * ctx->esize)};
  for (int i = 0; i < 3; ++i)
    if (bytes[i] > ctx->max_smem_optin - 1024) return BPX_OK;
  // resident CTAs per SM the bond kernel is compiled for: 4 (128 registers) / 2 for ComplexF64; one more (96 registers,
  // spills in the Jacobi steps) measured 20 % / 3 % slower (profiles/r2as_*); BPX_APPLY_BOND_CTAS=5 / 3 selects it
  int bond_ctas = cplx ? 2 : 4;
  if (const char* e = getenv("BPX_APPLY_BOND_CTAS")) bond_ctas = atoi(e);
  int rc, per_sm[3] = {0, 0, 0};
  if ((rc = cplx ? apply_prepare_v3<c64>(ctx, bytes, bond_ctas, per_sm) : apply_prepare_v3<double>(ctx, bytes, bond_ctas, per_sm))) return rc;
  if (per_sm[0] < 1 || per_sm[1] < 1 || per_sm[2] < 1) return BPX_OK;
  int grid[3] = {ctx->num_sms * per_sm[0], ctx->num_sms * per_sm[1], ctx->num_sms * per_sm[2]};
  // experiments: fewer CTAs in flight (is a kernel bound by latency or by a shared resource?)
  if (const char* e = getenv("BPX_APPLY_SIDES_GRID")) grid[0] = std::max(1, std::min(grid[0], atoi(e)));
  if (const char* e = getenv("BPX_APPLY_BOND_GRID")) grid[1] = std::max(1, std::min(grid[1], atoi(e)));
  if (const char* e = getenv("BPX_APPLY_FINAL_GRID")) grid[2] = std::max(1, std::min(grid[2], atoi(e)));
  // chunks: as many gates as fit the work-space budget, in whole waves of the side kernel when there are several chunks
  int64_t budget = (int64_t)3 << 29;  // 1.5 GiB: stays in the context's cached arena
  if (const char* e = getenv("BPX_APPLY_WS_BYTES")) budget = std::max<int64_t>(1, atoll(e));  // tests: force several chunks
  int64_t chunk = std::max<int64_t>(1, budget / (ws_stride * ctx->esize));
  if (chunk < ng && chunk > grid[0] / 2) chunk -= chunk % (grid[0] / 2);
  chunk = std::min(chunk, ng);
  const bool timing = getenv("BPX_APPLY_TIMING") != nullptr;  // debug: per-phase clock64 stamps, summary on stderr
  // OPT-IN (BPX_APPLY_OVERLAP=1), more than one chunk: the bond kernel beside the dense kernels of the neighbouring chunks
  // (apply_overlapped_v3).  Measured 7 - 16 % SLOWER than the kernels one after the other (profiles/r2bb_overlap.log).
  bool overlap = false;
  if (const char* e = getenv("BPX_APPLY_OVERLAP")) overlap = atoi(e) != 0 && chunk < ng && !timing && chunk >= 2;
  if (overlap) {
    chunk = chunk / 2;
    if (chunk > ctx->num_sms) chunk -= chunk % ctx->num_sms;  // whole waves of the one-CTA-per-SM side kernel (2 items per gate)
  }
  char *d_ws = nullptr, *d_tt = nullptr;
  int32_t* d_status = nullptr;
  if ((rc = ws_get(ctx, bpx_ctx::WS_WORK, (size_t)chunk * (overlap ? 2 : 1) * ws_stride * ctx->esize, &d_ws))) return rc;
  if ((rc = ws_get(ctx, bpx_ctx::WS_SCRATCH, (size_t)std::min<int64_t>(2 * chunk, grid[0]) * tt_stride * ctx->esize, &d_tt))) return rc;
  if ((rc = ws_get(ctx, bpx_ctx::WS_LIST, (size_t)ng * sizeof(int32_t), &d_status))) return rc;
  BPX_CUDA(ctx, cudaMemsetAsync(d_status, 0, (size_t)ng * sizeof(int32_t), ctx->stream));
  applyk3::ApplyArgs3 a3;
  a3.base.gates = d_gates;
  a3.base.n_gates = ng;
  a3.base.sites = ctx->d_sites;
  a3.base.msgs = ctx->d_msg[ctx->cur];
  a3.base.ops = d_ops;
  a3.base.ws = d_ws;
  a3.base.sv_out = sv_dev;
  a3.base.sv_stride = sv_stride;
  a3.base.normalize = normalize;
  a3.ws_stride = ws_stride;
  a3.tt = d_tt;
  a3.tt_stride = tt_stride;
  a3.status = d_status;
  a3.stamps = nullptr;
  constexpr int SS = applyk3::STAMP_SLOTS;
  long long* d_stamps = nullptr;
  if (timing) {
    if ((rc = ws_get(ctx, bpx_ctx::WS_OUT, (size_t)ng * SS * sizeof(long long), &d_stamps))) return rc;
    BPX_CUDA(ctx, cudaMemsetAsync(d_stamps, 0, (size_t)ng * SS * sizeof(long long), ctx->stream));
    a3.stamps = d_stamps;
  }
  std::vector<cudaEvent_t> evs;
  if (overlap) {
    char* ws2[2] = {d_ws, d_ws + (size_t)chunk * ws_stride * ctx->esize};
    if ((rc = cplx ? apply_overlapped_v3<c64>(ctx, a3, bytes, bond_ctas, ng, chunk, ws2)
                   : apply_overlapped_v3<double>(ctx, a3, bytes, bond_ctas, ng, chunk, ws2)))
      return rc;
  }
  for (int64_t g0 = 0; g0 < ng && !overlap; g0 += chunk) {
    a3.g0 = g0;
    a3.g1 = std::min(ng, g0 + chunk);
    cudaEvent_t* ev = nullptr;
    if (timing) {
      evs.resize(evs.size() + 4);
      ev = evs.data() + evs.size() - 4;
      for (int i = 0; i < 4; ++i) BPX_CUDA(ctx, cudaEventCreate(&ev[i]));
    }
    if ((rc = cplx ? apply_launch_v3<c64>(ctx, a3, bytes, grid, bond_ctas, ev) : apply_launch_v3<double>(ctx, a3, bytes, grid, bond_ctas, ev)))
      return rc;
  }
  std::vector<int32_t> status((size_t)ng);
  BPX_CUDA(ctx, cudaMemcpyAsync(status.data(), d_status, (size_t)ng * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (timing) {
    float ms[3] = {0, 0, 0};
    for (size_t c = 0; c + 3 < evs.size(); c += 4)
      for (int i = 0; i < 3; ++i) {
        float t = 0;
        cudaEventElapsedTime(&t, evs[c + i], evs[c + i + 1]);
        ms[i] += t;
      }
    for (cudaEvent_t e : evs) cudaEventDestroy(e);
    fprintf(stderr, "gate kernels v3: %lld gates in %zu chunk(s); kernel time sides %.3f ms, bond %.3f ms, final %.3f ms\n", (long long)ng,
            evs.size() / 4, ms[0], ms[1], ms[2]);
    std::vector<long long> st((size_t)ng * SS);
    BPX_CUDA(ctx, cudaMemcpy(st.data(), d_stamps, st.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    // (clock64 is per SM: only differences inside one kernel of one gate / side mean anything)
    static const struct { const char* name; int from, to; } ph[] = {
        {"sides: msg check 0", 0, 1}, {"sides: absorb 0", 1, 2},   {"sides: gram 0", 2, 3},    {"sides: msg check 1", 8, 9},
        {"sides: absorb 1", 9, 10},   {"sides: gram 1", 10, 11},   {"bond: eig x 2", 16, 17},  {"bond: theta + gate", 17, 18},
        {"bond: svd", 18, 19},        {"bond: Y, W", 19, 20},      {"final: side 0 + msgs", 24, 25}, {"final: side 1", 26, 27}};
    constexpr int NP = sizeof(ph) / sizeof(ph[0]);
    double acc[NP] = {0}, sw[2] = {0}, ab[5] = {0};
    int64_t cnt = 0;
    for (int64_t g = 0; g < ng; ++g) {
      const long long* t = st.data() + SS * g;
      if (status[g] != 0 || t[27] == 0) continue;
      for (int i = 0; i < NP; ++i) acc[i] += (double)(t[ph[i].to] - t[ph[i].from]);
      sw[0] += (double)t[21];
      sw[1] += (double)t[22];
      for (int i = 0; i < 4; ++i) ab[i] += (double)(t[33 + i] - t[32 + i]);
      ab[4] += (double)(t[39] - t[36]);
      ++cnt;
    }
    const double c1 = (double)std::max<int64_t>(cnt, 1);
    double tot = 0;
    for (int i = 0; i < NP; ++i) tot += acc[i];
    fprintf(stderr,
            "gate kernels v3, phase clocks per gate (mean over %lld gates; chunk %lld; CTAs/SM sides %d bond %d final %d): sum %.0f\n",
            (long long)cnt, (long long)chunk, per_sm[0], per_sm[1], per_sm[2], tot / c1);
    for (int i = 0; i < NP; ++i) fprintf(stderr, "  %-22s %10.0f  %5.1f %%\n", ph[i].name, acc[i] / c1, 100.0 * acc[i] / std::max(tot, 1.0));
    fprintf(stderr, "  Jacobi sweeps: svd %.1f, eig %.1f;  first column batch of absorb 0: gather %.0f, legs %.0f %.0f %.0f, scatter %.0f\n",
            sw[0] / c1, sw[1] / c1, ab[0] / c1, ab[1] / c1, ab[2] / c1, ab[3] / c1, ab[4] / c1);
  }
  *taken = true;
  for (int64_t g = 0; g < ng; ++g)
    if (status[g] != 0) rest.push_back(gates[g]);
  ctx->n_gates_v3 += ng - (int64_t)rest.size();
  ctx->n_gates_declined += (int64_t)rest.size();
  return BPX_OK;
}

// gates: descriptors with ws_off still unset.  Two-site batches go to version 3 first (BPX_APPLY_V3=0 disables it); what
// is left runs on version 1 (or the opt-in version 2) in chunks bounded by the work-space budget.
static int apply_run(bpx_ctx* ctx, std::vector<applyk::GateDesc>& gates_in, const void* ops_packed, size_t ops_elems,
                     int normalize, double* sv_dev, int64_t sv_stride, bool needs_ws = true) {
  if (gates_in.empty()) return BPX_OK;
  for (size_t g = 0; g < gates_in.size(); ++g) gates_in[g].sv_row_p1 = (int32_t)(g + 1);
  char* d_ops = nullptr;
  int rc = ws_get(ctx, bpx_ctx::WS_OPS, ops_elems * ctx->esize, &d_ops);
  if (rc) return rc;
  BPX_CUDA(ctx, cudaMemcpyAsync(d_ops, ops_packed, ops_elems * ctx->esize, cudaMemcpyHostToDevice, ctx->stream));
  applyk::GateDesc* d_gates = nullptr;
  std::vector<applyk::GateDesc> rest;
  std::vector<applyk::GateDesc>* gates_p = &gates_in;
  const char* v3env = getenv("BPX_APPLY_V3");
  if (needs_ws && gates_in[0].nsides == 2 && !(v3env && atoi(v3env) == 0)) {
    if ((rc = ws_upload(ctx, bpx_ctx::WS_DESC, gates_in, &d_gates))) return rc;
    bool taken = false;
    if ((rc = apply_run_v3(ctx, gates_in, d_gates, d_ops, normalize, sv_dev, sv_stride, rest, &taken))) {
      cudaStreamSynchronize(ctx->stream);
      return rc;
    }
    if (taken) {
      ctx->sites_dirty = true;
      if (rest.empty()) {
        ws_trim(ctx);
        return BPX_OK;
      }
      gates_p = &rest;
    }
  }
  std::vector<applyk::GateDesc>& gates = *gates_p;
  const int64_t ng = (int64_t)gates.size();
  // chunks: consecutive gates whose work space fits the budget (at least one gate per chunk)
  size_t free_b = 0, total_b = 0;
  BPX_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
  free_b += ctx->ws_bytes[bpx_ctx::WS_WORK];  // the cached work space is ours to re-use
  int64_t budget = (int64_t)std::max<size_t>(std::min<size_t>(free_b / 2, (size_t)8 << 30), (size_t)1 << 20) / ctx->esize;
  if (const char* e = getenv("BPX_APPLY_WS_BYTES")) budget = std::max<int64_t>(1, atoll(e) / ctx->esize);  // tests: force several chunks
  // OPT-IN version 2 of the two-site kernel (bpx_apply2.cuh: gauging and TSQR staged through shared memory)
  const char* v2env = getenv("BPX_APPLY_V2");
  bool v2 = v2env && atoi(v2env) != 0 && needs_ws;
  int64_t smem_elems = ((int64_t)ctx->max_smem_optin - 2048) / ctx->esize;
  if (const char* e = getenv("BPX_APPLY_V2_SMEM_KB"))  // tuning knob: smaller panels, more resident CTAs per SM
    smem_elems = std::max<int64_t>(64, std::min<int64_t>(smem_elems, atoll(e) * 1024 / ctx->esize));
  for (int64_t g = 0; v2 && g < ng; ++g)
    v2 = gates[g].nsides == 2 && applyk2::block_rows(gates[g].s[0], smem_elems) > 0 &&
         applyk2::block_rows(gates[g].s[1], smem_elems) > 0;
  std::vector<int64_t> chunk_begin{0};
  int64_t cur_total = 0, max_total = 0;
  for (int64_t g = 0; g < ng; ++g) {
    // plain one-site gates work in place
    const int64_t need = !needs_ws ? 2 : (v2 ? applyk2::layout2_of(gates[g], smem_elems).total : applyk::layout_of(gates[g]).total) + 2;
    if (cur_total > 0 && cur_total + need > budget) {
      chunk_begin.push_back(g);
      cur_total = 0;
    }
    gates[g].ws_off = cur_total;
    cur_total += need;
    max_total = std::max(max_total, cur_total);
  }
  chunk_begin.push_back(ng);
  char* d_ws = nullptr;
  // context-cached work space (grow-only; buffers above 1 GiB go back at the end of the call)
  rc = ws_upload(ctx, bpx_ctx::WS_DESC, gates, &d_gates);
  if (!rc) rc = ws_get(ctx, bpx_ctx::WS_WORK, (size_t)max_total * ctx->esize, &d_ws);
  cudaError_t ce = cudaSuccess;
  if (!rc) {
    for (size_t c = 0; c + 1 < chunk_begin.size() && ce == cudaSuccess; ++c) {
      applyk::ApplyArgs a;
      a.gates = d_gates + chunk_begin[c];
      a.n_gates = chunk_begin[c + 1] - chunk_begin[c];
      a.sites = ctx->d_sites;
      a.msgs = ctx->d_msg[ctx->cur];
      a.ops = d_ops;
      a.ws = d_ws;
      a.sv_out = sv_dev;  // rows through GateDesc::sv_row_p1
      a.sv_stride = sv_stride;
      a.normalize = normalize;
      const int grid = (int)std::min<int64_t>(a.n_gates, (int64_t)ctx->num_sms * 8);
      if (v2) {
        applyk2::ApplyArgs2 a2;
        a2.base = a;
        a2.smem_elems = smem_elems;
        const int bytes = (int)(smem_elems * ctx->esize);
        if (ctx->dtype == BPX_F64) {
          if ((rc = set_smem(ctx, applyk2::bp_apply_gates_v2<double>, bytes))) break;
          applyk2::bp_apply_gates_v2<double><<<grid, applyk::NT, bytes, ctx->stream>>>(a2);
        } else {
          if ((rc = set_smem(ctx, applyk2::bp_apply_gates_v2<c64>, bytes))) break;
          applyk2::bp_apply_gates_v2<c64><<<grid, applyk::NT, bytes, ctx->stream>>>(a2);
        }
      } else if (ctx->dtype == BPX_F64)
        applyk::bp_apply_gates<double><<<grid, applyk::NT, 0, ctx->stream>>>(a);
      else
        applyk::bp_apply_gates<c64><<<grid, applyk::NT, 0, ctx->stream>>>(a);
      ctx->n_launches++;
      ce = cudaGetLastError();
    }
  }
  const cudaError_t ce2 = cudaStreamSynchronize(ctx->stream);
  ws_trim(ctx);
  if (rc) return rc;
  BPX_CUDA(ctx, ce);
  BPX_CUDA(ctx, ce2);
  ctx->sites_dirty = true;  // the private tensor images of the update kernels are stale now
  return BPX_OK;
}

static int apply_common_checks(bpx_ctx* ctx, const char* name) {
  REQUIRE(ctx, ctx->mode == BPX_MODE_NORM, "%s: NORM mode only (the state is the ket layer of a norm network)", name);
  // partitioned contexts: every rank applies the gates that lie inside its own block (apply_owned_check); the incoming
  // boundary messages of its vertices include cut-edge messages pushed by the peers during the last sweep
  if (ctx->nranks > 1) return halo_gate(ctx);
  return BPX_OK;
}

// a gate can run on this rank if all its vertices are resident here (gates across a cut edge need the peer's tensor)
static int apply_owned_check(bpx_ctx* ctx, const char* name, int64_t g, int64_t v) {
  if (ctx->nranks > 1 && ctx->owner[v] != ctx->rank) {
    set_error(ctx, "%s: gate %lld touches vertex %lld, which rank %d owns: on a partitioned context every rank applies the "
              "gates inside its own block; gates across a cut edge are not supported yet", name, (long long)g, (long long)v,
              (int)ctx->owner[v]);
    return BPX_ERR_UNSUPPORTED;
  }
  return BPX_OK;
}

extern "C" int bpx_apply_two_site_gates(bpx_ctx* ctx, int64_t n_gates, const int64_t* edges, const void* ops_packed,
                                        int max_rank, int normalize, double* singular_values_out) {
  PAD(ctx, bpx::pad::apply_two(ctx, n_gates, edges, ops_packed, max_rank, normalize, singular_values_out));
  MULTI(ctx, bpx::multi::apply_two(ctx, n_gates, edges, ops_packed, max_rank, normalize, singular_values_out));
  NEED_DIMS(ctx, "bpx_apply_two_site_gates");
  int rc = apply_common_checks(ctx, "bpx_apply_two_site_gates");
  if (rc) return rc;
  REQUIRE(ctx, n_gates >= 0 && max_rank >= 0, "bpx_apply_two_site_gates: bad arguments");
  if (n_gates == 0) return BPX_OK;
  REQUIRE(ctx, edges && ops_packed, "bpx_apply_two_site_gates: NULL edges / operators");
  std::vector<applyk::GateDesc> gates((size_t)n_gates);
  std::vector<char> used((size_t)ctx->nv, 0);
  int64_t op_off = 0, sv_stride = 1;
  for (int64_t g = 0; g < n_gates; ++g) {
    const int64_t e = edges[g];
    REQUIRE(ctx, e >= 0 && e < ctx->ne, "bpx_apply_two_site_gates: gate %lld: edge %lld out of range", (long long)g, (long long)e);
    const int64_t v1 = ctx->src[e], v2 = ctx->dst[e], r = ctx->rev[e];
    REQUIRE(ctx, !used[v1] && !used[v2],
            "bpx_apply_two_site_gates: gate %lld shares a vertex with an earlier gate of the batch (gates of one call "
            "must be vertex-disjoint; apply overlapping gates in successive calls)", (long long)g);
    used[v1] = used[v2] = 1;
    if ((rc = apply_owned_check(ctx, "bpx_apply_two_site_gates", g, v1)) || (rc = apply_owned_check(ctx, "bpx_apply_two_site_gates", g, v2))) return rc;
    applyk::GateDesc& gd = gates[g];
    memset(&gd, 0, sizeof(gd));
    gd.nsides = 2;
    gd.chi_b = ctx->link_dim[e];
    apply_fill_side(ctx, gd.s[0], v1, ctx->slot[e], gd.chi_b);
    apply_fill_side(ctx, gd.s[1], v2, ctx->slot[r], gd.chi_b);
    REQUIRE(ctx, gd.s[0].d <= 16 && gd.s[1].d <= 16, "bpx_apply_two_site_gates: physical dimension > 16");
    const int m = gd.s[0].nref * gd.s[0].d, n = gd.s[1].nref * gd.s[1].d;
    int k = max_rank > 0 ? std::min(max_rank, gd.chi_b) : gd.chi_b;
    gd.k = std::min(k, std::min(m, n));
    gd.msg12 = ctx->msg_off[e];
    gd.msg21 = ctx->msg_off[r];
    gd.op_off = op_off;
    const int64_t dd = (int64_t)gd.s[0].d * gd.s[1].d;
    op_off += dd * dd;
    sv_stride = std::max<int64_t>(sv_stride, gd.chi_b);
  }
  double* d_sv = nullptr;
  if (singular_values_out && (rc = ws_get(ctx, bpx_ctx::WS_SV, (size_t)(n_gates * sv_stride) * sizeof(double), &d_sv))) return rc;
  rc = apply_run(ctx, gates, ops_packed, (size_t)op_off, normalize, d_sv, sv_stride);
  if (!rc && d_sv) {
    std::vector<double> h((size_t)(n_gates * sv_stride));
    BPX_CUDA(ctx, cudaMemcpy(h.data(), d_sv, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
    int64_t o = 0;  // packed: link_dim[edges[g]] values per gate, gate order
    for (int64_t g = 0; g < n_gates; ++g)
      for (int i = 0; i < gates[g].chi_b; ++i) singular_values_out[o++] = h[(size_t)(g * sv_stride + i)];
  }
  return rc;
}

extern "C" int bpx_apply_one_site_gates(bpx_ctx* ctx, int64_t n_gates, const int64_t* vertices, const void* ops_packed,
                                        int normalize) {
  MULTI(ctx, bpx::multi::apply_one(ctx, n_gates, vertices, ops_packed, normalize));
  NEED_DIMS(ctx, "bpx_apply_one_site_gates");
  int rc = apply_common_checks(ctx, "bpx_apply_one_site_gates");
  if (rc) return rc;
  REQUIRE(ctx, n_gates >= 0, "bpx_apply_one_site_gates: bad arguments");
  if (n_gates == 0) return BPX_OK;
  REQUIRE(ctx, vertices && ops_packed, "bpx_apply_one_site_gates: NULL vertices / operators");
  std::vector<applyk::GateDesc> gates((size_t)n_gates);
  std::vector<char> used((size_t)ctx->nv, 0);
  int64_t op_off = 0;
  for (int64_t g = 0; g < n_gates; ++g) {
    const int64_t v = vertices[g];
    REQUIRE(ctx, v >= 0 && v < ctx->nv, "bpx_apply_one_site_gates: gate %lld: vertex %lld out of range", (long long)g, (long long)v);
    REQUIRE(ctx, !used[v], "bpx_apply_one_site_gates: vertex %lld appears twice in the batch", (long long)v);
    used[v] = 1;
    if ((rc = apply_owned_check(ctx, "bpx_apply_one_site_gates", g, v))) return rc;
    applyk::GateDesc& gd = gates[g];
    memset(&gd, 0, sizeof(gd));
    gd.nsides = 1;
    apply_fill_side(ctx, gd.s[0], v, -1, 0);
    REQUIRE(ctx, gd.s[0].d <= 16, "bpx_apply_one_site_gates: physical dimension > 16");
    gd.op_off = op_off;
    op_off += (int64_t)gd.s[0].d * gd.s[0].d;
  }
  return apply_run(ctx, gates, ops_packed, (size_t)op_off, normalize, nullptr, 0, normalize != 0);
}

extern "C" int bpx_apply_stats(bpx_ctx* ctx, int64_t out[2], int reset) {
  if (!ctx) return BPX_ERR_INVALID;
  int64_t acc[2] = {0, 0};
  if (!ctx->children.empty()) {
    for (bpx_ctx* c : ctx->children) {
      int64_t t[2];
      bpx_apply_stats(c, t, reset);
      acc[0] += t[0];
      acc[1] += t[1];
    }
  } else {
    acc[0] = ctx->n_gates_v3;
    acc[1] = ctx->n_gates_declined;
    if (reset) ctx->n_gates_v3 = ctx->n_gates_declined = 0;
  }
  if (out) memcpy(out, acc, sizeof(acc));
  return BPX_OK;
}

// ---- two-site expectation values in the BP environment (bpx_expect2.cuh) ------------------------------------
extern "C" int bpx_edge_expect(bpx_ctx* ctx, int64_t n_edges, const int64_t* edges, const void* ops_packed, void* num_out,
                               void* den_out) {
  MULTI(ctx, bpx::multi::edge_expect(ctx, n_edges, edges, ops_packed, num_out, den_out));
  NEED_DIMS(ctx, "bpx_edge_expect");
  REQUIRE(ctx, ctx->mode == BPX_MODE_NORM, "bpx_edge_expect: NORM mode only");
  REQUIRE(ctx, n_edges >= 0, "bpx_edge_expect: bad arguments");
  if (n_edges == 0) return BPX_OK;
  REQUIRE(ctx, edges && ops_packed && num_out && den_out, "bpx_edge_expect: NULL argument");
  int rc = halo_gate(ctx);
  if (rc) return rc;
  std::vector<expect2::EdgeDesc> desc((size_t)n_edges);
  int64_t op_off = 0;
  for (int64_t g = 0; g < n_edges; ++g) {
    const int64_t e = edges[g];
    REQUIRE(ctx, e >= 0 && e < ctx->ne, "bpx_edge_expect: entry %lld: edge %lld out of range", (long long)g, (long long)e);
    const int64_t v1 = ctx->src[e], v2 = ctx->dst[e], r = ctx->rev[e];
    if ((rc = apply_owned_check(ctx, "bpx_edge_expect", g, v1)) || (rc = apply_owned_check(ctx, "bpx_edge_expect", g, v2))) return rc;
    expect2::EdgeDesc& d = desc[g];
    memset(&d, 0, sizeof(d));
    d.chi_b = ctx->link_dim[e];
    apply_fill_side(ctx, d.s[0], v1, ctx->slot[e], d.chi_b);
    apply_fill_side(ctx, d.s[1], v2, ctx->slot[r], d.chi_b);
    d.op_off = op_off;
    const int64_t dd = (int64_t)d.s[0].d * d.s[1].d;
    op_off += dd * dd;
  }
  // chunks bounded by the work-space budget, like the gate layers
  size_t free_b = 0, total_b = 0;
  BPX_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
  int64_t budget = (int64_t)std::max<size_t>(std::min<size_t>(free_b / 2, (size_t)8 << 30), (size_t)1 << 20) / ctx->esize;
  if (const char* e = getenv("BPX_APPLY_WS_BYTES")) budget = std::max<int64_t>(1, atoll(e) / ctx->esize);
  std::vector<int64_t> chunk_begin{0};
  int64_t cur_total = 0, max_total = 0;
  for (int64_t g = 0; g < n_edges; ++g) {
    const int64_t need = expect2::layout_of(desc[g]).total + 2;
    if (cur_total > 0 && cur_total + need > budget) {
      chunk_begin.push_back(g);
      cur_total = 0;
    }
    desc[g].ws_off = cur_total;
    cur_total += need;
    max_total = std::max(max_total, cur_total);
  }
  chunk_begin.push_back(n_edges);
  expect2::EdgeDesc* d_desc = nullptr;
  char *d_ws = nullptr, *d_ops = nullptr, *d_out = nullptr;
  rc = ws_upload(ctx, bpx_ctx::WS_DESC, desc, &d_desc);
  if (!rc) rc = ws_get(ctx, bpx_ctx::WS_WORK, (size_t)max_total * ctx->esize, &d_ws);
  if (!rc) rc = ws_get(ctx, bpx_ctx::WS_OPS, (size_t)op_off * ctx->esize, &d_ops);
  if (!rc) rc = ws_get(ctx, bpx_ctx::WS_OUT, (size_t)(2 * n_edges) * ctx->esize, &d_out);
  cudaError_t ce = cudaSuccess;
  if (!rc) {
    ce = cudaMemcpyAsync(d_ops, ops_packed, (size_t)op_off * ctx->esize, cudaMemcpyHostToDevice, ctx->stream);
    for (size_t c = 0; c + 1 < chunk_begin.size() && ce == cudaSuccess; ++c) {
      expect2::ExpectArgs a;
      a.edges = d_desc + chunk_begin[c];
      a.n_edges = chunk_begin[c + 1] - chunk_begin[c];
      a.sites = ctx->d_sites;
      a.msgs = ctx->d_msg[ctx->cur];
      a.ops = d_ops;
      a.ws = d_ws;
      a.num_out = d_out + (size_t)chunk_begin[c] * ctx->esize;
      a.den_out = d_out + (size_t)(n_edges + chunk_begin[c]) * ctx->esize;
      const int grid = (int)std::min<int64_t>(a.n_edges, (int64_t)ctx->num_sms * 8);
      if (ctx->dtype == BPX_F64)
        expect2::bp_edge_expect<double><<<grid, applyk::NT, 0, ctx->stream>>>(a);
      else
        expect2::bp_edge_expect<c64><<<grid, applyk::NT, 0, ctx->stream>>>(a);
      ctx->n_launches++;
      ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(num_out, d_out, (size_t)n_edges * ctx->esize, cudaMemcpyDeviceToHost, ctx->stream);
    if (ce == cudaSuccess)
      ce = cudaMemcpyAsync(den_out, d_out + (size_t)n_edges * ctx->esize, (size_t)n_edges * ctx->esize, cudaMemcpyDeviceToHost, ctx->stream);
  }
  const cudaError_t ce2 = cudaStreamSynchronize(ctx->stream);
  ws_trim(ctx);
  if (rc) return rc;
  BPX_CUDA(ctx, ce);
  BPX_CUDA(ctx, ce2);
  return BPX_OK;
}

extern "C" void* bpx_device_messages(bpx_ctx* ctx) {
  if (ctx && !ctx->children.empty()) return ctx->children.size() == 1 ? bpx_device_messages(ctx->children[0]) : nullptr;
  return (ctx && ctx->dims_set) ? ctx->d_msg[ctx->cur] : nullptr;
}
extern "C" void* bpx_device_site_tensors(bpx_ctx* ctx) {
  if (ctx && !ctx->children.empty()) return ctx->children.size() == 1 ? bpx_device_site_tensors(ctx->children[0]) : nullptr;
  return (ctx && ctx->dims_set) ? ctx->d_sites : nullptr;
}

extern "C" int bpx_synchronize(bpx_ctx* ctx) {
  MULTI(ctx, bpx::multi::each(ctx, [&](bpx_ctx* c) -> int { return bpx_synchronize(c); }));
  if (!ctx) return BPX_ERR_INVALID;
  BPX_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->dims_set) {
    int rc_ = halo_gate(ctx);
    if (rc_) return rc_;
  }
  BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPX_OK;
}

// ---- device-side synthetic inputs (benchmarks at sizes the host cannot stage) ----------------------------------
extern "C" int bpx_fill_synthetic(bpx_ctx* ctx, uint64_t seed) {
  PAD(ctx, bpx::pad::fill_synthetic(ctx, seed));
  MULTI(ctx, bpx::multi::each(ctx, [&](bpx_ctx* c) -> int { return bpx_fill_synthetic(c, seed); }));
  NEED_DIMS(ctx, "bpx_fill_synthetic");
  REQUIRE(ctx, ctx->mode == BPX_MODE_NORM, "bpx_fill_synthetic: NORM mode only");
  const int dpe = ctx->esize / 8;
  if (ctx->nv > 0) {
    dim3 grid(32, (unsigned)std::min<int64_t>(ctx->nv, 4096));
    fill_sites_randn<<<grid, 256, 0, ctx->stream>>>(ctx->d_vdesc, ctx->nv, seed, dpe, (double*)ctx->d_sites);
    ctx->n_launches++;
    BPX_CUDA(ctx, cudaGetLastError());
  }
  ctx->sites_dirty = true;
  if (ctx->ne > 0) {
    int32_t* d_chi = nullptr;
    int rc = upload(ctx, &d_chi, ctx->link_dim);
    if (rc) return rc;
    ctx->cur = 0;
    const int64_t blocks = (ctx->ne * 32 + 255) / 256;
    fill_messages_positive<<<(unsigned)blocks, 256, 0, ctx->stream>>>(ctx->d_msg_off, nullptr, ctx->ne, ctx->nv, seed, dpe, d_chi,
                                                                    (double*)ctx->d_msg[0]);
    ctx->n_launches++;
    cudaError_t ce = cudaGetLastError();
    if (ce == cudaSuccess)
      ce = cudaMemcpyAsync(ctx->d_msg[1], ctx->d_msg[0], (size_t)ctx->msg_off[ctx->ne] * ctx->esize, cudaMemcpyDeviceToDevice, ctx->stream);
    if (ce == cudaSuccess && collective_fence(ctx) != BPX_OK) ce = cudaErrorUnknown;
    cudaError_t ce2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_chi);
    BPX_CUDA(ctx, ce);
    BPX_CUDA(ctx, ce2);
  }
  BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPX_OK;
}

// ---- shared RNG -----------------------------------------------------------------------------------
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

static inline double randn_at(uint64_t seed, uint64_t stream, uint64_t i) {
  const uint64_t base = splitmix64(seed ^ splitmix64(stream + 0x632BE59BD9B4E019ull));
  const uint64_t a = splitmix64(base + 2 * i), b = splitmix64(base + 2 * i + 1);
  const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);  // (0, 1]
  const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);          // [0, 1)
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

extern "C" int bpx_fill_randn(uint64_t seed, uint64_t stream, int dtype, int64_t n, void* out) {
  if (n < 0 || (n > 0 && !out) || (dtype != BPX_F64 && dtype != BPX_C64)) return BPX_ERR_INVALID;
  double* o = (double*)out;
  if (dtype == BPX_F64) {
    for (int64_t i = 0; i < n; ++i) o[i] = randn_at(seed, stream, (uint64_t)i);
  } else {
    const double s = 0.70710678118654752440;
    for (int64_t i = 0; i < 2 * n; ++i) o[i] = s * randn_at(seed, stream, (uint64_t)i);
  }
  return BPX_OK;
}

// debug only (not part of include/bpx.h): per-phase timestamps of CTA 0 in BPX_ONCHIP_TIMING builds
extern "C" int bpx_debug_timing(bpx_ctx* ctx, long long* out, int n) {
  MULTI(ctx, bpx_debug_timing(MULTI0(ctx), out, n));
  if (!ctx) return BPX_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (!ctx->d_timing) {
    if (cudaMalloc(&ctx->d_timing, 8 * 32 * 16 * sizeof(long long)) != cudaSuccess) return BPX_ERR_ALLOC;
    cudaMemset(ctx->d_timing, 0, 8 * 32 * 16 * sizeof(long long));
    return BPX_OK;
  }
  cudaStreamSynchronize(ctx->stream);
  cudaMemcpy(out, ctx->d_timing, sizeof(long long) * std::min(n, 8 * 32 * 16), cudaMemcpyDeviceToHost);
  return BPX_OK;
}
