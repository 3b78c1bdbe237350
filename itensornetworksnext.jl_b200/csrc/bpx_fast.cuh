// Specialised (bucket-specific) kernels: dispatch.  Kernels live in bpx_onchip.cuh / bpx_sliced.cuh.
#pragma once
#include <algorithm>

#include "bpx_ctx.h"
#include "bpx_onchip.cuh"
#include "bpx_sliced.cuh"
#include "bpx_sliced2.cuh"
#include "bpx_onchip16.cuh"
#include "bpx_onchip16c.cuh"
#include "bpx_onchip8c.cuh"
#include "bpx_vertex.cuh"

namespace bpx {

// ComplexF64 on-chip families: link dims <= 8 with degree 2..4 -> 8 (bpx_onchip8c.cuh); link dims <= 16 with degree
// 1..3 -> 16 (bpx_onchip16c.cuh); smaller / per-leg different dims are zero-padded.  0: neither.
inline int complex_family(const Bucket& b) {
  if (b.d < 1 || b.z < 1) return 0;
  if (b.max_dim <= 8 && b.z >= 2 && b.z <= 4) return 8;
  if (b.max_dim <= 16 && b.z <= 3) return 16;
  return 0;
}

// Float64 buckets for the hand-tuned kernels: exact shapes only
inline bool real_tuned_c8(const Bucket& b) { return b.d == 2 && b.z >= 2 && b.z <= 4 && b.chi == 8; }
inline bool real_tuned_c16(const Bucket& b) { return b.d == 2 && ((b.z == 3 && b.chi == 16) || (b.z == 6 && b.chi == 4)); }
// buckets served by the templated slice kernel of bpx_onchip8c.cuh: ComplexF64 family 8, and every Float64 bucket with
// degree 2..4 and link dims <= 8 (any d, per-leg dims) that the tuned kernels do not take
inline bool uses_c8x(const bpx_ctx* ctx, const Bucket& b) {
  if (ctx->dtype == BPX_C64) return complex_family(b) == 8;
  return b.d >= 1 && b.z >= 2 && b.z <= 4 && b.max_dim <= 8 && !real_tuned_c8(b);
}

// ... and by the 16-wide slice kernel of bpx_onchip16c.cuh: ComplexF64 family 16, and the remaining Float64 buckets with
// degree 1..3 and link dims <= 16
inline bool uses_c16x(const bpx_ctx* ctx, const Bucket& b) {
  if (ctx->dtype == BPX_C64) return complex_family(b) == 16;
  return b.d >= 1 && b.z >= 1 && b.z <= 3 && b.max_dim <= 16 && !real_tuned_c8(b) && !real_tuned_c16(b) && !uses_c8x(ctx, b);
}

inline bool fast_kernel_supported(bpx_ctx* ctx, const Bucket& b, int kernel) {
  if (kernel == BPX_KERNEL_GENERIC) return true;
  if (kernel == BPX_KERNEL_VERTEX)  // single-layer networks, uniform link dim 2..4, factor of <= 64 doubles (bpx_vertex.cuh)
    return ctx->mode == BPX_MODE_SINGLE && b.chi >= 2 && vertexk::shape_supported(ctx->dtype == BPX_C64, b.z, b.chi) &&
           ctx->msg_off[ctx->ne] < (1ll << 31);
  if (ctx->mode != BPX_MODE_NORM) return false;
  if (kernel == BPX_KERNEL_ONCHIP) {
    // ComplexF64: chi = 16, degree 1..3, any physical dimension (bpx_onchip16c.cuh)
    // ... and chi = 8, degree 2..4 (bpx_onchip8c.cuh)
    if (ctx->dtype == BPX_C64) {
      const int fam = complex_family(b);
      if (fam == 16) return (size_t)ctx->max_smem_optin >= onchip16c::SMEM_BYTES16C;
      if (fam == 8) return (size_t)ctx->max_smem_optin >= onchip8c::SMEM_BYTES8C;
      return false;
    }
    if (ctx->dtype != BPX_F64) return false;
    if (real_tuned_c8(b)) return (size_t)ctx->max_smem_optin >= onchip::SMEM_BYTES;
    // 16-wide on-chip kernel: degree 3 / chi 16, or degree 6 / chi 4 with legs paired into super-legs
    if (real_tuned_c16(b)) return (size_t)ctx->max_smem_optin >= onchip16::SMEM_BYTES16;
    if (uses_c8x(ctx, b)) return (size_t)ctx->max_smem_optin >= onchip8c::SMEM_BYTES8C;
    if (uses_c16x(ctx, b)) return (size_t)ctx->max_smem_optin >= onchip16c::SMEM_BYTES16C;
    return false;
  }
  if (kernel == BPX_KERNEL_SLICED)
    return ctx->dtype == BPX_F64 && b.z == 4 && b.chi == 16 && b.d == 2 && (size_t)ctx->max_smem_optin >= sliced::SMEM_BYTES;
  return false;
}

inline int fast_kernel_for(bpx_ctx* ctx, const Bucket& b) {
  if (fast_kernel_supported(ctx, b, BPX_KERNEL_ONCHIP)) return BPX_KERNEL_ONCHIP;
  if (fast_kernel_supported(ctx, b, BPX_KERNEL_SLICED)) return BPX_KERNEL_SLICED;
  if (fast_kernel_supported(ctx, b, BPX_KERNEL_VERTEX)) return BPX_KERNEL_VERTEX;
  return BPX_KERNEL_GENERIC;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libbpx links cudart only)
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TensorMapEncodeTiledFn tensor_map_encoder() {
  static TensorMapEncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess)
      fn = (TensorMapEncodeTiledFn)p;
    cudaGetLastError();
  }
  return fn;
}
// the two strided views of a buffer of `n` doubles that the sliced kernel's half slices use (bpx_sliced2.cuh)
inline int sliced2_encode_views(bpx_ctx* ctx, void* base, size_t n, CUtensorMap* a0h, CUtensorMap* a3h) {
  TensorMapEncodeTiledFn enc = tensor_map_encoder();
  if (!enc) {
    set_error(ctx, "cuTensorMapEncodeTiled is not available from this driver");
    return BPX_ERR_CUDA;
  }
  const cuuint32_t ones[3] = {1, 1, 1};
  {
    const cuuint64_t dims[3] = {256, 32, (n + 8191) / 8192};
    const cuuint64_t strides[2] = {256 * 8, 8192 * 8};
    const cuuint32_t box[3] = {256, 1, 16};
    const CUresult r = enc(a0h, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, getenv("BPX_TMAP_NOPROMO") ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error(ctx, "cuTensorMapEncodeTiled(a0-half-slice view) failed: %d", (int)r);
      return BPX_ERR_CUDA;
    }
  }
  {
    const cuuint64_t dims[3] = {128, 2, (n + 255) / 256};
    const cuuint64_t strides[2] = {128 * 8, 256 * 8};
    const cuuint32_t box[3] = {128, 1, 32};
    const CUresult r = enc(a3h, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, getenv("BPX_TMAP_NOPROMO") ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error(ctx, "cuTensorMapEncodeTiled(a3-half-slice view) failed: %d", (int)r);
      return BPX_ERR_CUDA;
    }
  }
  return BPX_OK;
}

// Build the launch groups: all ONCHIP buckets share ONE persistent launch over a cost-sorted item list
// (degree 4 first); the first of them is the group leader, the others are skipped by the sweep loop.
inline int fast_prepare(bpx_ctx* ctx) {
  if (ctx->d_onchip_items) {
    cudaFree(ctx->d_onchip_items);
    ctx->d_onchip_items = nullptr;
  }
  if (ctx->d_sites_swz) {
    cudaFree(ctx->d_sites_swz);
    ctx->d_sites_swz = nullptr;
  }
  ctx->n_onchip_items = 0;
  std::vector<int> group;
  int generic_leader = -1;
  for (int i = 0; i < (int)ctx->buckets.size(); ++i) {
    ctx->buckets[i].leader = i;
    if (ctx->dtype == BPX_F64 && ctx->buckets[i].kernel == BPX_KERNEL_ONCHIP && real_tuned_c8(ctx->buckets[i]) && !ctx->buckets[i].my_vertices.empty())
      group.push_back(i);
    if (ctx->buckets[i].kernel == BPX_KERNEL_GENERIC && !ctx->buckets[i].my_edges.empty()) {
      if (generic_leader < 0) generic_leader = i;
      ctx->buckets[i].leader = generic_leader;  // all generic buckets share one launch
    }
  }
  // ---- VERTEX buckets (single-layer, thread per vertex): structure-of-arrays descriptors, one launch per bucket ----
  for (Bucket& b : ctx->buckets) {
    if (b.d_vx_site) cudaFree(b.d_vx_site);
    if (b.d_vx_moff) cudaFree(b.d_vx_moff);
    b.d_vx_site = nullptr;
    b.d_vx_moff = nullptr;
    if (b.kernel != BPX_KERNEL_VERTEX || b.my_vertices.empty()) continue;
    const size_t n = b.my_vertices.size();
    std::vector<int64_t> site(n);
    std::vector<int32_t> moff(2 * (size_t)b.z * n);
    b.vx_out_contig = 1;
    for (size_t i = 0; i < n; ++i) {
      const int32_t v = b.my_vertices[i];
      site[i] = ctx->dev_site_off[v];
      for (int k = 0; k < b.z; ++k) {
        const int32_t e = ctx->out_edge[v][k];
        moff[(size_t)k * n + i] = (int32_t)ctx->msg_off[ctx->rev[e]];      // message arriving on leg k
        moff[(size_t)(b.z + k) * n + i] = (int32_t)ctx->msg_off[e];        // message leaving on leg k
        if (ctx->msg_off[e] != ctx->msg_off[ctx->out_edge[v][0]] + (int64_t)k * b.chi) b.vx_out_contig = 0;
      }
    }
    cudaError_t e = cudaMalloc((void**)&b.d_vx_site, n * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMalloc((void**)&b.d_vx_moff, moff.size() * sizeof(int32_t));
    if (e != cudaSuccess) {
      set_error(ctx, "cudaMalloc(vertex-kernel descriptors) failed: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return BPX_ERR_ALLOC;
    }
    BPX_CUDA(ctx, cudaMemcpy(b.d_vx_site, site.data(), n * sizeof(int64_t), cudaMemcpyHostToDevice));
    BPX_CUDA(ctx, cudaMemcpy(b.d_vx_moff, moff.data(), moff.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  if (ctx->d_sliced_items) {
    cudaFree(ctx->d_sliced_items);
    ctx->d_sliced_items = nullptr;
  }
  if (ctx->d_fast_scratch) {
    cudaFree(ctx->d_fast_scratch);
    ctx->d_fast_scratch = nullptr;
  }
  ctx->n_sliced_items = 0;
  bool need_image = false;
  if (ctx->d_onchip16_items) {
    cudaFree(ctx->d_onchip16_items);
    ctx->d_onchip16_items = nullptr;
  }
  ctx->n_onchip16_items = 0;
  if (ctx->d_onchip16c_items) {
    cudaFree(ctx->d_onchip16c_items);
    ctx->d_onchip16c_items = nullptr;
  }
  ctx->n_onchip16c_slots = 0;
  ctx->onchip16c_grid = 0;
  {
    // ---- 16-wide slice-kernel buckets (complex family 16 / general real, degree 1..3): one launch; items scheduled onto
    // the CTAs here (LPT), laid out as rounds ----
    const bool cplx = ctx->dtype == BPX_C64;
    std::vector<onchip16c::ItemDesc> its;
    std::vector<double> cost;
    int leader = -1;
    int64_t img_total = 0;  // doubles of the zero-padded image
    for (int i = 0; i < (int)ctx->buckets.size(); ++i) {
      Bucket& b = ctx->buckets[i];
      if (b.kernel != BPX_KERNEL_ONCHIP || !uses_c16x(ctx, b) || b.my_vertices.empty()) continue;
      if (leader < 0) leader = i;
      b.leader = leader;
      for (int32_t v : b.my_vertices) {
        auto edge_of = [&](onchip16c::ItemDesc& d, int o, int l) {
          const int32_t e = ctx->out_edge[v][l];
          d.out_edge[o] = e;
          d.out_off[o] = ctx->msg_off[e];
          d.out_dim[o] = ctx->h_vdesc[v].dim[l];
          d.need = std::max<int64_t>(d.need, ctx->upload_end[e]);
          d.peer[o] = (!ctx->owner.empty() && ctx->owner[ctx->dst[e]] != ctx->rank) ? ctx->owner[ctx->dst[e]] : -1;
        };
        auto in_of = [&](int l) { return ctx->msg_off[ctx->rev[ctx->out_edge[v][l]]]; };
        onchip16c::ItemDesc d;
        memset(&d, 0, sizeof(d));
        d.site_off = img_total;
        d.phys = b.d;
        d.d = cplx ? b.d : (b.d + 1) / 2;  // slices: physical values (complex) or physical pairs (real)
        img_total += (int64_t)d.d * (b.z == 3 ? onchip16c::NSL3 : (b.z == 2 ? onchip16c::NSL2 : onchip16c::NSL1));
        d.canon_off = ctx->dev_site_off[v];
        d.peer[0] = d.peer[1] = -1;
        d.first = 1;
        for (int l = 0; l < 3; ++l) d.dim[l] = l < b.z ? ctx->h_vdesc[v].dim[l] : 1;
        for (int l = 0; l < b.z; ++l) d.need = std::max<int64_t>(d.need, ctx->upload_end[ctx->rev[ctx->out_edge[v][l]]]);
        if (b.z == 3) {
          // one item per output leg; (first, second) absorbed message: out2 (M0, M1), out1 (M0, M2), out0 (M2, M1)
          static const int first_leg[3] = {2, 0, 0}, second_leg[3] = {1, 2, 1};
          for (int leg = 2; leg >= 0; --leg) {
            d.kind = 0;
            d.leg = leg;
            d.in_off[0] = in_of(first_leg[leg]);
            d.in_off[1] = in_of(second_leg[leg]);
            d.in_dim[0] = d.dim[first_leg[leg]];
            d.in_dim[1] = d.dim[second_leg[leg]];
            edge_of(d, 0, leg);
            its.push_back(d);
            cost.push_back((cplx ? 16000.0 : 4000.0) * d.d + 3000.0);
            d.first = 0;
          }
        } else if (b.z == 2) {
          d.kind = 1;
          d.in_off[0] = in_of(0);
          d.in_off[1] = in_of(1);
          d.in_dim[0] = d.dim[0];
          d.in_dim[1] = d.dim[1];
          edge_of(d, 0, 0);
          edge_of(d, 1, 1);
          its.push_back(d);
          cost.push_back((cplx ? 3000.0 : 1000.0) * d.d + 3000.0);
        } else {
          d.kind = 2;
          edge_of(d, 0, 0);
          its.push_back(d);
          cost.push_back(3000.0);
        }
      }
    }
    if (!its.empty()) {
      const int G = std::min<int>((int)its.size(), ctx->num_sms);
      std::vector<int> order(its.size());
      for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
      std::vector<std::vector<int>> per_cta(G);
      std::vector<double> load(G, 0.0);
      for (int i : order) {  // longest processing time first onto the least loaded CTA
        const int c = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        per_cta[c].push_back(i);
        load[c] += cost[i];
      }
      size_t rounds = 0;
      for (auto& l : per_cta) rounds = std::max(rounds, l.size());
      onchip16c::ItemDesc null_item;
      memset(&null_item, 0, sizeof(null_item));
      null_item.kind = -1;
      std::vector<onchip16c::ItemDesc> slots(rounds * G, null_item);
      for (int c = 0; c < G; ++c)
        for (size_t r = 0; r < per_cta[c].size(); ++r) slots[r * G + c] = its[per_cta[c][r]];
      ctx->n_onchip16c_slots = (int)slots.size();
      ctx->onchip16c_grid = G;
      if (getenv("BPX_IO_DEBUG")) {
        int64_t mx = 0;
        for (auto& it : its) mx = std::max(mx, it.need);
        fprintf(stderr, "[bpx c16c] items %zu slots %zu grid %d max need %lld img doubles %lld\n", its.size(), slots.size(), G, (long long)mx,
                (long long)img_total);
      }
      cudaError_t e = cudaMalloc((void**)&ctx->d_onchip16c_items, slots.size() * sizeof(onchip16c::ItemDesc));
      if (e != cudaSuccess) {
        set_error(ctx, "cudaMalloc(complex on-chip items) failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return BPX_ERR_ALLOC;
      }
      BPX_CUDA(ctx, cudaMemcpy(ctx->d_onchip16c_items, slots.data(), slots.size() * sizeof(onchip16c::ItemDesc), cudaMemcpyHostToDevice));
      BPX_CUDA(ctx, cudaFuncSetAttribute(onchip16c::bp_update_onchip_c16x<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)onchip16c::SMEM_BYTES16C));
      BPX_CUDA(ctx, cudaFuncSetAttribute(onchip16c::bp_update_onchip_c16x<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)onchip16c::SMEM_BYTES16C));
      if (ctx->d_img16c) cudaFree(ctx->d_img16c);
      ctx->d_img16c = nullptr;
      e = cudaMalloc(&ctx->d_img16c, std::max<size_t>(16, (size_t)img_total * sizeof(double)));
      if (e != cudaSuccess) {
        set_error(ctx, "cudaMalloc(complex 16-wide tensor image) failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return BPX_ERR_ALLOC;
      }
      ctx->sites_dirty = true;
    }
  }
  if (ctx->d_onchip8c_items) {
    cudaFree(ctx->d_onchip8c_items);
    ctx->d_onchip8c_items = nullptr;
  }
  ctx->n_onchip8c_slots = 0;
  ctx->onchip8c_grid = 0;
  {
    // ---- slice-kernel buckets (complex family 8 / general real, degree 2..4): one launch; degree-4 vertices as two half
    // items (branch P / Q) ----
    const bool cplx = ctx->dtype == BPX_C64;
    std::vector<onchip8c::ItemDesc> its;
    std::vector<double> cost;
    int leader = -1;
    int64_t img_total = 0;  // doubles of the zero-padded image
    for (int i = 0; i < (int)ctx->buckets.size(); ++i) {
      Bucket& b = ctx->buckets[i];
      if (b.kernel != BPX_KERNEL_ONCHIP || !uses_c8x(ctx, b) || b.my_vertices.empty()) continue;
      if (leader < 0) leader = i;
      b.leader = leader;
      for (int32_t v : b.my_vertices) {
        onchip8c::ItemDesc d;
        memset(&d, 0, sizeof(d));
        d.site_off = img_total;
        d.phys = b.d;
        d.d = cplx ? b.d : (b.d + 1) / 2;  // slices: physical values (complex) or physical pairs (real)
        img_total += (int64_t)d.d * (onchip::NELEM >> (3 * (4 - b.z)));
        d.canon_off = ctx->dev_site_off[v];
        d.first = 1;
        for (int l = 0; l < 4; ++l) d.dim[l] = l < b.z ? ctx->h_vdesc[v].dim[l] : 1;
        for (int l = 0; l < b.z; ++l) {
          const int32_t e = ctx->out_edge[v][l];
          d.in_off[l] = ctx->msg_off[ctx->rev[e]];
          d.need = std::max<int64_t>(d.need, std::max(ctx->upload_end[e], ctx->upload_end[ctx->rev[e]]));
        }
        auto tile = [&](int tl, int leg) {
          const int32_t e = ctx->out_edge[v][leg];
          d.out_edge[tl] = e;
          d.out_off[tl] = ctx->msg_off[e];
          d.out_dim[tl] = d.dim[leg];
          d.peer[tl] = (!ctx->owner.empty() && ctx->owner[ctx->dst[e]] != ctx->rank) ? ctx->owner[ctx->dst[e]] : -1;
        };
        for (int tl = 0; tl < onchip8c::MAXT; ++tl) d.peer[tl] = -1;
        if (b.z == 4) {
          d.kind = 0;  // branch P: out3, out2
          tile(0, 3);
          tile(1, 2);
          its.push_back(d);
          cost.push_back((cplx ? 6.0 * 4096 : 6.0 * 1024) * d.d + 1500.0);
          d.kind = 1;  // branch Q: out1, out0
          d.first = 0;
          tile(0, 1);
          tile(1, 0);
          its.push_back(d);
          cost.push_back((cplx ? 6.0 * 4096 : 6.0 * 1024) * d.d + 1500.0);
        } else if (b.z == 3) {
          d.kind = 2;
          tile(0, 2);
          tile(1, 1);
          tile(2, 0);
          its.push_back(d);
          cost.push_back((cplx ? 8.0 * 512 : 8.0 * 128) * d.d + 1000.0 * d.d + 1500.0);
        } else {
          d.kind = 3;
          tile(0, 1);
          tile(1, 0);
          its.push_back(d);
          cost.push_back(2000.0);
        }
      }
    }
    if (!its.empty()) {
      const int G = std::min<int>((int)its.size(), ctx->num_sms);
      // longest processing time first onto the least loaded CTA; equal-cost items keep their (lattice) order, which is
      // the order in which a streamed upload delivers their messages
      std::vector<int> order(its.size());
      for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
      std::vector<std::vector<int>> per_cta(G);
      std::vector<double> load(G, 0.0);
      for (int i : order) {
        const int c = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        per_cta[c].push_back(i);
        load[c] += cost[i];
      }
      size_t rounds = 0;
      for (auto& l : per_cta) rounds = std::max(rounds, l.size());
      onchip8c::ItemDesc null_item;
      memset(&null_item, 0, sizeof(null_item));
      null_item.kind = -1;
      std::vector<onchip8c::ItemDesc> slots(rounds * G, null_item);
      for (int c = 0; c < G; ++c)
        for (size_t r = 0; r < per_cta[c].size(); ++r) slots[r * G + c] = its[per_cta[c][r]];
      ctx->n_onchip8c_slots = (int)slots.size();
      ctx->onchip8c_grid = G;
      cudaError_t e = cudaMalloc((void**)&ctx->d_onchip8c_items, slots.size() * sizeof(onchip8c::ItemDesc));
      if (e != cudaSuccess) {
        set_error(ctx, "cudaMalloc(complex chi=8 items) failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return BPX_ERR_ALLOC;
      }
      BPX_CUDA(ctx, cudaMemcpy(ctx->d_onchip8c_items, slots.data(), slots.size() * sizeof(onchip8c::ItemDesc), cudaMemcpyHostToDevice));
      BPX_CUDA(ctx, cudaFuncSetAttribute(onchip8c::bp_update_onchip_c8x<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)onchip8c::SMEM_BYTES8C));
      BPX_CUDA(ctx, cudaFuncSetAttribute(onchip8c::bp_update_onchip_c8x<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)onchip8c::SMEM_BYTES8C));
      if (ctx->d_img8c) cudaFree(ctx->d_img8c);
      ctx->d_img8c = nullptr;
      e = cudaMalloc(&ctx->d_img8c, std::max<size_t>(16, (size_t)img_total * sizeof(double)));
      if (e != cudaSuccess) {
        set_error(ctx, "cudaMalloc(complex 8-wide tensor image) failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return BPX_ERR_ALLOC;
      }
      ctx->sites_dirty = true;
    }
  }
  // ---- 16-wide ON-CHIP buckets (degree 3 / chi 16; degree 6 / chi 4 in pair mode): one launch ----
  {
    std::vector<onchip16::ItemDesc> it16;
    int leader16 = -1;
    for (int i = 0; i < (int)ctx->buckets.size(); ++i) {
      Bucket& b = ctx->buckets[i];
      if (ctx->dtype != BPX_F64 || b.kernel != BPX_KERNEL_ONCHIP || !real_tuned_c16(b) || b.my_vertices.empty()) continue;
      if (leader16 < 0) leader16 = i;
      b.leader = leader16;
      for (int32_t v : b.my_vertices) {
        onchip16::ItemDesc d;
        memset(&d, 0, sizeof(d));
        d.site_off = ctx->dev_site_off[v];
        d.pair_mode = b.z == 6 ? 1 : 0;
        for (int l = 0; l < 6; ++l) d.peer[l] = -1;
        for (int l = 0; l < b.z; ++l) {
          const int32_t e = ctx->out_edge[v][l];
          d.out_edge[l] = e;
          d.out_off[l] = ctx->msg_off[e];
          d.in_off[l] = ctx->msg_off[ctx->rev[e]];
          d.peer[l] = (!ctx->owner.empty() && ctx->owner[ctx->dst[e]] != ctx->rank) ? ctx->owner[ctx->dst[e]] : -1;
          d.need = std::max<int64_t>(d.need, std::max(ctx->upload_end[e], ctx->upload_end[ctx->rev[e]]));
        }
        it16.push_back(d);
      }
    }
    if (!it16.empty()) {
      // equal-cost items: process them in the order in which a streamed upload (bpx_sweep_host) delivers their messages
      // (on a periodic lattice the wrap-around layers need the end of the message set and go last)
      std::stable_sort(it16.begin(), it16.end(), [](const onchip16::ItemDesc& a, const onchip16::ItemDesc& b) {
        return a.pair_mode != b.pair_mode ? a.pair_mode < b.pair_mode : a.need < b.need;
      });
      ctx->n_onchip16_items = (int)it16.size();
      cudaError_t e = cudaMalloc((void**)&ctx->d_onchip16_items, it16.size() * sizeof(onchip16::ItemDesc));
      if (e != cudaSuccess) {
        set_error(ctx, "cudaMalloc(on-chip 16 items) failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return BPX_ERR_ALLOC;
      }
      BPX_CUDA(ctx, cudaMemcpy(ctx->d_onchip16_items, it16.data(), it16.size() * sizeof(onchip16::ItemDesc), cudaMemcpyHostToDevice));
      BPX_CUDA(ctx, cudaFuncSetAttribute(onchip16::bp_update_onchip_c16, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)onchip16::SMEM_BYTES16));
      need_image = true;
    }
  }
  // ---- SLICED buckets (chi = 16, degree 4) ----
  // version 2 (bpx_sliced2.cuh, default): groups of G CTAs share one vertex; version 1 (BPX_SLICED_V1=1, or message
  // offsets that are not 16-byte aligned): two independent half items per vertex
  {
    auto F = [](auto*& p) {
      if (p) cudaFree(p);
      p = nullptr;
    };
    F(ctx->d_sliced2_items);
    F(ctx->d_sliced2_group_ptr);
    F(ctx->d_sliced2_partials);
    F(ctx->d_sliced2_part1);
    F(ctx->d_sliced2_gsync);
    ctx->n_sliced2_items = ctx->n_sliced2_groups = ctx->sliced2_G = ctx->sliced2_grid = 0;
    std::vector<sliced::ItemDesc> sit;
    std::vector<sliced2::VItem> vit;
    bool aligned = true;
    for (int i = 0; i < (int)ctx->buckets.size(); ++i) {
      Bucket& b = ctx->buckets[i];
      if (b.kernel != BPX_KERNEL_SLICED) continue;
      for (int32_t v : b.my_vertices) {
        sliced2::VItem w;
        memset(&w, 0, sizeof(w));
        w.site_off = ctx->dev_site_off[v];
        for (int br = 0; br < 2; ++br) {
          sliced::ItemDesc d;
          memset(&d, 0, sizeof(d));
          d.site_off = ctx->dev_site_off[v];
          d.branch = br;
          for (int l = 0; l < 4; ++l) {
            const int32_t e = ctx->out_edge[v][l];
            d.out_edge[l] = e;
            d.out_off[l] = ctx->msg_off[e];
            d.in_off[l] = ctx->msg_off[ctx->rev[e]];
            d.peer[l] = (!ctx->owner.empty() && ctx->owner[ctx->dst[e]] != ctx->rank) ? ctx->owner[ctx->dst[e]] : -1;
            d.need = std::max<int64_t>(d.need, std::max(ctx->upload_end[e], ctx->upload_end[ctx->rev[e]]));
            w.out_edge[l] = d.out_edge[l];
            w.out_off[l] = d.out_off[l];
            w.in_off[l] = d.in_off[l];
            w.peer[l] = d.peer[l];
            if (d.in_off[l] & 1) aligned = false;  // the staged message copies are TMA bulk loads (16-byte granules)
          }
          w.need = d.need;
          sit.push_back(d);
        }
        vit.push_back(w);
      }
    }
    if (!sit.empty()) {
      ctx->n_sliced_items = (int)sit.size();
      const bool v2 = aligned && !getenv("BPX_SLICED_V1") && ctx->num_sms >= 4;
      size_t scratch_doubles = (size_t)std::min(ctx->n_sliced_items, ctx->num_sms) * sliced::NTENSOR;
      std::vector<sliced2::VItem> grouped;
      std::vector<int32_t> group_ptr;
      if (v2) {
        int G = 8;
        if (const char* env = getenv("BPX_SLICED_G")) G = atoi(env) == 4 ? 4 : 8;
        if (ctx->num_sms < 8) G = 4;
        // capacity units of 4 CTAs: a group of 8 owns two units, a trailing group of 4 one; vertex i -> unit i % n_units
        // (vertices stay in lattice order inside a group: the order in which a streamed upload delivers their messages)
        const int n_units = ctx->num_sms / 4;
        const int upg = G / 4;
        const int n_groups = (n_units + upg - 1) / upg;
        std::vector<std::vector<int>> per_group(n_groups);
        for (size_t i = 0; i < vit.size(); ++i) per_group[(int)(i % n_units) / upg].push_back((int)i);
        group_ptr.push_back(0);
        for (auto& l : per_group) {
          for (int i : l) grouped.push_back(vit[i]);
          group_ptr.push_back((int32_t)grouped.size());
        }
        ctx->n_sliced2_items = (int)grouped.size();
        ctx->n_sliced2_groups = n_groups;
        ctx->sliced2_G = G;
        ctx->sliced2_grid = n_units * 4;
        scratch_doubles = (size_t)n_groups * 2 * sliced::NTENSOR;
      }
      cudaError_t e = cudaMalloc((void**)&ctx->d_sliced_items, sit.size() * sizeof(sliced::ItemDesc));
      if (e == cudaSuccess) e = cudaMalloc(&ctx->d_fast_scratch, scratch_doubles * sizeof(double));
      if (e == cudaSuccess && v2) e = cudaMalloc((void**)&ctx->d_sliced2_items, grouped.size() * sizeof(sliced2::VItem));
      if (e == cudaSuccess && v2) e = cudaMalloc((void**)&ctx->d_sliced2_group_ptr, group_ptr.size() * sizeof(int32_t));
      if (e == cudaSuccess && v2) e = cudaMalloc(&ctx->d_sliced2_partials, (size_t)ctx->n_sliced2_groups * sliced2::PART_PER_GROUP * sizeof(double));
      if (e == cudaSuccess && v2) e = cudaMalloc(&ctx->d_sliced2_part1, (size_t)ctx->sliced2_grid * sliced2::PART1_PER_CTA * sizeof(double));
      if (e == cudaSuccess && v2) e = cudaMalloc((void**)&ctx->d_sliced2_gsync, (size_t)ctx->n_sliced2_groups * sliced2::GS_STRIDE * sizeof(unsigned int));
      if (e != cudaSuccess) {
        set_error(ctx, "cudaMalloc(sliced kernel items/scratch) failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return BPX_ERR_ALLOC;
      }
      BPX_CUDA(ctx, cudaMemcpy(ctx->d_sliced_items, sit.data(), sit.size() * sizeof(sliced::ItemDesc), cudaMemcpyHostToDevice));
      BPX_CUDA(ctx, cudaFuncSetAttribute(sliced::bp_update_sliced_c16, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sliced::SMEM_BYTES));
      if (v2) {
        BPX_CUDA(ctx, cudaMemcpy(ctx->d_sliced2_items, grouped.data(), grouped.size() * sizeof(sliced2::VItem), cudaMemcpyHostToDevice));
        BPX_CUDA(ctx, cudaMemcpy(ctx->d_sliced2_group_ptr, group_ptr.data(), group_ptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        BPX_CUDA(ctx, cudaFuncSetAttribute(sliced2::bp_update_sliced_c16g<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sliced2::SMEM2_BYTES));
        BPX_CUDA(ctx, cudaFuncSetAttribute(sliced2::bp_update_sliced_c16g<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sliced2::SMEM2_BYTES));
      }
      need_image = true;
    }
  }
  if (!group.empty()) need_image = true;
  if (need_image) {
    // private pre-swizzled image of the site tensors (same offsets as the canonical buffer), refreshed lazily
    cudaError_t e = cudaMalloc(&ctx->d_sites_swz, std::max<size_t>(16, (size_t)ctx->dev_site_total * ctx->esize));
    if (e != cudaSuccess) {
      set_error(ctx, "cudaMalloc(pre-swizzled site image) failed: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return BPX_ERR_ALLOC;
    }
    ctx->sites_dirty = true;
  }
  if (ctx->n_sliced2_groups > 0) {
    // tensor-map views of the tensor image and of the groups' scratch images (both buffers exist now)
    sliced2::TensorMaps* tm = reinterpret_cast<sliced2::TensorMaps*>(ctx->sliced2_tmaps);
    int rc = sliced2_encode_views(ctx, ctx->d_sites_swz, (size_t)ctx->dev_site_total, &tm->a0h_sites, &tm->a3h_sites);
    if (!rc) rc = sliced2_encode_views(ctx, ctx->d_fast_scratch, (size_t)ctx->n_sliced2_groups * 2 * sliced::NTENSOR, &tm->a0h_scratch, &tm->a3h_scratch);
    if (rc) return rc;
  }
  if (group.empty()) return BPX_OK;
  std::sort(group.begin(), group.end(), [&](int a, int b) { return ctx->buckets[a].z > ctx->buckets[b].z; });
  std::vector<onchip::ItemDesc> items;
  for (int bi : group) {
    Bucket& b = ctx->buckets[bi];
    b.leader = group[0];
    for (int32_t v : b.my_vertices) {
      onchip::ItemDesc d;
      memset(&d, 0, sizeof(d));
      d.site_off = ctx->dev_site_off[v];
      d.z = b.z;
      for (int i = 0; i < b.z; ++i) {
        const int32_t e = ctx->out_edge[v][i];
        d.out_edge[i] = e;
        d.out_off[i] = ctx->msg_off[e];
        d.in_off[i] = ctx->msg_off[ctx->rev[e]];
        d.peer[i] = (!ctx->owner.empty() && ctx->owner[ctx->dst[e]] != ctx->rank) ? ctx->owner[ctx->dst[e]] : -1;
        d.need = std::max<int64_t>(d.need, std::max(ctx->upload_end[e], ctx->upload_end[ctx->rev[e]]));
      }
      for (int i = b.z; i < 4; ++i) d.peer[i] = -1;
      if (b.z == 4) {  // two half items (branch P, branch Q): finer granularity for the last wave
        d.branch = 0;
        items.push_back(d);
        d.branch = 1;
      }
      items.push_back(d);
    }
  }
  ctx->n_onchip_items = (int)items.size();
  cudaError_t e = cudaMalloc((void**)&ctx->d_onchip_items, items.size() * sizeof(onchip::ItemDesc));
  if (e != cudaSuccess) {
    set_error(ctx, "cudaMalloc(item descriptors) failed: %s", cudaGetErrorString(e));
    return BPX_ERR_ALLOC;
  }
  BPX_CUDA(ctx, cudaMemcpy(ctx->d_onchip_items, items.data(), items.size() * sizeof(onchip::ItemDesc), cudaMemcpyHostToDevice));
  BPX_CUDA(ctx, cudaFuncSetAttribute(onchip::bp_update_onchip_c8, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)onchip::SMEM_BYTES));
  return BPX_OK;
}

// re-derive kernel-private images of the site tensors after an upload
inline int fast_refresh_sites(bpx_ctx* ctx) {
  if (!ctx->sites_dirty) return BPX_OK;
  if (ctx->d_sites_swz && ctx->n_onchip_items > 0) {
    onchip::swizzle_sites<<<std::min(ctx->n_onchip_items, 4 * ctx->num_sms), 256, 0, ctx->stream>>>(
        (const onchip::ItemDesc*)ctx->d_onchip_items, ctx->n_onchip_items, (const double*)ctx->d_sites, (double*)ctx->d_sites_swz);
    ctx->n_launches++;
    BPX_CUDA(ctx, cudaGetLastError());
  }
  if (ctx->d_sites_swz && ctx->n_onchip16_items > 0) {
    onchip16::swizzle_sites_z3<<<std::min(ctx->n_onchip16_items, 8 * ctx->num_sms), 256, 0, ctx->stream>>>(
        (const onchip16::ItemDesc*)ctx->d_onchip16_items, ctx->n_onchip16_items, (const double*)ctx->d_sites, (double*)ctx->d_sites_swz);
    ctx->n_launches++;
    BPX_CUDA(ctx, cudaGetLastError());
  }
  if (ctx->d_img16c && ctx->n_onchip16c_slots > 0) {
    if (ctx->dtype == BPX_C64)
      onchip16c::swizzle_sites_c16<true><<<std::min(ctx->n_onchip16c_slots, 8 * ctx->num_sms), 256, 0, ctx->stream>>>(
          (const onchip16c::ItemDesc*)ctx->d_onchip16c_items, ctx->n_onchip16c_slots, (const double*)ctx->d_sites, (double*)ctx->d_img16c);
    else
      onchip16c::swizzle_sites_c16<false><<<std::min(ctx->n_onchip16c_slots, 8 * ctx->num_sms), 256, 0, ctx->stream>>>(
          (const onchip16c::ItemDesc*)ctx->d_onchip16c_items, ctx->n_onchip16c_slots, (const double*)ctx->d_sites, (double*)ctx->d_img16c);
    ctx->n_launches++;
    BPX_CUDA(ctx, cudaGetLastError());
  }
  if (ctx->d_img8c && ctx->n_onchip8c_slots > 0) {
    if (ctx->dtype == BPX_C64)
      onchip8c::swizzle_sites_c8<true><<<std::min(ctx->n_onchip8c_slots, 8 * ctx->num_sms), 256, 0, ctx->stream>>>(
          (const onchip8c::ItemDesc*)ctx->d_onchip8c_items, ctx->n_onchip8c_slots, (const double*)ctx->d_sites, (double*)ctx->d_img8c);
    else
      onchip8c::swizzle_sites_c8<false><<<std::min(ctx->n_onchip8c_slots, 8 * ctx->num_sms), 256, 0, ctx->stream>>>(
          (const onchip8c::ItemDesc*)ctx->d_onchip8c_items, ctx->n_onchip8c_slots, (const double*)ctx->d_sites, (double*)ctx->d_img8c);
    ctx->n_launches++;
    BPX_CUDA(ctx, cudaGetLastError());
  }
  if (ctx->d_sites_swz && ctx->n_sliced_items > 0) {
    sliced::swizzle_sites16<<<std::min(ctx->n_sliced_items, 8 * ctx->num_sms), 512, 0, ctx->stream>>>(
        (const sliced::ItemDesc*)ctx->d_sliced_items, ctx->n_sliced_items, (const double*)ctx->d_sites, (double*)ctx->d_sites_swz);
    ctx->n_launches++;
    BPX_CUDA(ctx, cudaGetLastError());
  }
  ctx->sites_dirty = false;
  return BPX_OK;
}

// SINGLE-mode buckets on the thread-per-vertex register kernel (bpx_vertex.cuh)
inline int launch_vertex_update(bpx_ctx* ctx, Bucket& b, const void* msg_in, void* msg_out, int normalize) {
  vertexk::Args k;
  k.site = b.d_vx_site;
  k.moff = b.d_vx_moff;
  k.sites = ctx->d_sites;
  k.msg_in = msg_in;
  k.msg_out = msg_out;
  k.resmax = ctx->cur_slot;
  k.n = (int64_t)b.my_vertices.size();
  k.normalize = normalize;
  k.out_contig = b.vx_out_contig;
  k.stop_key = ctx->stop_key;
  if (k.n == 0) return BPX_OK;
  const cudaError_t e = ctx->dtype == BPX_C64 ? vertexk::launch<c64>(k, b.z, b.chi, ctx->num_sms, ctx->stream)
                                              : vertexk::launch<double>(k, b.z, b.chi, ctx->num_sms, ctx->stream);
  ctx->n_launches++;
  BPX_CUDA(ctx, e);
  return BPX_OK;
}

inline int launch_fast_update(bpx_ctx* ctx, Bucket& b, const void* msg_in, void* msg_out, int normalize) {
  if (b.kernel == BPX_KERNEL_ONCHIP && uses_c8x(ctx, b)) {
    onchip8c::Args k;
    k.items = (const onchip8c::ItemDesc*)ctx->d_onchip8c_items;
    k.n_slots = ctx->n_onchip8c_slots;
    k.sites = (const double*)ctx->d_img8c;
    k.msg_in = (const double*)msg_in;
    k.msg_out = (double*)msg_out;
    k.resmax = ctx->cur_slot;
    k.normalize = normalize;
    k.peer = ctx->peer_args;
    k.io = ctx->io_args;
    k.stop_key = ctx->stop_key;
    if (ctx->onchip8c_grid == 0) return BPX_OK;
    if (ctx->dtype == BPX_C64)
      onchip8c::bp_update_onchip_c8x<true><<<ctx->onchip8c_grid, onchip8c::NT, onchip8c::SMEM_BYTES8C, ctx->stream>>>(k);
    else
      onchip8c::bp_update_onchip_c8x<false><<<ctx->onchip8c_grid, onchip8c::NT, onchip8c::SMEM_BYTES8C, ctx->stream>>>(k);
    ctx->n_launches++;
    BPX_CUDA(ctx, cudaGetLastError());
    return BPX_OK;
  }
  if (b.kernel == BPX_KERNEL_ONCHIP && uses_c16x(ctx, b)) {
    onchip16c::Args k;
    k.items = (const onchip16c::ItemDesc*)ctx->d_onchip16c_items;
    k.n_slots = ctx->n_onchip16c_slots;
    k.sites = (const double*)ctx->d_img16c;
    k.msg_in = (const double*)msg_in;
    k.msg_out = (double*)msg_out;
    k.resmax = ctx->cur_slot;
    k.normalize = normalize;
    k.peer = ctx->peer_args;
    k.io = ctx->io_args;
    k.stop_key = ctx->stop_key;
    k.timing = (long long*)ctx->d_timing;
    if (ctx->onchip16c_grid == 0) return BPX_OK;
    if (ctx->dtype == BPX_C64)
      onchip16c::bp_update_onchip_c16x<true><<<ctx->onchip16c_grid, onchip16c::NTHREADSC, onchip16c::SMEM_BYTES16C, ctx->stream>>>(k);
    else
      onchip16c::bp_update_onchip_c16x<false><<<ctx->onchip16c_grid, onchip16c::NTHREADSC, onchip16c::SMEM_BYTES16C, ctx->stream>>>(k);
    ctx->n_launches++;
    BPX_CUDA(ctx, cudaGetLastError());
    return BPX_OK;
  }
  if (b.kernel == BPX_KERNEL_ONCHIP && ctx->dtype == BPX_F64 && real_tuned_c16(b)) {
    onchip16::Args k;
    k.items = (const onchip16::ItemDesc*)ctx->d_onchip16_items;
    k.n_items = ctx->n_onchip16_items;
    k.sites = (const double*)ctx->d_sites_swz;
    k.msg_in = (const double*)msg_in;
    k.msg_out = (double*)msg_out;
    k.residual = nullptr;
    k.resmax = ctx->cur_slot;
    k.normalize = normalize;
    k.peer = ctx->peer_args;
    k.io = ctx->io_args;
    k.stop_key = ctx->stop_key;
    const int grid = std::min(k.n_items, ctx->num_sms);
    if (grid == 0) return BPX_OK;
    onchip16::bp_update_onchip_c16<<<grid, onchip16::NTHREADS16, onchip16::SMEM_BYTES16, ctx->stream>>>(k);
    ctx->n_launches++;
    BPX_CUDA(ctx, cudaGetLastError());
    return BPX_OK;
  }
  if (b.kernel == BPX_KERNEL_ONCHIP) {
    onchip::Args k;
    k.timing = (long long*)ctx->d_timing;
    k.items = (const onchip::ItemDesc*)ctx->d_onchip_items;
    k.n_items = ctx->n_onchip_items;
    k.sites = (const double*)ctx->d_sites_swz;
    k.msg_in = (const double*)msg_in;
    k.msg_out = (double*)msg_out;
    k.residual = nullptr;
    k.resmax = ctx->cur_slot;
    k.normalize = normalize;
    k.peer = ctx->peer_args;
    k.io = ctx->io_args;
    k.stop_key = ctx->stop_key;
    const int grid = std::min(k.n_items, ctx->num_sms);
    if (grid == 0) return BPX_OK;
    onchip::bp_update_onchip_c8<<<grid, onchip::NTHREADS, onchip::SMEM_BYTES, ctx->stream>>>(k);
    ctx->n_launches++;
    BPX_CUDA(ctx, cudaGetLastError());
    return BPX_OK;
  }
  if (b.kernel == BPX_KERNEL_SLICED && ctx->n_sliced2_groups > 0) {
    sliced2::Args k;
    k.items = (const sliced2::VItem*)ctx->d_sliced2_items;
    k.group_ptr = ctx->d_sliced2_group_ptr;
    k.n_groups = ctx->n_sliced2_groups;
    k.G = ctx->sliced2_G;
    k.sites = (const double*)ctx->d_sites_swz;
    k.scratch = (double*)ctx->d_fast_scratch;
    k.partials = (double*)ctx->d_sliced2_partials;
    k.part1 = (double*)ctx->d_sliced2_part1;
    k.gsync = ctx->d_sliced2_gsync;
    k.timing = (long long*)ctx->d_timing;
    k.flags = getenv("BPX_SLICED_FLAGS") ? atoi(getenv("BPX_SLICED_FLAGS")) : 0;
    k.msg_in = (const double*)msg_in;
    k.msg_out = (double*)msg_out;
    k.residual = nullptr;
    k.resmax = ctx->cur_slot;
    k.normalize = normalize;
    k.peer = ctx->peer_args;
    k.io = ctx->io_args;
    k.stop_key = ctx->stop_key;
    if (ctx->n_sliced2_items == 0) return BPX_OK;
    // the group counters start every launch at zero (a memset node when the step is captured into a CUDA graph)
    BPX_CUDA(ctx, cudaMemsetAsync(ctx->d_sliced2_gsync, 0, (size_t)ctx->n_sliced2_groups * sliced2::GS_STRIDE * sizeof(unsigned int), ctx->stream));
    // 8 compute warps (a column per warp and step, both products) is the default; BPX_SLICED_CW=16 selects the variant with
    // two warps per column (one product each, 96 registers): measured 9 % SLOWER at cfg5 -- the FP64 pipe, not the number of
    // warps that feed it, limits the compute phases
    const char* cw_env = getenv("BPX_SLICED_CW");
    if (cw_env && atoi(cw_env) == 16)
      sliced2::bp_update_sliced_c16g<16><<<ctx->sliced2_grid, sliced2::nthreads2<16>(), sliced2::SMEM2_BYTES, ctx->stream>>>(
          k, *reinterpret_cast<const sliced2::TensorMaps*>(ctx->sliced2_tmaps));
    else
      sliced2::bp_update_sliced_c16g<8><<<ctx->sliced2_grid, sliced2::nthreads2<8>(), sliced2::SMEM2_BYTES, ctx->stream>>>(
          k, *reinterpret_cast<const sliced2::TensorMaps*>(ctx->sliced2_tmaps));
    ctx->n_launches++;
    BPX_CUDA(ctx, cudaGetLastError());
    return BPX_OK;
  }
  if (b.kernel == BPX_KERNEL_SLICED) {
    sliced::Args k;
    k.items = (const sliced::ItemDesc*)ctx->d_sliced_items;
    k.n_items = ctx->n_sliced_items;
    k.sites = (const double*)ctx->d_sites_swz;
    k.scratch = (double*)ctx->d_fast_scratch;
    k.msg_in = (const double*)msg_in;
    k.msg_out = (double*)msg_out;
    k.residual = nullptr;
    k.resmax = ctx->cur_slot;
    k.normalize = normalize;
    k.peer = ctx->peer_args;
    k.io = ctx->io_args;
    k.stop_key = ctx->stop_key;
    const int grid = std::min(k.n_items, ctx->num_sms);
    if (grid == 0) return BPX_OK;
    sliced::bp_update_sliced_c16<<<grid, sliced::NTHREADS, sliced::SMEM_BYTES, ctx->stream>>>(k);
    ctx->n_launches++;
    BPX_CUDA(ctx, cudaGetLastError());
    return BPX_OK;
  }
  set_error(ctx, "no specialised kernel for this bucket");
  return BPX_ERR_UNSUPPORTED;
}

}  // namespace bpx
