// Specialised (bucket-specific) kernels: dispatch.  Kernels live in bpx_onchip.cuh / bpx_sliced.cuh.
#pragma once
#include "bpx_ctx.h"
#include "bpx_onchip.cuh"

namespace bpx {

inline bool fast_kernel_supported(bpx_ctx* ctx, const Bucket& b, int kernel) {
  if (kernel == BPX_KERNEL_GENERIC) return true;
  if (ctx->mode != BPX_MODE_NORM) return false;
  if (kernel == BPX_KERNEL_ONCHIP)
    return ctx->dtype == BPX_F64 && b.z == 4 && b.chi == 8 && b.d == 2 &&
           (size_t)ctx->max_smem_optin >= onchip::SMEM_BYTES;
  return false;
}

inline int fast_kernel_for(bpx_ctx* ctx, const Bucket& b) {
  if (fast_kernel_supported(ctx, b, BPX_KERNEL_ONCHIP)) return BPX_KERNEL_ONCHIP;
  return BPX_KERNEL_GENERIC;
}

inline int fast_prepare(bpx_ctx* ctx) {
  bool any = false;
  for (auto& b : ctx->buckets) any |= (b.kernel == BPX_KERNEL_ONCHIP);
  if (any)
    BPX_CUDA(ctx, cudaFuncSetAttribute(onchip::bp_update_onchip_z4c8, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)onchip::SMEM_BYTES));
  return BPX_OK;
}

inline int launch_fast_update(bpx_ctx* ctx, Bucket& b, const void* msg_in, void* msg_out, int normalize) {
  if (b.kernel == BPX_KERNEL_ONCHIP) {
    onchip::Args k;
    k.vdesc = ctx->d_vdesc;
    k.vertices = b.d_vertices;
    k.n_vertices = (int)b.my_vertices.size();
    k.msg_off = ctx->d_msg_off;
    k.sites = (const double*)ctx->d_sites;
    k.msg_in = (const double*)msg_in;
    k.msg_out = (double*)msg_out;
    k.residual = ctx->d_residual;
    k.normalize = normalize;
    const int grid = std::min(k.n_vertices, ctx->num_sms);
    if (grid == 0) return BPX_OK;
    onchip::bp_update_onchip_z4c8<<<grid, onchip::NTHREADS, onchip::SMEM_BYTES, ctx->stream>>>(k);
    ctx->n_launches++;
    BPX_CUDA(ctx, cudaGetLastError());
    return BPX_OK;
  }
  set_error(ctx, "no specialised kernel for this bucket");
  return BPX_ERR_UNSUPPORTED;
}

}  // namespace bpx
