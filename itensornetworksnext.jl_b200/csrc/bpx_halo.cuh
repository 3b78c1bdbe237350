// Multi-GPU halo: vertex partition, cut-edge message push over NVLink peer memory.
#pragma once
#include "bpx_ctx.h"

namespace bpx {
inline int halo_push(bpx_ctx*, void*) { return BPX_OK; }
}  // namespace bpx

extern "C" int bpx_set_partition(bpx_ctx* ctx, int, int, const int32_t*) {
  if (!ctx) return BPX_ERR_INVALID;
  bpx::set_error(ctx, "bpx_set_partition: not implemented yet");
  return BPX_ERR_UNSUPPORTED;
}
extern "C" int bpx_halo_export(bpx_ctx* ctx, void*) {
  if (!ctx) return BPX_ERR_INVALID;
  bpx::set_error(ctx, "bpx_halo_export: not implemented yet");
  return BPX_ERR_UNSUPPORTED;
}
extern "C" int bpx_halo_connect(bpx_ctx* ctx, int, const void*) {
  if (!ctx) return BPX_ERR_INVALID;
  bpx::set_error(ctx, "bpx_halo_connect: not implemented yet");
  return BPX_ERR_UNSUPPORTED;
}
extern "C" int64_t bpx_num_cut_edges(const bpx_ctx* ctx) { return ctx ? ctx->n_cut : -1; }
