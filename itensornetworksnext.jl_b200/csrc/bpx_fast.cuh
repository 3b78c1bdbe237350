// Specialised (bucket-specific) kernels.  Dispatch lives here; kernels in bpx_onchip.cuh / bpx_sliced.cuh.
#pragma once
#include "bpx_ctx.h"

namespace bpx {

inline int fast_kernel_for(bpx_ctx*, const Bucket&) { return BPX_KERNEL_GENERIC; }
inline bool fast_kernel_supported(bpx_ctx*, const Bucket&, int kernel) { return kernel == BPX_KERNEL_GENERIC; }
inline int fast_prepare(bpx_ctx*) { return BPX_OK; }
inline int launch_fast_update(bpx_ctx* ctx, Bucket&, const void*, void*, int) {
  set_error(ctx, "no specialised kernel for this bucket");
  return BPX_ERR_UNSUPPORTED;
}

}  // namespace bpx
