// Device-side pieces that let ONE persistent update kernel per sweep also do the multi-GPU exchange
// (bpx_halo.cuh describes the protocol): wait for the peers' posts of the previous sweep before the first
// cut-edge message is read, store new cut-edge messages straight into the owner's message set over NVLink from the
// epilogue, and let the last CTA to finish post (sweep id, local residual) into every peer's mailbox.
#pragma once
#include "bpx_common.cuh"

namespace bpx {

constexpr int MAILBOX_RING = 64;  // ranks are coupled through neighbours only, so they may be several sweeps apart
struct Mailbox {                  // one per source rank, lives in the RECEIVER's memory
  unsigned long long sweep_id;    // latest sweep the source rank has completed and pushed
  unsigned long long barrier_id;  // bpx_peer_barrier epochs
  double residual[MAILBOX_RING];  // its local residual maxima, indexed by sweep id % MAILBOX_RING
};

struct PeerArgs {
  int nranks;                              // <= 1: everything below is ignored
  int rank;
  Mailbox* my_mailbox;                     // [nranks]
  Mailbox* const* peer_mailbox;            // [nranks] (own entry included)
  unsigned long long wait_id;              // posts of this sweep id must have arrived before cut messages are read (0: none)
  unsigned long long wait_mask;            // bit p set: rank p sends this rank cut-edge messages (only those are awaited)
  unsigned long long post_id;              // id to post when this kernel's updates are complete (0: do not post)
  unsigned long long* prev_global_key;     // where CTA 0 folds the previous sweep's global residual (may be NULL)
  const unsigned long long* local_key;     // this sweep's local residual key (the kernels' resmax slot)
  double* const* peer_out;                 // [nranks] message set being written this sweep, per rank
  unsigned int* ticket;                    // CTA completion counter (self-resetting)
  int* error_flag;
};

// ---- streamed host I/O (bpx_sweep_host with pinned buffers): the sweep kernel runs WHILE the copy engine uploads
// the iterate in chunks -- an item waits until the messages it reads have arrived -- and every new message is also
// stored straight into the caller's (device-mapped) host buffer, so no download follows the kernel.
struct HostIO {
  long long* progress;        // elements of the message set that have arrived in msg_in so far (NULL: all resident);
                              // reset by the kernel's last CTA for the next step
  double* host_out;           // device-accessible alias of the host output buffer, same offsets as msg_out (NULL: none)
  int* error_flag;
  unsigned long long* host_key;           // where the last CTA stores the sweep's residual key (mapped host memory)
  unsigned long long* local_key;          // the kernels' resmax slot for this step (self-resetting)
  unsigned long long* ring_key;           // residual-history slot that receives a copy
  unsigned int* ticket;                   // CTA completion counter (self-resetting)
};
// Called by every thread at the end of the kernel (CTA-uniform): the last CTA to finish hands the sweep's residual key
// to the host, so the step needs no device-to-host copy at all.
__device__ __forceinline__ void hostio_finish(const HostIO& io) {
  if (!io.progress || !io.ticket) return;  // (only the LAST launch of a sweep is given the ticket)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(io.ticket, 1u) == gridDim.x - 1) {
      __threadfence();
      if (io.host_key) {
        const unsigned long long key = *reinterpret_cast<const volatile unsigned long long*>(io.local_key);
        *reinterpret_cast<volatile unsigned long long*>(io.host_key) = key;
        *io.ring_key = key;
        *io.local_key = 0ull;
      }
      // ready for the next step (stream-ordered): no memset nodes in front of the kernel.  Every item has passed its
      // gate, so the upload -- and with it the last write of the progress word -- is complete.
      *io.progress = 0ll;
      *io.ticket = 0u;
    }
  }
}

// Called by ONE warp per CTA before it touches messages written by peers.  Lane p waits for rank p's post.
// the calling warp waits (lane 0 spins, bounded, with back-off) until `need` elements of the upload have arrived.
// Keep the number of polling warps small: a thousand lanes hammering the one L2 line starve the copy engine's write of
// the progress word itself (observed: a sweep whose every item needs the whole upload never saw it arrive).
__device__ __forceinline__ void hostio_wait(const HostIO& io, long long need) {
  if (!io.progress) return;
  if ((threadIdx.x & 31) == 0) {
    const volatile long long* p = io.progress;  // p[0] progress, p[3] abort (same L2 line)
    const long long t0 = clock64();
    unsigned ns = 200;
    while (*p < need && p[3] == 0) {
      if (clock64() - t0 > 8000000000ll) {  // ~4 s: give up once for the whole grid instead of hanging the GPU; the host
        io.progress[3] = 1;                 // then repeats the step with a staged upload (e.g. under a profiler that
        if (io.error_flag) *reinterpret_cast<volatile int*>(io.error_flag) = 2;  // serialises kernel and copies)
        break;
      }
      __nanosleep(ns);
      if (ns < 2000) ns += ns;
    }
  }
  __syncwarp();
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by ONE warp per CTA before it touches messages written by peers.  Lane p waits for rank p's post.
__device__ __forceinline__ void peer_gate(const PeerArgs& pa, int lane) {
  if (pa.nranks <= 1 || pa.wait_id == 0) return;
  double v = -INFINITY;
  const bool fold = blockIdx.x == 0 && pa.prev_global_key != nullptr;  // folding the global residual needs every rank
  for (int p = lane; p < pa.nranks; p += 32) {
    if (!fold && !((pa.wait_mask >> p) & 1ull)) continue;
    const unsigned long long* flag = &pa.my_mailbox[p].sweep_id;
    const long long t0 = clock64();
    bool ok = true;
    // acquire loads instead of a trailing fence.sys (measured: ~4 us per system fence, on every CTA's critical path):
    // the peer's message stores are ordered before its release of the flag
    while (ld_acquire_sys(flag) < pa.wait_id) {
      if (clock64() - t0 > 8000000000ll) {  // ~4 s: raise the error flag instead of hanging the GPU
        ok = false;
        break;
      }
      __nanosleep(20);
    }
    if (ok) {
      const double r = *reinterpret_cast<volatile double*>(&pa.my_mailbox[p].residual[pa.wait_id % MAILBOX_RING]);
      v = (r != r || v != v) ? NAN : fmax(v, r);
    } else {
      atomicExch(pa.error_flag, 1);
    }
  }
  asm volatile("fence.proxy.async;\n" ::: "memory");  // the messages are read through TMA (async proxy) as well
  if (fold) {
    bool has_nan = v != v;
    has_nan = __any_sync(0xffffffffu, has_nan);
    double m = has_nan ? -INFINITY : v;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
    if (lane == 0) *pa.prev_global_key = residual_key(has_nan ? NAN : m);
  }
  __syncwarp();
}

// Called by every thread of every CTA when all of the CTA's global/peer stores have been issued (CTA-uniform).
// `wrote`: this thread has peer stores that it did not release yet.  The fused kernels release (fence.sys) right
// after the peer stores of a cut edge -- in an epilogue warp, off the critical path -- and pass false: a system fence
// costs ~4 us (measured) and would sit on the kernel's tail here.  The last CTA to arrive posts (post_id, local
// residual) to every rank.
__device__ __forceinline__ void peer_post_when_last(const PeerArgs& pa, bool wrote = true) {
  if (pa.nranks <= 1 || pa.post_id == 0) return;
  __shared__ unsigned int s_last;
  if (wrote) __threadfence_system();  // release this thread's message / residual-key stores
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(pa.ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x < pa.nranks) {
      // every storing thread of the grid fenced (system scope) before its CTA took a ticket; the release store orders
      // those stores and the residual before the flag -- one fence-equivalent instead of two full system fences
      Mailbox* mb = pa.peer_mailbox[threadIdx.x] + pa.rank;
      mb->residual[pa.post_id % MAILBOX_RING] = residual_from_key(*reinterpret_cast<const volatile unsigned long long*>(pa.local_key));
      st_release_sys(&mb->sweep_id, pa.post_id);
    }
    if (threadIdx.x == 0) *pa.ticket = 0u;  // ready for the next launch (stream-ordered)
  }
}

}  // namespace bpx
