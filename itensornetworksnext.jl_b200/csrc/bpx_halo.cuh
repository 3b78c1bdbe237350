// Multi-GPU: vertex partition, one context (process) per GPU.
//
// Under the synchronous schedule an update of edge (u -> v) reads only previous-sweep messages INTO u, so a
// rank that owns u needs, besides its own messages, the messages on cut edges that point into its block.
// Per sweep and rank:
//   1. update kernels for the owned vertices (bucket launches, unchanged),
//   2. halo_push_kernel: copy every owned message whose head lives on another rank straight into THAT
//      rank's message set over NVLink (peer pointers from cudaIpcOpenMemHandle; st.global to peer memory),
//   3. local residual max, then residual_post_kernel writes (sweep id, local max) into the mailbox of
//      every peer with a system-scope release,
//   4. before the next sweep, residual_gate_kernel spins (bounded) until every peer's mailbox entry carries
//      the current sweep id, and folds the maxima: that is the cross-rank barrier AND the all-reduce(MAX)
//      of the convergence residual, in one hop over NVLink -- no NCCL on the data path.
// Both message sets are IPC-exported; all ranks flip them in lockstep (bpx_set_messages resets parity).
#pragma once
#include <cstring>

#include "bpx_ctx.h"
#include "bpx_peer.cuh"

namespace bpx {

// warp per cut edge: local out-message -> peer's out buffer (same offset)
__global__ void halo_push_kernel(const int32_t* __restrict__ cut /* (edge, peer) pairs */, int64_t n_cut,
                                 const int64_t* __restrict__ msg_off, const double* __restrict__ local_out, double* const* peer_out,
                                 int doubles_per_elem) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_cut) return;
  const int e = cut[2 * w], peer = cut[2 * w + 1];
  const int64_t off = msg_off[e] * doubles_per_elem, n = (msg_off[e + 1] - msg_off[e]) * doubles_per_elem;
  const double* src = local_out + off;
  double* dst = peer_out[peer] + off;
  for (int64_t i = lane; i < n; i += 32) dst[i] = src[i];
}

__global__ void residual_post_kernel(const unsigned long long* __restrict__ local_key, Mailbox* const* peer_mailbox, int rank, int nranks,
                                     unsigned long long sweep_id) {
  const int p = threadIdx.x;
  if (p >= nranks) return;
  // the kernels before this one (updates, halo push) completed in stream order; make their peer writes visible
  // system-wide before the flag
  __threadfence_system();
  Mailbox* mb = peer_mailbox[p] + rank;
  mb->residual[sweep_id % MAILBOX_RING] = residual_from_key(*local_key);
  __threadfence_system();
  *reinterpret_cast<volatile unsigned long long*>(&mb->sweep_id) = sweep_id;
}

// One thread per source rank waits for that rank's post of `sweep_id`; then the maxima are folded.
// The spin is bounded (~4 s of SM clock): on time-out the error flag is raised instead of hanging the GPU.
__global__ void residual_gate_kernel(Mailbox* my_mailbox, int nranks, unsigned long long sweep_id, unsigned long long* global_key,
                                     int* error_flag) {
  __shared__ double vals[64];
  const int p = threadIdx.x;
  double v = -INFINITY;
  if (p < nranks) {
    volatile unsigned long long* flag = &my_mailbox[p].sweep_id;
    const long long t0 = clock64();
    bool ok = true;
    while (*flag < sweep_id) {
      if (clock64() - t0 > 8000000000ll) {
        ok = false;
        break;
      }
      __nanosleep(200);
    }
    __threadfence_system();
    if (ok)
      v = *reinterpret_cast<volatile double*>(&my_mailbox[p].residual[sweep_id % MAILBOX_RING]);
    else
      atomicExch(error_flag, 1);
  }
  if (p < 64) vals[p] = v;
  __syncthreads();
  if (p == 0) {
    double m = -INFINITY;
    bool has_nan = false;
    for (int i = 0; i < nranks; ++i) {
      if (vals[i] != vals[i]) has_nan = true;
      m = fmax(m, vals[i]);
    }
    if (has_nan) m = nan("");
    if (global_key) *global_key = residual_key(m);
  }
}

// device-side barrier over all ranks: announce epoch `id` in every peer's mailbox, then wait for everybody's
__global__ void peer_barrier_kernel(Mailbox* my_mailbox, Mailbox* const* peer_mailbox, int rank, int nranks, unsigned long long id,
                                    int* error_flag) {
  const int p = threadIdx.x;
  if (p < nranks) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(&(peer_mailbox[p] + rank)->barrier_id) = id;
    volatile unsigned long long* flag = &my_mailbox[p].barrier_id;
    const long long t0 = clock64();
    while (*flag < id) {
      if (clock64() - t0 > 8000000000ll) {
        atomicExch(error_flag, 1);
        break;
      }
      __nanosleep(100);
    }
    __threadfence_system();
  }
}

inline int halo_push(bpx_ctx* ctx, void* msg_out) {
  if (ctx->nranks <= 1 || ctx->n_cut == 0) return BPX_OK;
  if (!ctx->halo_connected) {
    set_error(ctx, "partitioned sweep before bpx_halo_connect() was called for every peer");
    return BPX_ERR_INVALID;
  }
  const int threads = 256;
  const int64_t blocks = (ctx->n_cut * 32 + threads - 1) / threads;
  double* const* peer_out = reinterpret_cast<double* const*>(ctx->d_peer_msg) + (size_t)(ctx->cur ^ 1) * ctx->nranks;
  halo_push_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(ctx->d_cut, ctx->n_cut, ctx->d_msg_off, (const double*)msg_out, peer_out,
                                                                 ctx->esize / 8);
  ctx->n_launches++;
  BPX_CUDA(ctx, cudaGetLastError());
  return BPX_OK;
}

// after the local residual max of sweep `sweep_id`: tell every rank (including ourselves)
inline int halo_post_residual(bpx_ctx* ctx) {
  if (ctx->nranks <= 1) return BPX_OK;
  ctx->sweep_id++;
  residual_post_kernel<<<1, 64, 0, ctx->stream>>>(ctx->d_reskeys_local + ctx->gate_hist_idx, reinterpret_cast<Mailbox* const*>(ctx->d_peer_mailbox),
                                                 ctx->rank, ctx->nranks, ctx->sweep_id);
  ctx->n_launches++;
  ctx->gate_pending = true;
  BPX_CUDA(ctx, cudaGetLastError());
  return BPX_OK;
}

// wait for every rank's post of the last sweep; d_resmax[0] (and the history slot) receive the global max
inline int halo_gate(bpx_ctx* ctx) {
  if (ctx->nranks <= 1 || !ctx->gate_pending) return BPX_OK;
  residual_gate_kernel<<<1, 64, 0, ctx->stream>>>(reinterpret_cast<Mailbox*>(ctx->d_mailbox), ctx->nranks, ctx->sweep_id,
                                                 ctx->gate_hist_idx >= 0 ? ctx->d_reskeys + ctx->gate_hist_idx : nullptr, ctx->d_halo_error);
  ctx->n_launches++;
  ctx->gate_pending = false;
  BPX_CUDA(ctx, cudaGetLastError());
  return BPX_OK;
}

inline int halo_finalize(bpx_ctx* ctx) {
  if ((int)ctx->peers.size() != ctx->nranks - 1) {
    set_error(ctx, "halo: %d of %d peers connected", (int)ctx->peers.size(), ctx->nranks - 1);
    return BPX_ERR_INVALID;
  }
  std::vector<void*> msg(2 * ctx->nranks, nullptr), mb(ctx->nranks, nullptr);
  msg[ctx->rank] = ctx->d_msg[0];
  msg[ctx->nranks + ctx->rank] = ctx->d_msg[1];
  mb[ctx->rank] = ctx->d_mailbox;
  for (auto& q : ctx->peers) {
    msg[q.rank] = q.msg[0];
    msg[ctx->nranks + q.rank] = q.msg[1];
    mb[q.rank] = q.mailbox;
  }
  BPX_CUDA(ctx, cudaMalloc((void**)&ctx->d_peer_msg, msg.size() * sizeof(void*)));
  BPX_CUDA(ctx, cudaMemcpy(ctx->d_peer_msg, msg.data(), msg.size() * sizeof(void*), cudaMemcpyHostToDevice));
  BPX_CUDA(ctx, cudaMalloc((void**)&ctx->d_peer_mailbox, mb.size() * sizeof(void*)));
  BPX_CUDA(ctx, cudaMemcpy(ctx->d_peer_mailbox, mb.data(), mb.size() * sizeof(void*), cudaMemcpyHostToDevice));
  ctx->halo_connected = true;
  return BPX_OK;
}

inline void halo_release(bpx_ctx* ctx) {
  for (auto& p : ctx->peers) {
    if (!p.ipc) continue;  // siblings of a multi-device context: plain pointers owned by the sibling
    for (int k = 0; k < 2; ++k)
      if (p.msg[k] && p.rank != ctx->rank) cudaIpcCloseMemHandle(p.msg[k]);
    if (p.mailbox && p.rank != ctx->rank) cudaIpcCloseMemHandle(p.mailbox);
  }
  ctx->peers.clear();
  auto F = [](auto*& p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  F(ctx->d_peer_msg);
  F(ctx->d_peer_mailbox);
  F(ctx->d_cut);
  F(ctx->d_mailbox);
  F(ctx->d_halo_error);
  F(ctx->d_ticket);
  ctx->peer_args = bpx::PeerArgs{};
  ctx->n_cut = 0;
  ctx->halo_connected = false;
  ctx->gate_pending = false;
}

}  // namespace bpx

extern "C" int bpx_set_partition(bpx_ctx* ctx, int rank, int nranks, const int32_t* owner) {
  if (ctx && ctx->pad_active) ctx = ctx->children[0];  // zero-padded problem (bpx_pad.cuh): the child is the partitioned context
  if (!ctx) return BPX_ERR_INVALID;
  if (!ctx->children.empty()) {
    bpx::set_error(ctx, "multi-device contexts partition themselves (bpx_set_owner); the per-rank halo calls do not apply");
    return BPX_ERR_INVALID;
  }
  if (!ctx->dims_set) {
    bpx::set_error(ctx, "bpx_set_partition: call bpx_set_dims first");
    return BPX_ERR_INVALID;
  }
  if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks || (nranks > 1 && !owner)) {
    bpx::set_error(ctx, "bpx_set_partition: bad rank/nranks (1 <= nranks <= 64)");
    return BPX_ERR_INVALID;
  }
  BPX_CUDA(ctx, cudaSetDevice(ctx->device));
  BPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  bpx::halo_release(ctx);
  ctx->rank = rank;
  ctx->nranks = nranks;
  ctx->owner.clear();
  if (nranks > 1) {
    ctx->owner.assign(owner, owner + ctx->nv);
    for (int64_t v = 0; v < ctx->nv; ++v)
      if (owner[v] < 0 || owner[v] >= nranks) {
        bpx::set_error(ctx, "bpx_set_partition: owner[%lld] = %d out of range", (long long)v, owner[v]);
        ctx->owner.clear();
        ctx->nranks = 1;
        ctx->rank = 0;
        return BPX_ERR_INVALID;
      }
    // cut edges: owned edge (u -> v) whose head v lives elsewhere -> deliver to owner[v]
    std::vector<int32_t> cut;
    for (int64_t e = 0; e < ctx->ne; ++e)
      if (owner[ctx->src[e]] == rank && owner[ctx->dst[e]] != rank) {
        cut.push_back((int32_t)e);
        cut.push_back(owner[ctx->dst[e]]);
      }
    ctx->n_cut = (int64_t)cut.size() / 2;
    ctx->recv_mask = 0;
    for (int64_t e = 0; e < ctx->ne; ++e)
      if (owner[ctx->dst[e]] == rank && owner[ctx->src[e]] != rank) ctx->recv_mask |= 1ull << owner[ctx->src[e]];
    if (!cut.empty()) {
      BPX_CUDA(ctx, cudaMalloc((void**)&ctx->d_cut, cut.size() * sizeof(int32_t)));
      BPX_CUDA(ctx, cudaMemcpy(ctx->d_cut, cut.data(), cut.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    BPX_CUDA(ctx, cudaMalloc(&ctx->d_mailbox, 64 * sizeof(bpx::Mailbox)));
    BPX_CUDA(ctx, cudaMemset(ctx->d_mailbox, 0, 64 * sizeof(bpx::Mailbox)));
    BPX_CUDA(ctx, cudaMalloc((void**)&ctx->d_halo_error, sizeof(int)));
    BPX_CUDA(ctx, cudaMemset(ctx->d_halo_error, 0, sizeof(int)));
    BPX_CUDA(ctx, cudaMalloc((void**)&ctx->d_ticket, sizeof(unsigned int)));
    BPX_CUDA(ctx, cudaMemset(ctx->d_ticket, 0, sizeof(unsigned int)));
    ctx->sweep_id = 0;
    ctx->barrier_id = 0;
  }
  {
    int rc = bpx::relayout_sites(ctx);  // site tensors are stored on the owning rank only
    if (rc) return rc;
  }
  return bpx::rebuild_work_lists(ctx);
}

// handles: [0..63] message set 0, [64..127] message set 1, [128..191] mailbox
extern "C" int bpx_halo_export(bpx_ctx* ctx, void* handles_3x64) {
  if (ctx && ctx->pad_active) ctx = ctx->children[0];  // zero-padded problem (bpx_pad.cuh): the child is the partitioned context
  if (!ctx || !handles_3x64) return BPX_ERR_INVALID;
  if (!ctx->children.empty()) {
    bpx::set_error(ctx, "multi-device contexts partition themselves (bpx_set_owner); the per-rank halo calls do not apply");
    return BPX_ERR_INVALID;
  }
  if (!ctx->dims_set || ctx->nranks <= 1) {
    bpx::set_error(ctx, "bpx_halo_export: call bpx_set_partition (nranks > 1) first");
    return BPX_ERR_INVALID;
  }
  BPX_CUDA(ctx, cudaSetDevice(ctx->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h[3];
  BPX_CUDA(ctx, cudaIpcGetMemHandle(&h[0], ctx->d_msg[0]));
  BPX_CUDA(ctx, cudaIpcGetMemHandle(&h[1], ctx->d_msg[1]));
  BPX_CUDA(ctx, cudaIpcGetMemHandle(&h[2], ctx->d_mailbox));
  memcpy(handles_3x64, h, sizeof(h));
  return BPX_OK;
}

// Connect peer `peer_rank` (handles from ITS bpx_halo_export).  Must be called for every rank != own rank;
// the last call finalises the device-side peer tables.
extern "C" int bpx_halo_connect(bpx_ctx* ctx, int peer_rank, const void* handles_3x64) {
  if (ctx && ctx->pad_active) ctx = ctx->children[0];  // zero-padded problem (bpx_pad.cuh): the child is the partitioned context
  if (!ctx || !handles_3x64) return BPX_ERR_INVALID;
  if (!ctx->children.empty()) {
    bpx::set_error(ctx, "multi-device contexts partition themselves (bpx_set_owner); the per-rank halo calls do not apply");
    return BPX_ERR_INVALID;
  }
  if (!ctx->dims_set || ctx->nranks <= 1 || peer_rank < 0 || peer_rank >= ctx->nranks || peer_rank == ctx->rank) {
    bpx::set_error(ctx, "bpx_halo_connect: bad peer rank %d", peer_rank);
    return BPX_ERR_INVALID;
  }
  BPX_CUDA(ctx, cudaSetDevice(ctx->device));
  for (auto& p : ctx->peers)
    if (p.rank == peer_rank) {
      bpx::set_error(ctx, "bpx_halo_connect: peer %d already connected", peer_rank);
      return BPX_ERR_INVALID;
    }
  cudaIpcMemHandle_t h[3];
  memcpy(h, handles_3x64, sizeof(h));
  bpx::Peer p;
  p.rank = peer_rank;
  BPX_CUDA(ctx, cudaIpcOpenMemHandle(&p.msg[0], h[0], cudaIpcMemLazyEnablePeerAccess));
  BPX_CUDA(ctx, cudaIpcOpenMemHandle(&p.msg[1], h[1], cudaIpcMemLazyEnablePeerAccess));
  BPX_CUDA(ctx, cudaIpcOpenMemHandle(&p.mailbox, h[2], cudaIpcMemLazyEnablePeerAccess));
  ctx->peers.push_back(p);
  if ((int)ctx->peers.size() == ctx->nranks - 1) return bpx::halo_finalize(ctx);
  return BPX_OK;
}

extern "C" int64_t bpx_num_cut_edges(const bpx_ctx* ctx) {
  if (ctx && ctx->pad_active) ctx = ctx->children[0];
  if (ctx && !ctx->children.empty()) {
    int64_t n = 0;
    for (const bpx_ctx* c : ctx->children) n += c->n_cut;
    return n;
  }
  return ctx ? ctx->n_cut : -1;
}

// Enqueue a cross-rank barrier on the context's stream (every rank must call it the same number of times).
extern "C" int bpx_peer_barrier(bpx_ctx* ctx) {
  if (ctx && ctx->pad_active) ctx = ctx->children[0];  // zero-padded problem (bpx_pad.cuh): the child is the partitioned context
  if (!ctx) return BPX_ERR_INVALID;
  if (!ctx->children.empty()) {
    for (bpx_ctx* c : ctx->children) {
      const int rc = bpx_peer_barrier(c);
      if (rc) {
        ctx->err = c->err;
        return rc;
      }
    }
    return BPX_OK;
  }
  if (ctx->nranks <= 1) return BPX_OK;
  if (!ctx->halo_connected) {
    bpx::set_error(ctx, "bpx_peer_barrier: peers are not connected");
    return BPX_ERR_INVALID;
  }
  BPX_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->barrier_id++;
  bpx::peer_barrier_kernel<<<1, 64, 0, ctx->stream>>>(reinterpret_cast<bpx::Mailbox*>(ctx->d_mailbox),
                                                    reinterpret_cast<bpx::Mailbox* const*>(ctx->d_peer_mailbox), ctx->rank, ctx->nranks,
                                                    ctx->barrier_id, ctx->d_halo_error);
  ctx->n_launches++;
  BPX_CUDA(ctx, cudaGetLastError());
  return BPX_OK;
}
