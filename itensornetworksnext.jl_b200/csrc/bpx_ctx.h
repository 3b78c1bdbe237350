// Internal context layout of libbpx (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "bpx_common.cuh"
#include "bpx_peer.cuh"

namespace bpx {

struct Bucket {
  int z = 0, d = 1, chi = 0;          // chi == 0: link dims not uniform
  int max_dim = 0;                    // largest link dimension (all vertices of a bucket share the per-leg dims)
  std::vector<int32_t> vertices;      // all vertices of the bucket
  std::vector<int32_t> my_vertices;   // ... owned by this rank
  std::vector<int32_t> my_edges;      // out-edges of my_vertices (vertex-major, slot order)
  int32_t* d_vertices = nullptr;
  int32_t* d_edges = nullptr;
  int64_t* d_vx_site = nullptr;       // VERTEX kernel descriptors (bpx_vertex.cuh): site offsets [n] ...
  int32_t* d_vx_moff = nullptr;       // ... and message offsets [2 z][n], n = my_vertices.size()
  int vx_out_contig = 0;              // every vertex's out-messages are adjacent in the packed layout (slot order)
  int kernel = BPX_KERNEL_GENERIC;
  int leader = 0;  // bucket index whose launch covers this bucket (launch groups, bpx_fast.cuh)
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing;  // profiling: one event pair per timed launch
  double timed_ms = 0.0;
  int64_t timed_launches = 0;
};

struct Peer {
  int rank = -1;
  void* msg[2] = {nullptr, nullptr};  // peer's two message sets (cudaIpcOpenMemHandle, or plain pointers: same process)
  void* mailbox = nullptr;            // peer's mailbox array
  bool ipc = true;                    // opened from IPC handles (closed on release); false: bpx_create_multi siblings
};

}  // namespace bpx

struct bpx_ctx {
  int device = 0;
  int num_sms = 0;
  int max_smem_optin = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;

  // graph (host)
  bool graph_set = false, dims_set = false;
  int64_t nv = 0, ne = 0, n_und = 0;
  std::vector<int32_t> src, dst, slot, rev, deg;
  std::vector<std::vector<int32_t>> out_edge;  // per vertex, slot order
  std::vector<int32_t> link_dim, phys_dim;
  std::vector<int64_t> site_off, msg_off;  // packed HOST layouts (all vertices / edges)
  std::vector<int64_t> dev_site_off;       // device layout: owned vertices only, compacted (-1: not resident)
  int64_t dev_site_total = 0;
  std::vector<bpx::VDesc> h_vdesc;
  int dtype = BPX_F64, mode = BPX_MODE_NORM, esize = 8;
  int64_t max_site_elems = 1, max_msg_elems = 1;

  // device
  bpx::VDesc* d_vdesc = nullptr;
  int32_t *d_src = nullptr, *d_slot = nullptr, *d_rev = nullptr, *d_und_edge = nullptr;
  int32_t* d_first_out_edge = nullptr;  // per vertex: its first out-edge (-1: none), built on the first belief pass
  int64_t* d_msg_off = nullptr;
  void* d_sites = nullptr;
  void* d_msg[2] = {nullptr, nullptr};
  void* d_msg_snapshot = nullptr;
  int cur = 0;
  double* d_residual = nullptr;               // per-edge residual terms of generic-kernel updates
  unsigned long long* d_reskeys = nullptr;    // ring of per-sweep residual keys (residual_key), [history_cap + 1]
  unsigned long long* d_reskeys_local = nullptr;  // partitioned runs: this rank's own maxima
  unsigned long long* cur_slot = nullptr;     // where the kernels of the sweep being enqueued record
  int history_cap = 0, history_len = 0;       // history_len: sweeps recorded since the ring was last cleared
  int32_t* d_generic_edges = nullptr;         // owned edges updated by generic buckets (need bp_residual_max)
  int64_t n_generic_edges = 0;
  void* d_scratch = nullptr;
  int gen_smem_elems = 0, gen_smem_bytes = 0, gen_grid = 0;
  void* d_onchip16_items = nullptr;
  int n_onchip16_items = 0;
  void* d_sliced_items = nullptr;
  int n_sliced_items = 0;
  // group-cooperative version of the sliced kernel (bpx_sliced2.cuh): vertices grouped per CTA group
  void* d_sliced2_items = nullptr;
  int32_t* d_sliced2_group_ptr = nullptr;
  int n_sliced2_items = 0, n_sliced2_groups = 0, sliced2_G = 0, sliced2_grid = 0;
  void* d_sliced2_partials = nullptr;
  void* d_sliced2_part1 = nullptr;
  unsigned int* d_sliced2_gsync = nullptr;
  alignas(64) unsigned char sliced2_tmaps[4 * 128];  // sliced2::TensorMaps (four CUtensorMap), rebuilt with the buffers
  void* d_onchip16c_items = nullptr;  // complex chi = 16 kernel: items laid out as rounds (slot r * grid + cta)
  int n_onchip16c_slots = 0, onchip16c_grid = 0;
  void* d_onchip8c_items = nullptr;   // complex chi = 8 kernel, same round layout
  int n_onchip8c_slots = 0, onchip8c_grid = 0;
  void *d_img8c = nullptr, *d_img16c = nullptr;  // zero-padded private tensor images of the complex kernels
  void* d_timing = nullptr;  // debug: per-phase clock64 stamps (BPX_ONCHIP_TIMING builds)
  void* d_onchip_items = nullptr;
  void* d_sites_swz = nullptr;  // pre-swizzled site tensors for the ONCHIP kernel (bpx_onchip.cuh)
  bool sites_dirty = true;
  int n_onchip_items = 0;
  void* d_fast_scratch = nullptr;
  size_t fast_scratch_bytes = 0;

  std::vector<bpx::Bucket> buckets;
  int kernel_policy = BPX_KERNEL_AUTO;
  bool profiling = false;

  // partition
  int rank = 0, nranks = 1;
  std::vector<int32_t> owner;
  int32_t* d_owned_vertices = nullptr;
  int64_t n_owned_vertices = 0;
  int32_t* d_owned_edges = nullptr;
  int32_t* d_all_edges = nullptr;
  int64_t n_owned_edges = 0;
  std::vector<bpx::Peer> peers;
  void** d_peer_msg = nullptr;  // [2][nranks] device table of peer message-set pointers
  int32_t* d_cut = nullptr;     // (edge, peer) pairs of owned edges whose head lives on another rank
  int64_t n_cut = 0;
  void** d_peer_mailbox = nullptr;  // [nranks] device table of mailbox arrays (own entry included)
  void* d_mailbox = nullptr;        // this rank's mailbox array: one slot per source rank
  int* d_halo_error = nullptr;
  unsigned long long barrier_id = 0;
  unsigned long long sweep_id = 0;  // sweeps posted since bpx_set_partition (identical on all ranks)
  bool gate_pending = false, halo_connected = false;
  int gate_hist_idx = -1;
  bpx::PeerArgs peer_args = {};     // what the next fast launch is told about the exchange (nranks 0: nothing)
  unsigned int* d_ticket = nullptr;
  unsigned long long recv_mask = 0;  // ranks that own the tail of an edge pointing into this rank's block
  unsigned long long stop_key = 0;  // device-side convergence test of the sweep being enqueued (0: none)
  bool single_launch = false;       // the whole sweep of this rank is ONE fast launch: exchange fused into it

  // streamed host I/O (bpx_sweep_host with pinned buffers)
  cudaStream_t copy_stream = nullptr;
  cudaStream_t apply_stream = nullptr;  // second stream of the gate kernels (bond kernel of chunk i beside the side kernel of chunk i + 1)
  cudaEvent_t ev_io_start = nullptr, ev_io_done = nullptr;
  long long* d_io_progress = nullptr;
  long long* h_io_progress = nullptr;  // pinned: cumulative element counts, one per chunk; [32] residual key, [33] error flag
  bpx::HostIO io_args = {};            // what the next fast launch is told (all NULL: nothing)
  struct IoGraph {                     // one captured step per (host buffers, normalize, chunking)
    const void* in;
    void* out;
    int normalize, chunks;
    uint64_t epoch;
    int64_t launches;
    cudaGraphExec_t exec;
  };
  std::vector<IoGraph> io_graphs;
  // host <-> device transfers of an iterate move the messages this rank owns (all of them on a single rank): coalesced
  // runs of owned edges in edge order; upload_end[e] = elements of the upload (in that order) up to the end of message e
  // (0 for edges that are not uploaded -- cut edges into this rank arrive from the peers)
  std::vector<std::pair<int64_t, int64_t>> owned_runs;  // [element begin, element end)
  std::vector<int64_t> upload_end;
  int64_t owned_elems = 0;
  uint64_t work_epoch = 0;             // bumped whenever launch lists / buffers are rebuilt (invalidates io_graphs)
  bool io_graph_disabled = false;
  bool io_stream_disabled = false;     // a streamed step timed out once: this context stages its host steps from now on
  unsigned long long* slot_override = nullptr;  // residual slot of the step being enqueued (streamed steps)
  bool ring_dirty = false;             // residual ring slots were re-used without a clear (streamed steps)

  // counters
  int64_t n_launches = 0, n_updates = 0, n_sweeps = 0;
  int64_t n_gates_v3 = 0, n_gates_declined = 0;  // two-site gates applied by version 3 / handed on to versions 1, 2

  // grow-only work space of the belief / gate / expectation-value calls (one buffer per role; no cudaMalloc / cudaFree on
  // the call path once warm).  Buffers above WS_KEEP_BYTES (2 GiB) are released at the end of the call that needed them.
  enum { WS_SCALARS = 0, WS_EDGE_SCALARS, WS_LOGSUM, WS_OPS, WS_OP_OFF, WS_LIST, WS_DESC, WS_WORK, WS_SCRATCH, WS_SV, WS_OUT, WS_COUNT };
  void* ws[WS_COUNT] = {};
  size_t ws_bytes[WS_COUNT] = {};
  void* h_logsum = nullptr;  // pinned: result block of bpx_bethe_free_energy

  // single-process multi-GPU (bpx_create_multi, bpx_multi.cuh): the parent owns one partitioned child per device and holds
  // no device memory itself
  std::vector<bpx_ctx*> children;
  std::vector<int32_t> multi_owner;  // owner[v] = index of the child that updates the out-edges of v
  bool is_child = false;
  // internal zero-padding of link dims 9..15 to 16 (bpx_pad.cuh): this context is a thin parent that keeps the caller's
  // dims / packed layouts (link_dim, phys_dim, site_off, msg_off) and owns ONE child with the padded problem
  bool pad_active = false;
  std::vector<int> pad_devices;  // the caller created this context with bpx_create_multi: its device list (the child is a multi context)
  bool no_pad = false;  // never pad (children of multi-device / padding parents)
  // children only: element runs of the messages on cut edges that point INTO this device's block; a host iterate
  // (bpx_sweep_host) uploads them too, so that one call depends on its host buffer alone
  std::vector<std::pair<int64_t, int64_t>> halo_in_runs;
};

namespace bpx {

void set_error(bpx_ctx* ctx, const char* fmt, ...);

#define BPX_CUDA(ctx, call)                                                                       \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      bpx::set_error(ctx, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      cudaGetLastError();                                                                         \
      return BPX_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

int rebuild_work_lists(bpx_ctx* ctx);
int relayout_sites(bpx_ctx* ctx);  // (re)allocate the site buffer for the vertices this rank owns, keeping resident data
int launch_generic_update(bpx_ctx* ctx, const void* msg_in, void* msg_out, const int32_t* d_work, int64_t n_work,
                          int normalize, unsigned long long* resmax);
// specialised kernels (bpx_fast.cuh)
int fast_kernel_for(bpx_ctx* ctx, const Bucket& b);
bool fast_kernel_supported(bpx_ctx* ctx, const Bucket& b, int kernel);
int fast_prepare(bpx_ctx* ctx);
int fast_refresh_sites(bpx_ctx* ctx);
int launch_fast_update(bpx_ctx* ctx, Bucket& b, const void* msg_in, void* msg_out, int normalize);
int launch_vertex_update(bpx_ctx* ctx, Bucket& b, const void* msg_in, void* msg_out, int normalize);
// kernel families that do not implement the fused multi-GPU exchange / streamed host I/O hooks (the sweep runs the
// exchange as separate small kernels and stages host iterates for them)
inline bool plain_family(int kernel) { return kernel == BPX_KERNEL_GENERIC || kernel == BPX_KERNEL_VERTEX; }
// multi-GPU (bpx_halo.cuh)
int halo_push(bpx_ctx* ctx, void* msg_out);
int halo_post_residual(bpx_ctx* ctx);
int halo_gate(bpx_ctx* ctx);
void halo_release(bpx_ctx* ctx);
int halo_finalize(bpx_ctx* ctx);  // all peers known: build the device-side tables of peer message sets / mailboxes

}  // namespace bpx
