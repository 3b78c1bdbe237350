/*
 * bpx.h -- C ABI of libbpx.so: B200-native (sm_100a) belief-propagation message updates for
 * ITensorNetworksNext.jl's `beliefpropagation` hot path.
 *
 * The reference (pure Julia, /root/reference) has no FFI for this path; its extension point is
 *   "Plug in a new strategy by subtyping `MessageUpdateAlgorithm` and overloading `message_update!`"
 *   (src/beliefpropagation/beliefpropagation.jl:214-220), reached through the
 *   `message_update_algorithm=` keyword of `beliefpropagation` (:69-92) and `select_algorithm`
 *   (src/select_algorithm.jl:41-48).
 * The entry points below are what a Julia `ccall` glue for that strategy binds (julia/BPX.jl,
 * INTEGRATION.md).  Each one cites the reference code it replaces.
 *
 * Conventions
 *   - every function returns BPX_OK (0) or a negative bpx_status; the message of the last failure on
 *     a context is available from bpx_last_error().  No C++ exception crosses this boundary.
 *   - plain pointers and sizes only.  Host buffers are owned by the caller and are only touched during
 *     the call (all calls are synchronous w.r.t. the host buffers they read or write).
 *   - the library owns all device memory.  A context made by bpx_create is bound to ONE CUDA device; multi-GPU runs
 *     either use ONE context over a device list (bpx_create_multi: single process, what a `beliefpropagation()` call
 *     needs) or one context per rank/process connected with bpx_halo_* (what torchrun / MPI launchers need).
 *   - a context is not thread-safe; several contexts may coexist.
 *   - there is no CPU fallback: every compute entry point fails with BPX_ERR_CUDA if no sm_100 device
 *     is usable.
 *
 * Canonical data layout at the boundary (column-major, first index fastest, like Julia arrays)
 *   graph      vertices 0..nv-1; DIRECTED edges 0..ne-1, both orientations of every link present.
 *              slot[e] = position of link e among the link legs of src[e]  (0..deg(src[e])-1).
 *   site       NORM mode:   A_v[s, l_0, ..., l_{z-1}]  (physical leg fastest, link legs in slot order)
 *              SINGLE mode: T_v[l_0, ..., l_{z-1}]
 *              packed for all vertices in vertex order (bpx_site_offset gives element offsets).
 *   message    NORM mode:   M_e[bra, ket], chi_e x chi_e (src/beliefpropagation/messagecache.jl:205-225)
 *              SINGLE mode: m_e[chi_e]
 *              packed for all directed edges in edge order (bpx_message_offset).
 *   element    BPX_F64: double;  BPX_C64: interleaved (re, im) doubles == Julia ComplexF64.
 */
#ifndef BPX_H
#define BPX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPX_VERSION 100
#define BPX_MAX_DEGREE 12

typedef struct bpx_ctx bpx_ctx;

typedef enum {
  BPX_OK = 0,
  BPX_ERR_INVALID = -1, /* bad argument / call order */
  BPX_ERR_CUDA = -2,    /* CUDA runtime failure or no usable device */
  BPX_ERR_ALLOC = -3,   /* host or device allocation failed */
  BPX_ERR_UNSUPPORTED = -4
} bpx_status;

enum { BPX_F64 = 0, BPX_C64 = 1 };
enum { BPX_MODE_NORM = 0, BPX_MODE_SINGLE = 1 };

/* Kernel families a bucket of updates can be routed to (bpx_bucket_info / bpx_set_kernel_policy). */
enum {
  BPX_KERNEL_AUTO = 0,
  BPX_KERNEL_GENERIC = 1, /* edge-centric, any shape, strided loops            */
  BPX_KERNEL_ONCHIP = 2,  /* vertex-centric, site tensor resident in shared memory, FP64 DMMA chain */
  BPX_KERNEL_SLICED = 3,  /* vertex-centric, site tensor streamed in leg slices, FP64 DMMA chain    */
  BPX_KERNEL_VERTEX = 4   /* SINGLE mode, small uniform link dims: one thread per vertex, factor in registers (HBM bound) */
};

int bpx_version(void);

/* ---- context ------------------------------------------------------------------------------------ */
int bpx_create(int device, bpx_ctx** out);
/* Single-process multi-GPU (the reference is one call tree on one task, beliefpropagation.jl:69-92): ONE context over
 * `ndev` devices.  Every entry point of this header works on it unchanged: the problem is described once, the library
 * partitions the vertices over the devices (balanced contiguous blocks of the vertex order; bpx_set_owner overrides), each
 * device stores the site tensors it owns and updates the out-edges of its vertices, cut-edge messages are stored
 * straight into the owning device's message set over NVLink and the residual is max-reduced through peer mailboxes inside
 * the sweep kernels (cudaDeviceEnablePeerAccess + plain pointers; no IPC handles, no process group, no NCCL).  All
 * launches are asynchronous and issued from the calling thread.  Results do not depend on the number of devices for the
 * on-chip kernel families (same per-vertex arithmetic); see DESIGN.md for the sliced kernel.
 * Not available on such a context: bpx_set_stream (one internal stream per device), bpx_device_* pointers, the per-rank
 * halo calls, bpx_sweep_sequence / bpx_iterate_diff (single device only), gates and two-site expectation values across
 * a cut edge (BPX_ERR_UNSUPPORTED, as on per-rank contexts). */
int bpx_create_multi(const int* devices, int ndev, bpx_ctx** out);
int bpx_num_devices(const bpx_ctx* ctx);
/* multi-device contexts: owner[v] = index into `devices` of the device that updates the out-edges of v (after
 * bpx_set_dims; site tensors already uploaded stay where they are needed).  bpx_get_owner reads the partition in use. */
int bpx_set_owner(bpx_ctx* ctx, const int32_t* owner);
int bpx_get_owner(const bpx_ctx* ctx, int32_t* owner_out /* nv */);
int bpx_destroy(bpx_ctx* ctx);
/* ctx may be NULL: returns the message of the last failed bpx_create on this thread. */
const char* bpx_last_error(const bpx_ctx* ctx);

/* ---- problem description ------------------------------------------------------------------------
 * Replaces the graph/dictionary walk of `incoming_messages` (messagecache.jl:124-131) and the factor
 * fetch `factors[src(edge)]` (beliefpropagation.jl:244, normnetwork.jl:49-54) by a one-off upload. */
int bpx_set_graph(bpx_ctx* ctx, int64_t nv, int64_t ne, const int64_t* src, const int64_t* dst,
                  const int32_t* slot);
/* phys_dim[nv] (ignored, may be NULL, in SINGLE mode); link_dim[ne] with link_dim[e] == link_dim[rev e].
 * Allocates device storage for site tensors and two message sets (synchronous ping-pong), buckets the
 * directed edges by (degree, link dims, physical dim).
 * Link dimensions 9..15 on degree-4 Float64 vertices are zero-padded to 16 INSIDE the library (the chi = 16 tensor-pipe
 * kernels then serve them; exact, see csrc/bpx_pad.cuh): every packed layout at this ABI keeps the caller's dimensions;
 * bpx_bucket_info and the raw device pointers (bpx_device_*) show the internal ones.  BPX_NO_PAD=1 disables it. */
int bpx_set_dims(bpx_ctx* ctx, int dtype, int mode, const int32_t* phys_dim, const int32_t* link_dim);

int64_t bpx_num_vertices(const bpx_ctx* ctx);
int64_t bpx_num_edges(const bpx_ctx* ctx);
int64_t bpx_rev(const bpx_ctx* ctx, int64_t e);
/* element (not byte) offsets / sizes in the packed host layouts; totals with v == nv / e == ne */
int64_t bpx_site_offset(const bpx_ctx* ctx, int64_t v);
int64_t bpx_message_offset(const bpx_ctx* ctx, int64_t e);
/* element offset of A_v inside the DEVICE site buffer (bpx_device_site_tensors); -1 if v is not resident on this
 * rank (partitioned contexts store only the tensors of the vertices they own) */
int64_t bpx_site_device_offset(const bpx_ctx* ctx, int64_t v);

/* kettensor(nn, v) for every v (normnetwork.jl:77), canonical layout, packed. */
int bpx_set_site_tensors(bpx_ctx* ctx, const void* packed);
int bpx_set_site_tensor(bpx_ctx* ctx, int64_t v, const void* data);
/* The iterate: `MessageCache(messages)` (beliefpropagation.jl:76, messagecache.jl:33-49). */
int bpx_set_messages(bpx_ctx* ctx, const void* packed);
/* (On connected ranks of a partitioned run -- bpx_halo_connect -- bpx_set_messages and bpx_fill_synthetic are COLLECTIVE: they
 * end with a cross-rank barrier, so that no rank starts sweeping into this rank's halo slots before every rank has
 * rewritten its message sets.) */
int bpx_get_messages(bpx_ctx* ctx, void* packed);
int bpx_get_message(bpx_ctx* ctx, int64_t e, void* data);

/* ---- the hot path -------------------------------------------------------------------------------
 * bpx_sweep: up to `max_sweeps` SYNCHRONOUS sweeps.  One sweep performs, for every directed edge,
 * `message_update!(::SimpleMessageUpdate, cache, factors, edge)` (beliefpropagation.jl:242-257) from the
 * previous sweep's messages, with the sum-normalisation (:248-253, `normalize` != 0) and the per-edge
 * term of `iterate_diff` (:261-267) fused into the kernel epilogue.  The loop stops after the first sweep
 * whose maximum residual is < tol (StopWhenConverged, AlgorithmsInterfaceExtensions.jl:84-119) or after
 * max_sweeps (StopAfterIteration); tol <= 0 disables the convergence test.  On a single device the test runs ON
 * THE DEVICE: sweeps are enqueued in batches (4, 8, 16, ...) ahead of the host and every launch of a sweep first
 * checks the previous sweep's residual key, turning into a no-op once it is below tol -- the iterate is exactly
 * the one after the first converged sweep, and the host synchronises once per batch instead of once per sweep.
 * Partitioned / multi-device contexts check the GLOBAL residual on the host after every sweep.
 * residual_out: residual of the last executed sweep; sweeps_done: number executed (either may be NULL). */
int bpx_sweep(bpx_ctx* ctx, int max_sweeps, double tol, int normalize, double* residual_out,
              int* sweeps_done);
/* Enqueue `n_sweeps` synchronous sweeps on the context's stream and return without waiting (no
 * convergence test).  Pair with bpx_synchronize / bpx_last_residual / bpx_residual_history. */
int bpx_sweep_async(bpx_ctx* ctx, int n_sweeps, int normalize);
/* One step with HOST buffers: upload the iterate (packed_in), one synchronous sweep, download the new iterate
 * (packed_out) and the fused residual; a single host synchronisation.  What the per-sweep Julia hook
 * (AI.step! specialisation + StopWhenConverged) calls when the messages live on the host.
 *  - Partitioned contexts move only the messages of the edges this rank owns (the out-edges of its vertices): those
 *    slices of packed_in are read, those slices of packed_out are written; messages on cut edges travel between the
 *    devices; residual_out is the GLOBAL residual of the sweep (the call waits for every rank's post).
 *  - If both buffers are page-locked (bpx_host_register, cudaHostAlloc, torch pin_memory) and the sweep is one
 *    launch of an on-chip kernel, the step is STREAMED: the kernel starts at once and every work item waits until
 *    the chunk of the upload that holds its messages has landed; new messages are stored straight into packed_out
 *    by the kernel (no download); on a single rank the whole step is one cached CUDA-graph launch.
 *    Results are identical to the staged path (pageable buffers, or sweeps of several launches). */
int bpx_sweep_host(bpx_ctx* ctx, const void* packed_in, void* packed_out, int normalize, double* residual_out);
/* Page-lock (and map) a caller-owned host buffer so that bpx_sweep_host can stream through it; a thin wrapper of
 * cudaHostRegister / cudaHostUnregister for hosts without their own CUDA binding (the Julia glue). */
int bpx_host_register(bpx_ctx* ctx, void* ptr, size_t bytes);
int bpx_host_unregister(bpx_ctx* ctx, void* ptr);
/* Reference schedule: in-place (Gauss-Seidel) updates along an explicit directed-edge list
 * (beliefpropagation.jl:200-210, 255).  Runs of consecutive, mutually independent updates are batched. */
int bpx_sweep_sequence(bpx_ctx* ctx, const int64_t* edge_seq, int64_t n_seq, int max_sweeps, double tol,
                       int normalize, double* residual_out, int* sweeps_done);
/* residuals of the sweeps of the last bpx_sweep* call (up to n). Returns the count via *n_out. */
int bpx_residual_history(bpx_ctx* ctx, double* out, int n, int* n_out);
/* residual of the most recent sweep (synchronises the stream) */
int bpx_last_residual(bpx_ctx* ctx, double* out);
/* `iterate_diff(cache, other)` against a host copy of another message set (beliefpropagation.jl:261). */
int bpx_iterate_diff(bpx_ctx* ctx, const void* other_packed, double* out);

/* ---- beliefs (messagecache.jl:139-201), same contraction kernels with all z messages absorbed ---- */
/* partitioned contexts fill in the vertices they own and report 0 for the others (combine across ranks on the host) */
int bpx_vertex_scalars(bpx_ctx* ctx, void* out /* nv elements */);
int bpx_edge_scalars(bpx_ctx* ctx, void* out /* ne/2 elements, edges with e < rev(e) in edge order */);
/* `bethe_free_energy(factors, messages)` (messagecache.jl:185-201) without leaving the device: vertex scalars on the
 * update kernels, edge scalars, then ONE reduction kernel forms sum(log.(numerators)) - sum(log.(denominators)); 7 doubles
 * come back.  out = (re, im); *promoted (may be NULL) = 1 when the reference's result is Complex (complex element type, or a
 * term with negative real part, :189-194 -- then im carries the phases, a multiple of pi for real networks), 0 when it is
 * a real number (im = 0).  A zero edge scalar gives (-inf, 0) like :196-198. */
int bpx_bethe_free_energy(bpx_ctx* ctx, double out[2], int* promoted);
/* the same reduction, raw, for process-per-GPU contexts (each rank reports the vertices it owns and the undirected edges
 * whose first orientation starts there; add [0..3] and OR [4..6] across ranks): parts = { sum log|num|, sum arg(num),
 * sum log|den|, sum arg(den), any real(num) < 0, any real(den) < 0, any den == 0 } */
int bpx_bethe_free_energy_parts(bpx_ctx* ctx, double parts[7]);
/* numerator of <O_v>: vertex contraction with the d x d operator `op[s_out, s_in]` applied to the ket
 * site leg, for every vertex (ops packed per vertex, d_v*d_v elements each).  Build-defined extension:
 * the reference has no `expect`. */
int bpx_vertex_expect_numerators(bpx_ctx* ctx, const void* ops_packed, void* out /* nv elements */);

/* ---- the consumer of the messages: BP simple-update gate application ------------------------------------
 * Replaces `apply_gate_bp!` (src/apply/apply_operators.jl:213-283, the default `BPApplyGate` strategy of
 * `apply_operator`, :190-211) for a BATCH of vertex-disjoint gates, e.g. one layer of a Trotter circuit: the
 * reference applies one gate per call on the host; here one CTA owns one gate and one launch covers the layer.
 * Site tensors (the ket layer) and the current message set are updated IN PLACE on the device
 * (the reference copies state and env first, `initialize_output`, :204-208 -- the Julia glue copies if it must).
 *
 * Two-site gates (:246-283): edges[g] = a directed edge v1 -> v2 of the graph; ops_packed holds for every gate the
 *   operator `op[o1, o2, i1, i2]` (column-major, d1*d2*d1*d2 elements; 1 = src, 2 = dst of the edge; outputs first
 *   like ITensorBase `operator`s, test/test_apply_operator.jl:19-27), packed in gate order.  Per gate: gauges
 *   from the incoming boundary messages (gram_eigh_full_with_pinv) -> QR -> gate -> truncated SVD -> sqrt(S) split ->
 *   inverse gauges; the messages of BOTH directions of the gate edge become diag(S) (:273-277).
 *   max_rank: `trunc` rank; 0 = keep the bond dimension.  The link leg keeps its dimension chi_e: the kept rank is
 *   k = min(max_rank, chi_e, rank bound of the bond matrix) and tensors / messages are zero-padded from k to chi_e
 *   (the reference would shrink or grow the leg; to grow a bond re-declare the dims with bpx_set_dims).
 *   normalize != 0: S <- S / |S| (:262-264).
 *   singular_values_out (may be NULL): link_dim[edges[g]] doubles per gate, packed in gate order (kept values, then 0).
 * One-site gates (:226-244): vertices[g]; ops_packed holds `op[o, i]` (d_v*d_v elements) per gate; normalize != 0
 *   divides the new tensor by the norm of its gauged version (all incoming messages).
 * BPX_ERR_INVALID if two gates of one call share a vertex.  Partitioned contexts: every rank passes the gates that lie
 * inside its own block (all vertices owned by the rank; the call first waits for the cut-edge messages of the last
 * sweep); a gate that touches a foreign vertex -- a gate across a cut edge -- is BPX_ERR_UNSUPPORTED. */
int bpx_apply_two_site_gates(bpx_ctx* ctx, int64_t n_gates, const int64_t* edges, const void* ops_packed,
                             int max_rank, int normalize, double* singular_values_out);
int bpx_apply_one_site_gates(bpx_ctx* ctx, int64_t n_gates, const int64_t* vertices, const void* ops_packed,
                             int normalize);
/* how the two-site gates of this context were applied so far: out[0] = by the Gram-path kernel (version 3,
 * csrc/bpx_apply3.cuh), out[1] = declined by it at run time (rank-deficient / indefinite boundary message, ill-conditioned
 * Gram matrix) and applied by the step-by-step kernel instead.  reset != 0 clears the counters. */
int bpx_apply_stats(bpx_ctx* ctx, int64_t out[2], int reset);
/* download one site tensor (canonical layout) -- `state[v]` after gates were applied */
int bpx_get_site_tensor(bpx_ctx* ctx, int64_t v, void* data);

/* Two-site expectation values in the BP environment, the quantity a simple-update evolution monitors (bond energies):
 * for every listed directed edge e = v1 -> v2, both norm-network factors contracted with every incoming message except
 * the two on the shared link, `op[o1, o2, i1, i2]` (column-major, d1*d2*d1*d2 elements per edge, packed in list order)
 * applied to the two ket site legs.  num_out[g] / den_out[g] = <O_e>; den is the same contraction with the identity.
 * Build-defined extension like bpx_vertex_expect_numerators: the two-vertex analogue of `vertex_scalar`
 * (messagecache.jl:139-143); the reference has no `expect`.  Read-only: edges may share vertices. */
int bpx_edge_expect(bpx_ctx* ctx, int64_t n_edges, const int64_t* edges, const void* ops_packed, void* num_out,
                    void* den_out);

/* ---- introspection --------------------------------------------------------------------------------- */
int bpx_num_buckets(const bpx_ctx* ctx);
/* info[0]=degree, [1]=chi (0 if non-uniform), [2]=phys dim, [3]=#vertices, [4]=#directed edges,
 * [5]=kernel family in use, [6]=bucket whose launch covers this bucket (buckets of one kernel family and
 * chi/d share a launch; only the leader is timed by bpx_bucket_time), [7]=reserved */
int bpx_bucket_info(const bpx_ctx* ctx, int bucket, int64_t info[8]);
/* force a kernel family for all buckets that support it (testing / profiling); BPX_KERNEL_AUTO resets */
int bpx_set_kernel_policy(bpx_ctx* ctx, int kernel);
/* per-bucket device timing: when enabled, every bucket launch of a sweep is bracketed by CUDA events on
 * the context's stream.  bpx_bucket_time returns the summed milliseconds and the number of launches
 * timed since profiling was (re-)enabled (it synchronises the stream). */
int bpx_set_profiling(bpx_ctx* ctx, int enable);
int bpx_bucket_time(bpx_ctx* ctx, int bucket, double* total_ms, int64_t* launches);
/* counters since the last reset: [0]=kernel launches issued by this library, [1]=message updates,
 * [2]=sweeps */
int bpx_counters(bpx_ctx* ctx, int64_t out[3], int reset);

/* ---- device-side access for callers that already hold device memory ------------------------------ */
/* use an external cudaStream_t (e.g. torch's current stream); NULL restores the internal stream */
int bpx_set_stream(bpx_ctx* ctx, void* cuda_stream);
void* bpx_device_messages(bpx_ctx* ctx);     /* current message set, packed, device pointer */
void* bpx_device_site_tensors(bpx_ctx* ctx); /* packed, device pointer */
int bpx_synchronize(bpx_ctx* ctx);

/* ---- multi-GPU: vertex partition, one context per rank ------------------------
 * owner[v] = rank that updates the out-edges of v.  A rank stores all site tensors it owns and all
 * messages; per sweep it updates only its own edges and pushes the messages on cut edges straight into
 * the peers' message sets over NVLink (peer pointers from bpx_halo_export / bpx_halo_connect).  Site tensors are
 * kept only for owned vertices (tensors uploaded earlier are preserved; later uploads skip foreign vertices). */
int bpx_set_partition(bpx_ctx* ctx, int rank, int nranks, const int32_t* owner);
/* three 64-byte cudaIpcMemHandle_t: message set 0, message set 1, residual mailbox.  Exchange them between
 * the ranks by any host-side means (torch.distributed / MPI) and connect every peer; sweeps then push
 * cut-edge messages into the peers and exchange the residual through the mailboxes (no NCCL involved). */
int bpx_halo_export(bpx_ctx* ctx, void* handles_3x64);
int bpx_halo_connect(bpx_ctx* ctx, int peer_rank, const void* handles_3x64);
int64_t bpx_num_cut_edges(const bpx_ctx* ctx);
/* enqueue a device-side barrier over all connected ranks on the context's stream (collective: every rank calls it) */
int bpx_peer_barrier(bpx_ctx* ctx);

/* Fill the RESIDENT site tensors and messages with the synthetic benchmark recipe on the device (site tensor v =
 * randn(seed, stream v) / sqrt(n_v); message e = (I + 0.1 |randn(seed, stream nv + e)|) sum-normalised): same
 * counter-based generator as bpx_fill_randn, evaluated with device math (may differ from the host in the last ulp).
 * For workloads whose inputs the host cannot stage (BASELINE config 5: 63 GiB of site tensors). */
int bpx_fill_synthetic(bpx_ctx* ctx, uint64_t seed);

/* ---- shared deterministic RNG (host): splitmix64 counter -> Box-Muller standard normals ----------
 * out[i] depends only on (seed, stream, i); complex: (N(0,1) + i N(0,1)) / sqrt(2).                 */
int bpx_fill_randn(uint64_t seed, uint64_t stream, int dtype, int64_t n, void* out);

#ifdef __cplusplus
}
#endif
#endif /* BPX_H */
