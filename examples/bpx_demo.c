/* Plain-C client of libbpx.so (include/bpx.h): what a non-Python, non-Julia host does with the drop-in boundary.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/bpx_demo.c -o bpx_demo -Litensornetworksnext.jl_b200/csrc -lbpx \
 *       -Wl,-rpath,$PWD/itensornetworksnext.jl_b200/csrc -lm
 *
 * Builds a 3x3 open square-lattice PEPS norm network (chi = 2, d = 2, Float64) from the library's deterministic RNG,
 * runs synchronous BP sweeps to 1e-10 (bpx_sweep: beliefpropagation.jl:69-92 + 200-210 + 242-267 of the reference),
 * applies one two-site gate by BP simple update (bpx_apply_two_site_gates: apply_operators.jl:246-283) and sweeps again.
 * Exit code 0 when every call succeeded, 2 when no sm_100 device is usable (there is no CPU fallback), 1 on any other failure. */
#include <stdio.h>
#include <stdlib.h>

#include "bpx.h"

#define NX 3
#define NY 3
#define CHI 2
#define D 2

static int fail(bpx_ctx* ctx, const char* what, int rc) {
  fprintf(stderr, "%s failed (%d): %s\n", what, rc, bpx_last_error(ctx));
  if (ctx) bpx_destroy(ctx);
  return rc == BPX_ERR_CUDA ? 2 : 1;
}

int main(void) {
  bpx_ctx* ctx = NULL;
  int rc = bpx_create(0, &ctx);
  if (rc) return fail(NULL, "bpx_create", rc);

  /* graph: vertices row-major, directed edges grouped by source in neighbour order (slot = position in the row) */
  int64_t src[4 * NX * NY], dst[4 * NX * NY];
  int32_t slot[4 * NX * NY], phys[NX * NY], link[4 * NX * NY];
  int64_t ne = 0;
  for (int y = 0; y < NY; ++y)
    for (int x = 0; x < NX; ++x) {
      const int v = x + NX * y;
      const int nb[4][2] = {{x - 1, y}, {x + 1, y}, {x, y - 1}, {x, y + 1}};
      int k = 0;
      phys[v] = D;
      for (int i = 0; i < 4; ++i) {
        if (nb[i][0] < 0 || nb[i][0] >= NX || nb[i][1] < 0 || nb[i][1] >= NY) continue;
        src[ne] = v;
        dst[ne] = nb[i][0] + NX * nb[i][1];
        slot[ne] = k++;
        link[ne] = CHI;
        ++ne;
      }
    }
  if ((rc = bpx_set_graph(ctx, NX * NY, ne, src, dst, slot))) return fail(ctx, "bpx_set_graph", rc);
  if ((rc = bpx_set_dims(ctx, BPX_F64, BPX_MODE_NORM, phys, link))) return fail(ctx, "bpx_set_dims", rc);

  /* site tensors A_v[s, l_0..l_{z-1}] and messages M_e[bra, ket], packed; offsets come from the library */
  const int64_t n_site = bpx_site_offset(ctx, NX * NY), n_msg = bpx_message_offset(ctx, ne);
  double* sites = (double*)malloc(sizeof(double) * (size_t)n_site);
  double* msgs = (double*)malloc(sizeof(double) * (size_t)n_msg);
  if (!sites || !msgs) return 1;
  for (int64_t v = 0; v < NX * NY; ++v)
    bpx_fill_randn(123, (uint64_t)v, BPX_F64, bpx_site_offset(ctx, v + 1) - bpx_site_offset(ctx, v), sites + bpx_site_offset(ctx, v));
  for (int64_t i = 0; i < n_msg; ++i) msgs[i] = 1.0; /* all-ones messages (test/test_apply_operator.jl:72) */
  if ((rc = bpx_set_site_tensors(ctx, sites))) return fail(ctx, "bpx_set_site_tensors", rc);
  if ((rc = bpx_set_messages(ctx, msgs))) return fail(ctx, "bpx_set_messages", rc);

  double residual = 0.0;
  int done = 0;
  if ((rc = bpx_sweep(ctx, 200, 1e-10, 1, &residual, &done))) return fail(ctx, "bpx_sweep", rc);
  printf("BP: %d synchronous sweeps x %lld updates, residual %.3e\n", done, (long long)ne, residual);

  /* a two-site gate on directed edge 0 (vertex 0 -> its first neighbour): op[o1, o2, i1, i2], here a random 4x4 */
  double op[D * D * D * D], sv[CHI];
  const int64_t edge = 0;
  bpx_fill_randn(7, 0, BPX_F64, D * D * D * D, op);
  if ((rc = bpx_apply_two_site_gates(ctx, 1, &edge, op, CHI, 1, sv))) return fail(ctx, "bpx_apply_two_site_gates", rc);
  printf("gate on edge %lld -> %lld: kept singular values %.6f %.6f\n", (long long)src[0], (long long)dst[0], sv[0], sv[1]);
  if ((rc = bpx_sweep(ctx, 200, 1e-10, 1, &residual, &done))) return fail(ctx, "bpx_sweep", rc);
  printf("BP after the gate: %d sweeps, residual %.3e\n", done, residual);

  double scalars[NX * NY];
  if ((rc = bpx_vertex_scalars(ctx, scalars))) return fail(ctx, "bpx_vertex_scalars", rc);
  printf("vertex scalar of vertex 0: %.6e\n", scalars[0]);
  free(sites);
  free(msgs);
  bpx_destroy(ctx);
  return 0;
}
