"""OPT-IN version 2 of the two-site gate kernel (csrc/bpx_apply2.cuh, BPX_APPLY_V2=1) against version 1 on the GPU.
Version 2 has only run on the host so far (tests/test_apply_device_code.py); this file sorts after every other test."""
import numpy as np
import pytest

import itnn_b200 as B
from helpers import randn
from itnn_b200 import graphs, problems
from test_zz_gpu_apply import bond_invariant, device_tensors, matching, oracle_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("lattice,chi,max_rank,normalize", [((4, 4), 4, 0, False), ((3, 5), 3, 2, True), ((4, 4), 8, 8, True)])
def test_v2_matches_v1(monkeypatch, dtype, lattice, chi, max_rank, normalize):
    rng = np.random.default_rng(chi + max_rank)
    p = problems.synthetic_peps(graphs.named_grid(lattice), chi, 2, dtype, init="positive")
    edges = matching(p.ga, rng)
    ops = [randn(rng, dtype, (2, 2, 2, 2)) for _ in edges]
    results = []
    for v2 in ("0", "1"):
        monkeypatch.setenv("BPX_APPLY_V2", v2)
        with B.BPXContext(0) as ctx:
            problems.upload(ctx, p)
            ctx.sweep(5, 0.0, True)
            svs = ctx.apply_two_site_gates(edges, ops, max_rank=max_rank, normalize=normalize)
            results.append((svs, device_tensors(ctx, p), ctx.get_messages()))
    (sv1, t1, m1), (sv2, t2, m2) = results
    s1, s2 = oracle_state(p, t1), oracle_state(p, t2)
    for e, a, b in zip(edges, sv1, sv2):
        assert np.allclose(a, b, rtol=1e-10, atol=1e-14)
        v, w = p.ga.src[e], p.ga.dst[e]
        x, y = bond_invariant(s1, v, w), bond_invariant(s2, v, w)
        assert np.abs(x - y).max() <= 1e-9 * np.abs(x).max()
    assert all(np.allclose(a, b, rtol=1e-10, atol=1e-14) for a, b in zip(m1, m2))
