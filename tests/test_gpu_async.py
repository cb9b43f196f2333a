"""Asynchronous state accessors (gvib200_set_state_async / gvib200_get_*_async / gvib200_sync, include/gvib200.h):
two handles of the same problem pipelined the way bench.py's end-to-end leg does it reproduce the synchronous calls bit
for bit and agree with the oracle; a precision that is not positive definite is reported by the next synchronising call
on the handle and leaves it without a state."""
import numpy as np
import pytest
import torch

import oracle_bridge as ob
from gaussianvi_b200 import capi, problems

pytestmark = pytest.mark.gpu


def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64).pin_memory()
    t.numpy()[...] = a
    return t.numpy()


def c_layout(spec):
    return (pinned(np.ascontiguousarray(spec.mu0, dtype=np.float64)),
            pinned(np.ascontiguousarray(np.transpose(spec.prec0_D, (0, 2, 1)))),
            pinned(np.ascontiguousarray(np.transpose(spec.prec0_O, (0, 2, 1)))))


def test_pipelined_handles_reproduce_synchronous_calls(gpu_ctx):
    spec = problems.make_cfg3(N=3000)
    S, d = spec.mu0.size // 4, 4
    mu_h, pD_h, pO_h = c_layout(spec)
    opts = capi.Problem.default_opts()
    opts.reuse_accepted_sweep = 1

    ref_p = problems.build_device_problem(gpu_ctx, spec)
    ref_p.set_state_raw(mu_h, pD_h, pO_h)
    st_ref = ref_p.iterate(opts)
    mu_ref = np.zeros(S * d)
    cD_ref, cO_ref = np.zeros((S, d, d)), np.zeros((S - 1, d, d))
    ref_p.get_mean_into(mu_ref)
    ref_p.get_cov_blocks_into(cD_ref, cO_ref)

    handles = []
    for _ in range(2):
        h = problems.build_device_problem(gpu_ctx, spec)
        handles.append((h, pinned(np.zeros(S * d)), pinned(np.zeros((S, d, d))), pinned(np.zeros((S - 1, d, d)))))
    steps = 6
    costs = []
    handles[0][0].set_state_raw_async(mu_h, pD_h, pO_h)
    for i in range(steps):
        h, om, oD, oO = handles[i % 2]
        if i + 1 < steps:
            handles[(i + 1) % 2][0].set_state_raw_async(mu_h, pD_h, pO_h)
        st = h.iterate(opts)
        costs.append((st.cost, st.new_cost, st.accepted, st.n_backtrack))
        h.get_mean_into_async(om)
        h.get_cov_blocks_into_async(oD, oO)
    for h, om, oD, oO in handles:
        h.sync()
        assert np.array_equal(om, mu_ref)
        assert np.array_equal(oD, cD_ref)
        assert np.array_equal(oO, cO_ref)
    assert all(c == (st_ref.cost, st_ref.new_cost, st_ref.accepted, st_ref.n_backtrack) for c in costs)

    # and the step itself is the oracle's
    ref = ob.build_oracle(spec, niters=1)
    recs = ref.optimize()
    assert abs(st_ref.cost - recs[0].cost) < 1e-9 * max(1.0, abs(recs[0].cost))
    assert np.abs(mu_ref - ref.mean()).max() < 1e-7 * np.abs(ref.mean()).max()
    for h, *_ in handles:
        h.close()
    ref_p.close()


def test_async_set_state_reports_indefinite_precision_on_next_call(gpu_ctx):
    spec = problems.make_cfg3(N=40)
    mu_h, pD_h, pO_h = c_layout(spec)
    p = problems.build_device_problem(gpu_ctx, spec)
    bad = pinned(pD_h.copy())
    bad[7] = -bad[7]                       # one diagonal block negative definite
    p.set_state_raw_async(mu_h, bad, pO_h)  # returns without waiting
    with pytest.raises(capi.GviError) as e:
        p.iterate(capi.Problem.default_opts())
    assert e.value.code == capi.E_NOTSPD
    # the handle is usable again after a valid state
    p.set_state_raw_async(mu_h, pD_h, pO_h)
    p.sync()
    st = p.iterate(capi.Problem.default_opts())
    assert st.status == 0 and st.accepted
    p.close()
