"""GPU tier: the CUDA-graph replay of a training step (semseg_b200/graphs.py) is the same computation as the eager path.

The kernels are deterministic, so three optimiser steps taken through the captured graphs must leave the model in
bit-identical state to three eager steps: losses, parameter gradients, updated weights, BatchNorm running statistics and
num_batches_tracked. Also: a changed input shape falls back to eager warm-up and captures a second graph; eval mode and
SEMSEG_B200_GRAPH=0 never capture."""
import copy

import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _steps(model, batches, n_steps, lr=0.01):
    opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=0.9, weight_decay=1e-4)
    losses = []
    for k in range(n_steps):
        x, y = batches[k % len(batches)]
        _, ml, al = model(x, y)
        loss = ml + 0.4 * al
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append((ml.item(), al.item()))
    return losses


@pytest.mark.parametrize("arch", ["psp", "psa"])
def test_graphed_steps_bit_identical_to_eager(arch, monkeypatch):
    from semseg_b200 import graphs
    build = util.build_pspnet if arch == "psp" else util.build_psanet
    base = build(50, 21).cuda().train()
    batches = [util.synth(2, 65, 65, 21, seed=s, device="cuda") for s in (1, 2, 3)]
    n_steps = graphs.WARMUP_CALLS + 4                    # 3 eager warm-up calls, capture, then replays
    eager = copy.deepcopy(base)
    monkeypatch.setenv("SEMSEG_B200_GRAPH", "0")
    le = _steps(eager, batches, n_steps)
    assert graphs.launches_per_step(eager) == 0
    graphed = copy.deepcopy(base)
    monkeypatch.setenv("SEMSEG_B200_GRAPH", "1")
    lg = _steps(graphed, batches, n_steps)
    assert graphs.launches_per_step(graphed) > 100       # the step really was captured and replayed
    assert le == lg, (le, lg)
    se, sg = eager.state_dict(), graphed.state_dict()
    for k in se:
        assert torch.equal(se[k], sg[k]), k              # weights, running statistics, num_batches_tracked
    for (k, pe), (_, pg) in zip(eager.named_parameters(), graphed.named_parameters()):
        assert torch.equal(pe.grad, pg.grad), k
    # eval on the trained weights still goes through the eager single-kernel path and agrees
    eager.eval(), graphed.eval()
    with torch.no_grad():
        assert torch.equal(eager(batches[0][0]), graphed(batches[0][0]))
    # another input shape: eager warm-up again, then a second captured step; training continues to work
    graphed.train()
    other = [util.synth(2, 73, 73, 21, seed=9, device="cuda")]
    l2 = _steps(graphed, other, graphs.WARMUP_CALLS + 2)
    assert all(torch.isfinite(torch.tensor(v)).all() for v in l2)
    assert len([s for s in graphed.__dict__["_sb_graph_steps"].values() if s.fwd is not None]) == 2
