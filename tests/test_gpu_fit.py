"""End-to-end use of the path (rasterizer-level): the reference's iteration shape — two frames
x (front + back) in one batched chain, L1 loss, Adam — must actually fit target frames, i.e. the gradients of the
whole chain point downhill."""
import pytest

pytestmark = pytest.mark.gpu


def test_fit_window_converges(cuda_device):
    from examples.fit_window import fit
    losses = fit(cuda_device, iters=150, P=6000, W=192, H=128, Fr=192)
    assert all(l == l for l in losses)                      # no NaN
    first, last = sum(losses[:5]) / 5, sum(losses[-5:]) / 5
    assert last < 0.6 * first, (first, last)
    d = fit.last_densify                                    # the densification statistic rode along
    assert 0 < d["seen"] <= 6000 and 0 <= d["candidates"] <= d["seen"] and d["checksum"] > 0
