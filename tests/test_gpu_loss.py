"""GPU parity of the training loss (VGG16 features, Gram style loss, content loss, total variation), forward and
backward, against the oracle (oracle/rst_oracle.py, autograd for the gradient).  Bar (north_star): Gram-loss scalar
within 1e-3 relative."""
import numpy as np
import pytest
import torch

from oracle import rst_oracle as O
from realtime_style_transfer_b200 import _native
from realtime_style_transfer_b200.models import styleLoss

pytestmark = pytest.mark.gpu


def _inputs(b, h, w, seed=0):
    rng = np.random.default_rng(seed)
    pred = rng.uniform(0.05, 0.95, (b, h, w, 3)).astype(np.float32)
    content = rng.uniform(0, 1, (b, h, w, 3)).astype(np.float32)
    style = rng.uniform(0, 1, (b, 1, h, w, 3)).astype(np.float32)
    return pred, content, style


@pytest.mark.parametrize("b,h,w", [(2, 64, 96), (1, 128, 160)])
def test_loss_forward_matches_oracle(cuda_device, b, h, w):
    vgg = O.init_vgg16_weights(seed=3)
    pred, content, style = _inputs(b, h, w)
    ref = O.style_loss_vgg(vgg, pred, content, style)
    model = styleLoss.StyleLossModelVGG((h, w, 3))
    model.set_weights(vgg)
    compute_loss, m2 = styleLoss.make_style_loss_function(model, (h, w, 3), 1, with_depth_loss=False)
    assert m2 is model
    got = compute_loss(pred, {"content": content, "style": style})
    for key in ("loss", "feature_loss", "style_loss", "total_variation_loss"):
        r = ref[key].numpy()
        assert got[key].shape == (b,)
        rel = np.abs(got[key] - r).max() / max(np.abs(r).max(), 1e-12)
        print(key, got[key], r, rel)
        assert rel < 1e-3, key
    with pytest.raises(AssertionError):
        styleLoss.make_style_loss_function(model, (h, w, 3), 2, with_depth_loss=False)


def test_loss_backward_matches_autograd(cuda_device):
    b, h, w = 2, 64, 96
    vgg = O.init_vgg16_weights(seed=3)
    pred, content, style = _inputs(b, h, w, seed=1)
    p = torch.tensor(pred, dtype=torch.float64, requires_grad=True)
    out = O.style_loss_vgg(vgg, p, content, style, dtype=torch.float64)
    out["loss"].sum().backward()                  # Keras differentiates the (B,) loss vector = its batch sum
    ref_grad = p.grad.numpy()
    loss = _native.NativeLoss(h, w, b)
    loss.set_weights(vgg)
    d_pred = torch.tensor(pred).to(cuda_device)
    d_c = torch.tensor(content).to(cuda_device)
    d_s = torch.tensor(style[:, 0]).to(cuda_device)
    d_l = torch.empty((b, 4), device=cuda_device)
    d_g = torch.zeros((b, h, w, 3), device=cuda_device)
    st = torch.cuda.current_stream().cuda_stream
    loss.forward(d_pred.data_ptr(), d_c.data_ptr(), d_s.data_ptr(), d_l.data_ptr(), b, st)
    loss.backward(d_pred.data_ptr(), d_g.data_ptr(), b, st)
    got = d_g.cpu().numpy()
    num = np.sqrt(((got - ref_grad) ** 2).sum())
    den = np.sqrt((ref_grad ** 2).sum())
    print("grad rel l2", num / den, "max ref", np.abs(ref_grad).max())
    assert num / den < 1e-3
    assert np.abs(d_l.cpu().numpy()[:, 0] - out["loss"].detach().numpy()).max() / np.abs(out["loss"].detach().numpy()).max() < 1e-3
    loss.close()
