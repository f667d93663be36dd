"""Row N3 - the training step: `loss.backward()` of the reference's loop (train/ucf_train.py:60-106) through the drop-in
`MMFMIL` and `CLAS2`.  Gradients are checked against goldens produced by the reference's own autograd
(tests/golden/train_small.npz: every gradient tensor of a 128-d model; train_full.npz: norms + sampled entries of all 78
gradients of the 768-d model on 8 config-4 clips), eval mode (dropout off: its mask comes from PyTorch's generator).
Tolerance 1e-3 of each tensor's largest gradient entry (the GEMMs use bf16 operands with the 3-term split).  Needs a B200."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, params_from_npz

pytestmark = pytest.mark.gpu


def reference_loss(model, CLAS2, img, ev, lengths, labels, nu):
    """train/ucf_train.py:60-102 as the reference's caller writes it (torch ops on the module's outputs)."""
    outputs = model(img, ev, None, None, lengths)
    logits = outputs['logits']
    image_mu, event_mu = outputs['image_mu'], outputs['event_mu']
    image_logvar, event_logvar = outputs['image_logvar'], outputs['event_logvar']
    loss_classification = CLAS2(logits, labels, lengths, img.device)
    image_mu_norm = F.normalize(image_mu, p=2, dim=-1)
    event_mu_norm = F.normalize(event_mu, p=2, dim=-1)
    cos_sim = F.cosine_similarity(image_mu_norm, event_mu_norm, dim=-1)
    loss_reg = (1 - cos_sim).mean() + torch.abs(torch.norm(image_mu, p=2, dim=-1) - torch.norm(event_mu, p=2, dim=-1)).mean()
    eli = image_logvar + math.log(nu / (nu + 1))
    ele = event_logvar + math.log(nu / (nu + 1))
    loss_kl = (-0.5 * torch.mean(1 + eli - image_mu.pow(2) - eli.exp())
               - 0.5 * torch.mean(1 + ele - event_mu.pow(2) - ele.exp()))
    return loss_classification + loss_reg + loss_kl, loss_classification


@pytest.fixture(scope="module")
def pkg():
    assert torch.cuda.is_available()
    import iefvad_b200
    from iefvad_b200 import synth
    from iefvad_b200.loss import CLAS2
    return iefvad_b200, synth, CLAS2


def test_all_gradients_of_the_small_model_match_the_reference_autograd(pkg):
    iefvad_b200, synth, CLAS2 = pkg
    z = load_golden("train_small.npz")
    args = synth.default_args(noise_model="StudentT", nu=8, num_refinement_steps=3, visual_head=4)
    m = iefvad_b200.MMFMIL(14, 128, 256, 128, 8, 2, 8, 10, 10, "cuda", args)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params_from_npz(z).items()})
    m = m.cuda().eval()                                  # eval: dropout off, but grad enabled -> autograd path
    img, ev = torch.from_numpy(z["img"]).cuda(), torch.from_numpy(z["ev"]).cuda()
    lengths, labels = torch.from_numpy(z["lengths"]).cuda(), torch.from_numpy(z["labels"]).cuda()
    loss, lcls = reference_loss(m, CLAS2, img, ev, lengths, labels, 8)
    assert loss.grad_fn is not None
    assert abs(float(loss) - float(z["loss"])) < 2e-4 * abs(float(z["loss"]))
    assert abs(float(lcls) - float(z["loss_cls"])) < 2e-4
    loss.backward()
    worst = 0.0
    for k, p in m.named_parameters():
        ref = z["grad:" + k]
        assert p.grad is not None and p.grad.shape == ref.shape, k
        err = float(np.max(np.abs(p.grad.cpu().numpy() - ref)) / max(np.max(np.abs(ref)), 1e-12))
        worst = max(worst, err)
        assert err < 1e-3, (k, err)
    print("worst gradient error (max-norm relative):", worst)


def test_full_size_gradients_match_the_reference_autograd(pkg):
    iefvad_b200, synth, CLAS2 = pkg
    z = load_golden("train_full.npz")
    m = synth.build_model(iefvad_b200.MMFMIL, seed=0).cuda().eval()
    assert synth.state_digest(m.state_dict()) == str(z["digest"])
    img, ev, lengths, labels = synth.make_c4_batch(8)
    loss, _ = reference_loss(m, CLAS2, img.cuda(), ev.cuda(), lengths.cuda(), labels.cuda(), 8)
    assert abs(float(loss) - float(z["loss"])) < 2e-4 * abs(float(z["loss"]))
    loss.backward()
    n = 0
    for k, p in m.named_parameters():
        g = p.grad.cpu().numpy().reshape(-1)
        norm_ref = float(z["norm:" + k])
        assert abs(np.linalg.norm(g.astype(np.float64)) - norm_ref) < 2e-3 * norm_ref + 1e-12, k
        scale = max(norm_ref / math.sqrt(g.size), float(np.max(np.abs(z["val:" + k])))) + 1e-20
        err = np.abs(g[z["idx:" + k]] - z["val:" + k]) / scale
        # 90 % of the sampled entries within 1e-3; a lone outlier is a ReLU whose pre-activation sits within rounding of
        # zero (relu(W1 x + b1) of the refinement MLPs): its gate may flip between the fp32 reference and the split-bf16
        # GEMM, moving the gradients that pass through that one unit - bounded, not a precision defect
        assert np.quantile(err, 0.9) < 1e-3 and err.max() < 5e-2, (k, float(np.quantile(err, 0.9)), float(err.max()))
        n += 1
    assert n == 78


def test_one_optimizer_step_like_ucf_train(pkg):
    """The reference's loop body (train/ucf_train.py:28,60-106): AdamW step through model.train() with attention dropout."""
    iefvad_b200, synth, CLAS2 = pkg
    m = synth.build_model(iefvad_b200.MMFMIL, seed=0).cuda()
    m.train()
    opt = torch.optim.AdamW(m.parameters(), lr=2e-5)
    img, ev, lengths, labels = synth.make_c4_batch(4)
    img, ev, lengths, labels = img.cuda(), ev.cuda(), lengths.cuda(), labels.cuda()
    before = {k: v.detach().clone() for k, v in m.named_parameters()}
    losses = []
    for _ in range(2):
        loss, _ = reference_loss(m, CLAS2, img, ev, lengths, labels, m.temporal.nu)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses))
    assert all(not torch.equal(before[k], v.detach()) for k, v in m.named_parameters())
    m.eval()
    with torch.no_grad():                                 # the updated weights reach the inference kernels (version bump)
        out = m(img, ev, None, None, lengths)
    assert torch.isfinite(out["logits"]).all()


def test_dropout_is_a_pure_function_of_the_seed_and_backward_regenerates_it(pkg):
    iefvad_b200, synth, CLAS2 = pkg
    args = synth.default_args(num_refinement_steps=2, visual_head=4)
    torch.manual_seed(3)
    m = iefvad_b200.MMFMIL(14, 128, 256, 128, 8, 2, 8, 10, 10, "cuda", args).cuda().train()
    g = torch.Generator("cpu").manual_seed(1)
    img, ev = torch.randn(2, 48, 128, generator=g).cuda(), torch.randn(2, 48, 128, generator=g).cuda()

    def run(seed):
        torch.manual_seed(seed)
        m.zero_grad()
        out = m(img, ev, None, None, None)
        loss = out["logits"].square().mean() + out["image_mu"].mean()
        loss.backward()
        return float(loss), m.temporal.image_attn_layers[0].in_proj_weight.grad.clone()

    a, ga = run(11)
    b, gb = run(11)
    c, gc = run(12)
    assert a == b and torch.equal(ga, gb)                 # same seed: same mask in forward and backward
    assert a != c and not torch.equal(ga, gc)
    m.eval()
    e1 = m(img, ev, None, None, None)["logits"]
    m.train()
    torch.manual_seed(11)
    t1 = m(img, ev, None, None, None)["logits"]
    assert not torch.allclose(e1, t1)                     # dropout really acts in train() mode
    # directional derivative against a central difference under the SAME mask
    p = m.temporal.image_attn_layers[0].in_proj_weight
    d = torch.randn_like(p)
    _, g0 = run(21)
    eps = 1e-2
    with torch.no_grad():
        p.add_(eps * d)
    lp, _ = run(21)
    with torch.no_grad():
        p.add_(-2 * eps * d)
    lm, _ = run(21)
    with torch.no_grad():
        p.add_(eps * d)
    fd = (lp - lm) / (2 * eps)
    an = float((g0 * d).sum())
    assert abs(fd - an) < 0.05 * max(abs(an), 1e-3), (fd, an)


def test_standalone_backward_operators_against_torch_autograd(pkg):
    from iefvad_b200 import ops
    g = torch.Generator("cpu").manual_seed(2)
    B, T, H, dh = 2, 70, 4, 32
    D = H * dh
    qkv = torch.randn(B * T, 3 * D, generator=g).cuda().requires_grad_(True)
    dout = torch.randn(B * T, D, generator=g).cuda()
    q, k, v = (t.view(B, T, H, dh).transpose(1, 2) for t in qkv.split(D, dim=-1))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, D)
    ref.backward(dout)
    out, lse = ops.attention_train_fwd(qkv.detach(), B, T, H)
    assert torch.allclose(out, ref.detach(), atol=2e-5)
    dqkv = ops.attention_train_bwd(qkv.detach(), out, dout, lse, B, T, H)
    assert float((dqkv - qkv.grad).abs().max() / qkv.grad.abs().max()) < 1e-5
    # dropout: the expected value of the dropped weights is the undropped attention
    acc = sum(ops.attention_train_fwd(qkv.detach(), B, T, H, 0.1, s)[0] for s in range(64)) / 64
    assert float((acc - out).abs().mean() / out.abs().mean()) < 0.12
    # LayerNorm backward
    x = torch.randn(300, 128, generator=g).cuda().requires_grad_(True)
    w = (1 + 0.2 * torch.randn(128, generator=g)).cuda().requires_grad_(True)
    b = torch.randn(128, generator=g).cuda().requires_grad_(True)
    dy = torch.randn(300, 128, generator=g).cuda()
    F.layer_norm(x, (128,), w, b).backward(dy)
    dx, dw, db = ops.layernorm_bwd(x.detach(), w.detach(), dy)
    for got, want in ((dx, x.grad), (dw, w.grad), (db, b.grad)):
        assert float((got - want).abs().max() / want.abs().max()) < 1e-5
    # fusion backward (model/imf_vad.py:130-144, StudentT nu = 8) incl. gradients through the returned w_i / w_e
    ins = [torch.randn(500, 128, generator=g).cuda().requires_grad_(True) for _ in range(4)]
    mu_i, mu_e, lv_i, lv_e = ins
    wi_ = 1.125 * torch.exp(-lv_i)
    we_ = 1.125 * torch.exp(-lv_e)
    den = wi_ + we_ + 1e-8
    nwi, nwe = wi_ / den, we_ / den
    fused = nwi * mu_i + nwe * mu_e
    gs = [torch.randn(500, 128, generator=g).cuda() for _ in range(3)]
    (fused * gs[0] + nwi * gs[1] + nwe * gs[2]).sum().backward()
    got = ops.fuse_bwd(*(t.detach() for t in ins), gs[0], gs[1], gs[2])
    for a, t in zip(got, ins):
        assert float((a - t.grad).abs().max() / t.grad.abs().max()) < 2e-5
    # column sums / transpose
    a = torch.randn(1000, 96, generator=g).cuda()
    rw = torch.randn(1000, generator=g).cuda()
    assert torch.allclose(ops.colsum(a), a.sum(0), atol=1e-3)
    assert torch.allclose(ops.colsum(a, rw), (a * rw[:, None]).sum(0), atol=1e-3)
    t = ops.transpose(a, pad_to=64)
    assert t.shape == (96, 1024) and torch.equal(t[:, :1000], a.t()) and float(t[:, 1000:].abs().max()) == 0.0
