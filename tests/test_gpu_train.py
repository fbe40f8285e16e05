"""GPU tests of the training side (SURVEY.md §8 b / e-2 / f-3): the autograd boundary of the drop-in modules, the fused
AdamW kernel, the gradient-bucket kernels, and the reference's own trainers (unmodified, from baseline/_ref) driving the
drop-in model."""
import os
import sys

import pytest
import torch

import swinwnet_b200 as S
from oracle import swinwnet_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
D2 = [2, 2, 2, 2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_adamw_matches_torch_adamw():
    torch.manual_seed(0)
    shapes = [(48, 2, 2, 2), (48,), (144, 48), (81, 3), (1,), (7, 13), (4097,), (3 * 4096 + 5,)]
    ref = [torch.nn.Parameter(torch.randn(*s, device=DEV)) for s in shapes]
    mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    ref[3].requires_grad_(False)
    mine[3].requires_grad_(False)                                     # frozen parameter
    kw = dict(lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    o_ref, o_mine = torch.optim.AdamW(ref, **kw), S.train.FusedAdamW(mine, **kw)
    for step in range(5):
        for i, (a, b) in enumerate(zip(ref, mine)):
            if not a.requires_grad or (i == 5 and step % 2 == 0):     # parameter 5: no gradient on even steps
                a.grad = b.grad = None
                continue
            g = torch.randn_like(a) * (10.0 ** (step - 2))
            a.grad, b.grad = g.clone(), g.clone()
        o_ref.step()
        o_mine.step()
    torch.cuda.synchronize()
    for a, b in zip(ref, mine):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (a - b).abs().max()
    st = o_mine.state[mine[2]]
    assert torch.allclose(st["exp_avg"], o_ref.state[ref[2]]["exp_avg"], rtol=1e-4, atol=1e-5)


def test_grad_bucket_pack_unpack():
    ps = [torch.nn.Parameter(torch.randn(n, device=DEV)) for n in (5, 4096, 4100, 1, 12345)]
    for i, p in enumerate(ps):
        p.grad = None if i == 3 else torch.randn_like(p)
    want = [None if p.grad is None else p.grad.clone() for p in ps]
    red = S.train.FlatGradReducer(ps, average=True)
    n = red.reduce()                                     # world size 1: pack, (no all-reduce), unpack * 1
    torch.cuda.synchronize()
    assert n == 5 + 4096 + 4100 + 12345
    for p, w in zip(ps, want):
        assert (p.grad is None) if w is None else torch.equal(p.grad, w)
    off = red.tab.offsets
    assert torch.equal(red.flat[off[1]:off[1] + 4096], want[1]) and torch.equal(red.flat[off[4]:off[4] + 12345], want[4])


def _oracle_loss(sd, x, xh):
    seg, skips = O.segment_1(sd, x)
    up, sk2 = O.upscale(sd, xh, skips)
    seg2, _ = O.segment_2(sd, up, sk2)
    return seg.square().mean() + up.square().mean() + seg2.square().mean(), (seg, up, seg2)


def test_autograd_gradients_match_fp32_oracle(manifest):
    """segment_1 -> upscale (half-resolution input, Lq != Lkv cross attention) -> segment_2 with autograd on: the forward
    values are the kernels', the gradients come from KernelOp's recompute; both against the fp32 oracle differentiated by
    torch on the same device"""
    sd = O.make_state_dict(manifest["wnet_em"], seed=1)
    m = S.SwinWNet(error_matrix=True, depths=D2)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).train()
    x = (O.synthetic_diffractions(2, seed=3, H=40, W=60) / 100.0).to(DEV)
    xh = torch.nn.functional.interpolate(x, scale_factor=0.5, mode="bilinear", align_corners=False)
    seg, skips = m.segment_1(x)
    up, sk2 = m.upscale(xh, skips)
    seg2, _ = m.segment_2(up, sk2)
    loss = seg.square().mean() + up.square().mean() + seg2.square().mean()
    loss.backward()
    sdd = {k: v.to(DEV).clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    ref_loss, (rseg, rup, rseg2) = _oracle_loss(sdd, x, xh)
    ref_loss.backward()
    for a, b, n in ((seg, rseg, "seg"), (up, rup, "up"), (seg2, rseg2, "seg2")):
        assert (a - b).abs().max().item() <= 2e-2 * b.abs().max().item(), n
    assert abs(loss.item() - ref_loss.item()) <= 1e-2 * abs(ref_loss.item())
    worst = 0.0
    checked = 0
    for name, p in m.named_parameters():
        g, r = p.grad, sdd[name].grad
        if r is None or r.abs().max() == 0:
            continue
        assert g is not None, name
        e = (g - r).abs().max().item() / r.abs().max().item()
        worst = max(worst, e)
        checked += 1
        assert e <= 0.1, (name, e)
    print(f"gradients checked: {checked} tensors, worst max-norm relative error {worst:.3e}")
    assert checked > 500


def _ref_trainers():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from stage_reference import import_reference, ref_dir
    if ref_dir() is None:
        pytest.skip("the reference is not staged (baseline/_ref)")
    _, R, _ = import_reference()
    from Segmentator_pretrain import SegmentatorTrainer
    from FullModel_supervised_trainer import FullModelTrainer
    return R, SegmentatorTrainer, FullModelTrainer


def test_reference_trainers_drive_the_dropin(manifest):
    """the reference's SegmentatorTrainer and FullModelTrainer, UNMODIFIED, on the drop-in model with FusedAdamW: freeze
    logic, autocast + GradScaler, even / odd steps.
    (1) from identical weights, the even- and odd-step losses and their gradients equal those of the reference model
        driven by the same trainer code (fp16 autocast) within the 16-bit tolerances;
    (2) the trainers' own loops run on the drop-in and reduce the loss."""
    R, SegT, FullT = _ref_trainers()
    sd = O.make_state_dict(manifest["wnet_em"], seed=1)
    H, W = 40, 60
    x = O.synthetic_diffractions(8, seed=5, H=H, W=W, two_channel=False) / 100.0
    masks = (x[:, 0] > 3.0).long()
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, masks), batch_size=2, shuffle=False)

    def one_step(model, which, amp):
        model.load_state_dict(sd, strict=True)
        model = model.to(DEV).train()
        for p in model.parameters():
            p.requires_grad = True
            p.grad = None
        f = FullT(model, loader, loader, DEV, num_epochs=2, warmup_epochs=0, lr=1e-3, verbose=False)
        imgs, mk = f.ensure_2ch(x[:2].to(DEV)), masks[:2].unsqueeze(1).to(DEV)
        with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            loss, a, b = (f._even_step if which == "even" else f._odd_step)(imgs, mk)
        loss.float().backward()
        return loss.item(), a, b, {n: (None if p.grad is None else p.grad.float().clone()) for n, p in model.named_parameters()}

    for which in ("even", "odd"):
        # reference side in fp32 (its fp16-autocast gradients are themselves off by tens of percent on some tensors, e.g.
        # the cross-attention LayerNorm weights); drop-in side exactly as the trainer runs it: under fp16 autocast
        l_ref, a_ref, b_ref, g_ref = one_step(R.SwinWNet(error_matrix=True, depths=D2), which, False)
        l_me, a_me, b_me, g_me = one_step(S.SwinWNet(error_matrix=True, depths=D2), which, True)
        print(f"{which} step: loss ref {l_ref:.6f} / drop-in {l_me:.6f}; parts {a_ref:.5f},{b_ref:.5f} / {a_me:.5f},{b_me:.5f}")
        assert abs(l_me - l_ref) <= 1e-2 * abs(l_ref), which
        worst, n = 0.0, 0
        for name, gr in g_ref.items():
            gm = g_me[name]
            assert (gr is None) == (gm is None), (which, name)          # same unused-parameter pattern (even: ca_sr_to_seg)
            if gr is None or gr.abs().max() == 0:
                continue
            e = (gm - gr).abs().max().item() / gr.abs().max().item()
            worst, n = max(worst, e), n + 1
            assert e <= 0.1, (which, name, e)
        print(f"{which} step: {n} gradient tensors compared, worst max-norm relative difference {worst:.3e}")
        if which == "even":
            assert g_me["ca_sr_to_seg.blocks.0.gamma"] is None and g_me["ca_seg_to_sr.blocks.0.gamma"] is not None

    model = S.SwinWNet(error_matrix=True, depths=D2)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV)
    t = SegT(model, loader, loader, DEV, num_epochs=3, warmup_epochs=0, lr=1e-3, use_fp16=True, verbose=False)
    t.optimizer = S.train.FusedAdamW(filter(lambda p: p.requires_grad, model.parameters()), lr=1e-3, weight_decay=1e-4)
    t.scheduler = t._build_default_scheduler()
    assert not any(p.requires_grad for p in model.upscaler_encoder.parameters())      # Segmentator_pretrain.py:74-93
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    h = t.train()
    print("segmentator epochs (drop-in):", h["train_loss"])
    assert all(v == v and abs(v) < 1e6 for v in h["train_loss"]) and h["train_loss"][-1] < h["train_loss"][0]
    for n, p in model.named_parameters():                                             # frozen branches untouched
        if n.startswith(("upscaler_", "ca_")):
            assert torch.equal(p, before[n]), n
    for p in model.parameters():
        p.requires_grad = True
    f = FullT(model, loader, loader, DEV, num_epochs=2, warmup_epochs=0, lr=3e-4, verbose=False)
    f.optimizer = S.train.FusedAdamW(model.parameters(), lr=3e-4, weight_decay=1e-4)
    f.scheduler = f._build_default_scheduler()
    m0 = f._run_epoch(0, train=True)                                                  # 4 batches: even, odd, even, odd
    m1 = f._run_epoch(1, train=True)
    print("full-model epochs (drop-in):", m0, m1)
    assert all(v == v for v in m1.values()) and m1["loss"] < m0["loss"] * 1.05
