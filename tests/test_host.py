"""CPU-only tests: host logic (weight packing, module/state_dict contract, launch planning) and the C-ABI
library surface.  No kernel is launched here."""
import ctypes
import os
import re

import pytest
import torch

import swinwnet_b200 as S
from swinwnet_b200 import _lib, packing
from oracle import swinwnet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D2 = [2, 2, 2, 2]
OPD = S.ops.operand_dtype()   # 16-bit tensor-core operand dtype of the built library (fp16 default)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "swinwnet_b200.h")).read()
    declared = set(re.findall(r"\b(swn_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"swn_rowgemm_args"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.swn_abi_version() == 1
    assert ctypes.sizeof(_lib.RowGemmArgs) == lib.swn_sizeof_rowgemm_args()


def test_mlp_config_host_call_and_errors():
    lib = _lib.load()
    for C, exp in ((12, (48, 16)), (24, (96, 32)), (48, (64, 48)), (96, (96, 96)), (192, (64, 96)), (384, (128, 128))):
        hc, tr = ctypes.c_int(), ctypes.c_int()
        assert lib.swn_mlp_config(C, ctypes.byref(hc), ctypes.byref(tr)) == 0
        assert (hc.value, tr.value) == exp, (C, hc.value, tr.value)
    hc, tr = ctypes.c_int(), ctypes.c_int()
    assert lib.swn_mlp_config(13, ctypes.byref(hc), ctypes.byref(tr)) != 0
    assert b"unsupported" in lib.swn_last_error()


@pytest.mark.parametrize("name,ctor", [
    ("wnet_em", lambda: S.SwinWNet(error_matrix=True, depths=D2)),
    ("wnet", lambda: S.SwinWNet(depths=D2)),
    ("unet", lambda: S.SwinUNet(depths=D2)),
    ("unetsr", lambda: S.SwinUNetSR(depths=D2)),
    ("wnet_em_default_depths", lambda: S.SwinWNet(error_matrix=True)),
])
def test_state_dict_contract_matches_reference(manifest, name, ctor):
    """same keys and shapes as the reference modules (606 tensors for the multimodal model)."""
    m = ctor()
    sd = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert sd == manifest[name]
    m.load_state_dict(O.make_state_dict(manifest[name], seed=0), strict=True)
    if name == "wnet_em":
        assert len(sd) == 606 and sum(p.numel() for p in m.parameters()) == 29159743
        for attr in ("patch_embed", "segmentator_encoder", "segmentator_bottleneck", "segmentator_decoder",
                     "segmentator_head", "ca_seg_to_sr", "ca_sr_to_seg", "upscaler_encoder", "upscaler_bottleneck",
                     "upscaler_decoder", "upscaler_head"):
            assert len(list(getattr(m, attr).parameters())) > 0
        assert torch.equal(m.segmentator_encoder.layers[0].blocks[0].attn.relative_position_index, O.rel_pos_index(5))


def _tile_element(flat, tile_rows, tile_idx, r, k):
    """read element (r,k) of tile `tile_idx` the way the UMMA SWIZZLE_128B descriptor addresses it"""
    base = tile_idx * tile_rows * 64
    off = (r * 128 + ((((k >> 3) ^ r) & 7) << 4) + (k & 7) * 2) // 2
    return flat[base + off]


def test_pack_rowgemm_layout():
    N, K, nv = 288, 96, 144
    W = torch.randn(N, K)
    b = torch.randn(N)
    Wp, bp, NT, nch = packing.pack_rowgemm(W, b, nv)
    assert (NT, nch) == (144, 2) and Wp.dtype == OPD and Wp.numel() == nch * 2 * NT * 64
    Wb = W.to(OPD)
    for (n, kb, r, k) in [(0, 0, 0, 0), (1, 1, 143, 31), (1, 0, 77, 63), (0, 1, 9, 17)]:
        assert _tile_element(Wp, NT, n * 2 + kb, r, k) == Wb[n * nv + r, kb * 64 + k]
    assert _tile_element(Wp, NT, 1, 5, 40) == 0          # k = 64+40 >= K: zero padding
    assert torch.equal(bp.view(nch, NT)[:, :nv].reshape(-1), b)
    # padded rows (n_valid < NT)
    Wp, bp, NT, nch = packing.pack_rowgemm(torch.randn(72, 24), None, 72)
    assert (NT, nch) == (80, 1) and bp is None and _tile_element(Wp, NT, 0, 75, 3) == 0


def test_pack_mlp_stream_order():
    C, HC, TR = 96, 128, 96
    W1, W2, b2 = torch.randn(4 * C, C), torch.randn(C, 4 * C), torch.randn(C)
    Wp, b2p = packing.pack_mlp(W1, W2, b2, HC, TR)
    KB1, nj, nkk, nT = 2, 3, 2, 1
    g1, g2 = KB1 * HC * 64, nkk * nT * TR * 64
    assert Wp.numel() == nj * (g1 + g2) and torch.equal(b2p, b2)
    # stream: G1(0) G1(1) G2(0) G1(2) G2(1) G2(2)
    off = {"g1_0": 0, "g1_1": g1, "g2_0": 2 * g1, "g1_2": 2 * g1 + g2, "g2_1": 3 * g1 + g2, "g2_2": 3 * g1 + 2 * g2}
    W1b, W2b = W1.to(OPD), W2.to(OPD)
    assert _tile_element(Wp[off["g1_2"]:], HC, 1, 100, 20) == W1b[2 * HC + 100, 64 + 20]
    assert _tile_element(Wp[off["g2_1"]:], TR, 1, 50, 7) == W2b[50, 1 * HC + 64 + 7]
    assert _tile_element(Wp[off["g2_0"]:], TR, 0, 95, 63) == W2b[95, 63]


def test_choose_chunk():
    assert [packing.choose_chunk(n) for n in (144, 288, 576, 1152, 72, 36, 48, 384, 768)] == \
        [144, 144, 192, 192, 72, 36, 48, 192, 256]
    assert packing.choose_chunk(384, 128) == 128


def test_forward_refuses_cpu_and_autograd():
    m = S.SwinUNet(depths=D2)
    with pytest.raises(RuntimeError, match="CUDA"):
        with torch.no_grad():
            m(torch.zeros(1, 1, 20, 20))
    pkg = os.path.dirname(S.__file__)
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, f)).read(), f"product file {f} must not reference the oracle"


def test_gelu_fast_formula_against_exact_erf_gelu():
    """csrc/common.cuh::gelu_fast (epilogue GELU of every tensor-core MLP): 0.5 x (1 + tanh(x (a0 + a1 x^2 + a2 x^4))) with
    x^2 clamped at 49, restated in float64 — within 3.2e-5 of the exact erf GELU everywhere, correct sign / saturation
    far outside the fitted range (the quartic turns over at |x| ~ 11 without the clamp)."""
    import math
    import numpy as np
    a0, a1, a2 = 7.97458471e-1, 3.70503451e-2, -3.58732362e-4
    x = np.concatenate([np.linspace(-12, 12, 480001), np.array([-1e4, -50.0, 50.0, 1e4])])
    x2 = np.minimum(x * x, 49.0)
    g = 0.5 * x * (1.0 + np.tanh(x * (a0 + x2 * (a1 + x2 * a2))))
    ref = 0.5 * x * (1.0 + np.vectorize(math.erf)(x / math.sqrt(2.0)))
    assert np.abs(g - ref).max() <= 3.2e-5


def test_checkpoint_ingestion_variants():
    """nested / DataParallel-prefixed checkpoints of either depth configuration load strict=True, with
    depths and modality inferred from the keys (reference viewer: inference_gui/swinwnet_viewer_gui.py:129-151)."""
    import json
    import os
    import torch
    import swinwnet_b200 as S
    import benchdata
    man = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "manifest.json")))
    from swinwnet_b200 import checkpoint as ck
    sd = benchdata.make_state_dict(man["wnet_em"], seed=5)
    wrapped = {"epoch": 3, "model_state_dict": {"module." + k: v for k, v in sd.items()}}
    got = ck.load_state_dict_any(wrapped)
    assert set(got) == set(sd) and ck.infer_error_matrix_flag_from_sd(got)
    assert ck.infer_depths_from_sd(got) == [2, 2, 2, 2]
    m = ck.build_model_from_checkpoint(wrapped)
    assert isinstance(m, S.SwinWNet) and all(torch.equal(v, sd[k]) for k, v in m.state_dict().items())
    sd6 = benchdata.make_state_dict(man["wnet_em_default_depths"], seed=6)
    assert ck.infer_depths_from_sd(sd6) == [2, 2, 6, 2]
    assert isinstance(ck.build_model_from_checkpoint({"state_dict": sd6}), S.SwinWNet)
    assert not ck.infer_error_matrix_flag_from_sd(benchdata.make_state_dict(man["wnet"], seed=7))
    assert isinstance(ck.build_model_from_checkpoint(benchdata.make_state_dict(man["unet"], seed=8)), S.SwinUNet)
    assert isinstance(ck.build_model_from_checkpoint(benchdata.make_state_dict(man["unetsr"], seed=9)), S.SwinUNetSR)
    import pytest
    with pytest.raises(ValueError):
        ck.load_state_dict_any([1, 2, 3])


def test_fused_block_packing_folds_are_exact_algebra():
    """packing.pack_fused_block (host side of csrc/swin_fused.cu): un-swizzling the packed tiles and replaying the kernel's
    arithmetic in fp32 torch — (x - mean) * rstd against the folded weights, exp2 softmax against the bias-fragment
    images, row sums from the ones block — reproduces the oracle block's qkv, attention logits and fc1 pre-activations."""
    import math
    import torch
    from swinwnet_b200 import packing
    from oracle import swinwnet_oracle as O
    torch.manual_seed(0)
    C, nH = 24, 3
    hd = C // nH
    shp = [(C,), (C,), (3 * C, C), (3 * C,), (81, nH), (C, C), (C,), (C,), (C,), (4 * C, C), (4 * C,), (C, 4 * C), (C,)]
    n1w, n1b, Wqkv, bqkv, table, Wproj, bproj, n2w, n2b, W1, b1, W2, b2 = [torch.randn(*s) * 0.3 + (1.0 if i in (0, 7) else 0.0)
                                                                           for i, s in enumerate(shp)]
    Wpk, fpk = packing.pack_fused_block(n1w, n1b, Wqkv, bqkv, table, Wproj, bproj, n2w, n2b, W1, b1, W2, b2, nH)
    K16, NQ, HC, nj, ones = packing.fused_block_geometry(C)
    tiles = packing.unswizzle_tiles(Wpk[:NQ * 64].view(NQ, 64)).float()                 # Wqkv' image [NQ, 64]
    bq_fold, bq_plain = fpk[:NQ], fpk[NQ:2 * NQ]
    x = torch.randn(7, C) * 2 + 0.5
    xn = (x - x.mean(-1, keepdim=True)) * torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + 1e-5)
    qkv_k = xn @ tiles[:, :C].t() + bq_fold                                              # what the kernel's GEMM + bias gives
    ref = O.linear(O.layer_norm(x, n1w, n1b), Wqkv, bqkv)
    qs = hd ** -0.5 * math.log2(math.e)
    ref = torch.cat([ref[:, :C] * qs, ref[:, C:]], 1)
    tol = 2e-3 * ref.abs().max()                                                         # 16-bit rounding of the packed weights
    assert (qkv_k[:, :3 * C] - ref).abs().max() <= tol
    assert torch.equal(qkv_k[:, ones:ones + 8], torch.ones(7, 8))                        # ones block: zero weights, bias 1
    assert torch.allclose(bq_plain[:3 * C], torch.cat([bqkv[:C] * qs, bqkv[C:]]), atol=1e-6)   # q/k/v of zero-padded tokens
    # bias fragment images: element (mt, nt, lane, e) <-> (row, key) of the 32x32 tile
    frag = fpk[2 * NQ + 2 * K16 + 4 * C:].view(nH, 2, 4, 32, 4)
    idx = O.rel_pos_index()
    for (h, mt, nt, lane, e) in ((0, 0, 0, 0, 0), (1, 1, 2, 13, 3), (2, 0, 3, 31, 1), (2, 1, 1, 5, 2)):
        row, key = mt * 16 + lane // 4 + (e // 2) * 8, nt * 8 + (lane % 4) * 2 + e % 2
        want = -1e30 if key >= 25 else (0.0 if row >= 25 else table[idx[row, key], h].item() * math.log2(math.e))
        assert abs(frag[h, mt, nt, lane, e].item() - want) <= 1e-6 * max(1.0, abs(want))
    # fc1 with the norm2 affine folded in
    off = (NQ + K16) * 64
    W1img = packing.unswizzle_tiles(Wpk[off:off + nj * HC * 64].view(nj * HC, 64)).float()[:, :C]
    b1_fold = fpk[2 * NQ + K16:2 * NQ + K16 + 4 * C]
    h_k = xn @ W1img.t() + b1_fold
    h_ref = O.linear(O.layer_norm(x, n2w, n2b), W1, b1)
    assert (h_k - h_ref).abs().max() <= 2e-3 * h_ref.abs().max()


def test_gelu_pack2_half_pipeline_error_budget():
    """csrc/common.cuh::gelu_pack2 (fp16 build): numpy replay of its op sequence with a rounding to fp16 after every
    instruction (x -> fp16, x*x, min, two fmas, x*p, tanh, 0.5x, fma).  Against the exact erf GELU its rms error stays
    within 1.8x that of "fp32 evaluation followed by one rounding to fp16", the form it replaced, for pre-activation
    scales 0.5 .. 4."""
    import numpy as np
    from math import erf, sqrt
    rng = np.random.default_rng(0)
    verf = np.vectorize(erf)
    f16 = np.float16

    def fma16(a, b, c):
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f16)
    for sd in (0.5, 1.0, 2.0, 4.0):
        x = rng.normal(0, sd, 200_000).astype(np.float32)
        exact = x.astype(np.float64) * 0.5 * (1 + verf(x.astype(np.float64) / sqrt(2.0)))
        xh = x.astype(f16)
        x2 = np.minimum((xh.astype(np.float32) * xh.astype(np.float32)).astype(f16), f16(49))
        p = fma16(np.full_like(xh, f16(-3.58732362e-4)), x2, np.full_like(xh, f16(3.70503451e-2)))
        p = fma16(p, x2, np.full_like(xh, f16(7.97458471e-1)))
        u = (xh.astype(np.float32) * p.astype(np.float32)).astype(f16)
        th = np.tanh(u.astype(np.float64)).astype(f16)
        hx = (xh.astype(np.float32) * np.float32(0.5)).astype(f16)
        got = fma16(hx, th, hx).astype(np.float64)
        one_rounding = exact.astype(f16).astype(np.float64)
        assert (got - exact).std() <= 1.8 * (one_rounding - exact).std() + 1e-6
        assert np.abs(got - exact).max() <= 2.5e-3 * max(1.0, sd)


def test_npy_batch_ingestion_and_pathlike_checkpoint(tmp_path):
    """f-4: .npy batches in the viewer's formats (2-D / 3-D / 4-D arrays, dict files) -> [B,C,H,W]; checkpoints given as
    pathlib.Path."""
    import numpy as np
    from swinwnet_b200 import checkpoint as CK
    a2, a3 = np.random.rand(6, 8).astype(np.float32), np.random.rand(3, 6, 8)
    np.save(tmp_path / "one.npy", a2)
    np.save(tmp_path / "three.npy", a3)
    np.save(tmp_path / "d.npy", {"images": np.random.rand(2, 1, 6, 8), "other": 1}, allow_pickle=True)
    x = CK.load_npy_batch([tmp_path / "one.npy", tmp_path / "three.npy", tmp_path / "d.npy"], pin=False)
    assert x.shape == (6, 1, 6, 8) and x.dtype == torch.float32
    assert torch.equal(x[0, 0], torch.from_numpy(a2)) and torch.allclose(x[1:4, 0], torch.from_numpy(a3).float())
    np.save(tmp_path / "bad.npy", np.zeros((2, 5, 8)))
    with pytest.raises(ValueError):
        CK.load_npy_batch([tmp_path / "one.npy", tmp_path / "bad.npy"], pin=False)
    with pytest.raises(ValueError):
        CK.as_4d(np.zeros((1, 2, 3, 4, 5)))
    m = S.SwinUNet(depths=[2, 2, 2, 2])
    torch.save({"state_dict": {"module." + k: v for k, v in m.state_dict().items()}}, tmp_path / "ck.pth")
    m2 = CK.build_model_from_checkpoint(tmp_path / "ck.pth")           # pathlib.Path
    assert type(m2).__name__ == "SwinUNet" and len(m2.state_dict()) == len(m.state_dict())


def test_warp_block_packing_replays_the_kernel_dataflow():
    """packing.pack_warp_block (host side of csrc/swin_warp.cu): a lane-level replay of the kernel's mma.sync fragment
    dataflow in fp32 torch — m16n8k16 operand / accumulator layouts written out independently here from the PTX
    definition — turns one 25-token window (20 real + 5 zero-padded tokens) into the oracle's q (scaled), k, v, the
    attention output, and the whole block output."""
    import math
    torch.manual_seed(1)
    lane = torch.arange(32)
    g, t = lane // 4, lane % 4

    def mma(Af, Bf, Cf):            # Af [32, 8] (a0..a3 pairs), Bf [32, 4] (b0, b1 pairs), Cf [32, 4]
        A, Bm = torch.zeros(16, 16), torch.zeros(16, 8)
        for r in range(4):
            for e in range(2):
                A[g + 8 * (r % 2), 2 * t + e + 8 * (r // 2)] = Af[:, 2 * r + e]
        for r in range(2):
            for e in range(2):
                Bm[2 * t + e + 8 * r, g] = Bf[:, 2 * r + e]
        Cm = A @ Bm
        out = Cf.clone()
        for i in range(4):
            out[:, i] += Cm[g + 8 * (i // 2), 2 * t + i % 2]
        return out

    for C, nH in ((12, 3), (24, 3), (48, 3), (48, 6)):
        hd = C // nH
        K16, KT, NJ = packing.warp_block_geometry(C)
        biascol = K16 >= C + 2
        shp = [(C,), (C,), (3 * C, C), (3 * C,), (81, nH), (C, C), (C,), (C,), (C,), (4 * C, C), (4 * C,), (C, 4 * C), (C,)]
        prm = [torch.randn(*s) * (s[-1] ** -0.5 if len(s) == 2 else 0.2) + (1.0 if i in (0, 7) else 0.0) for i, s in enumerate(shp)]
        n1w, n1b, Wqkv, bqkv, table, Wproj, bproj, n2w, n2b, W1, b1, W2, b2 = prm
        Wpk, fpk = packing.pack_warp_block(*prm, nH)
        Wpk = Wpk.float()
        sizes = [KT * NJ * 128, KT * NJ * 128, KT * KT * 256, KT * NJ * 128, KT * (C // 2) * 128, (C // 4) * NJ * 128]
        assert Wpk.numel() == sum(sizes) and fpk.numel() == K16 + nH * 1024 + (0 if biascol else 11 * C)   # WbGeom::W_ELEMS / F_ELEMS
        xb = fpk[K16 + nH * 1024:]
        bqf, bkf, bvf, bqp, bkp, bvp, bpj, b1f = (torch.split(xb, [C] * 7 + [4 * C]) if not biascol else [None] * 8)
        Q, K, V, P, F1, F2 = torch.split(Wpk, sizes)
        Q, K, P = [m.view(KT, NJ, 32, 4) for m in (Q, K, P)]
        V, F1, F2 = V.view(KT, KT, 32, 8), F1.view(KT, C // 2, 32, 4), F2.view(C // 4, NJ, 32, 4)
        bias_frag = fpk[K16:K16 + nH * 1024].view(nH, 2, 4, 32, 4)
        # a 4 x 5 image: one window, its fifth row is zero padding (AFTER norm1) — oracle reference
        H, W = 4, 5
        x = torch.randn(1, H * W, C) * 1.5 + 0.3
        sd = dict(zip(["norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.relative_position_bias_table",
                       "attn.proj.weight", "attn.proj.bias", "norm2.weight", "norm2.bias", "mlp.0.weight", "mlp.0.bias",
                       "mlp.3.weight", "mlp.3.bias"], prm))
        ref = O.swin_block(sd, "", x, (H, W), nH, 0)[0]
        # ---- the kernel, lane by lane: rows in accumulator layout xr[mt][j] = [32, 4] ----
        rows = torch.zeros(32, K16)
        rows[:20, :C] = x[0]
        real = torch.zeros(32, dtype=torch.bool)
        real[:20] = True

        def to_frag(m):             # [32 rows, K16] -> acc-layout list [mt][j] of [32 lanes, 4]
            return [[torch.stack([m[16 * mt + g + 8 * (i // 2), 8 * j + 2 * t + i % 2] for i in range(4)], 1)
                     for j in range(K16 // 8)] for mt in range(2)]

        def from_frag(fr, ncols):
            m = torch.zeros(32, ncols)
            for mt in range(2):
                for j in range(ncols // 8):
                    for i in range(4):
                        m[16 * mt + g + 8 * (i // 2), 8 * j + 2 * t + i % 2] = fr[mt][j][:, i]
            return m

        def a_frags(fr):            # acc-layout tiles -> A fragments [mt][kt] of [32, 8]
            return [[torch.cat([fr[mt][2 * kt][:, 0:2], fr[mt][2 * kt][:, 2:4], fr[mt][2 * kt + 1][:, 0:2], fr[mt][2 * kt + 1][:, 2:4]], 1)
                     for kt in range(KT)] for mt in range(2)]

        def layer_norm_frags(m, first):
            mu = m[:, :C].mean(-1, keepdim=True)
            xn = (m[:, :C] - mu) * torch.rsqrt(m[:, :C].var(-1, unbiased=False, keepdim=True) + 1e-5)
            z = torch.zeros(32, K16)
            z[:, :C] = xn * (real[:, None] if first else 1.0)
            if biascol:
                z[:, C] = 1.0
                z[:, C + 1] = real.float() if first else 0.0
            return a_frags(to_frag(z))

        def rowbias(fold, plain, ncols):      # accumulator start without bias columns: per row the folded / the plain bias
            m = torch.zeros(32, ncols)
            m[:, :C] = torch.where(real[:, None], fold[None, :], plain[None, :])
            return to_frag_n(m, ncols)

        def to_frag_n(m, ncols):
            return [[torch.stack([m[16 * mt + g + 8 * (i // 2), 8 * j + 2 * t + i % 2] for i in range(4)], 1)
                     for j in range(ncols // 8)] for mt in range(2)]

        a1 = layer_norm_frags(rows, True)
        zero = torch.zeros(32, 4)
        q = [[None] * NJ for _ in range(2)]
        k = [[None] * NJ for _ in range(2)]
        q0 = rowbias(bqf, bqp, NJ * 8) if not biascol else None
        k0 = rowbias(bkf, bkp, NJ * 8) if not biascol else None
        for n in range(NJ):
            for mt in range(2):
                qc, kc = (zero, zero) if biascol else (q0[mt][n], k0[mt][n])
                for kt in range(KT):
                    qc, kc = mma(a1[mt][kt], Q[kt, n], qc), mma(a1[mt][kt], K[kt, n], kc)
                q[mt][n], k[mt][n] = qc, kc
        qkv_ref = O.linear(torch.cat([O.layer_norm(x[0], n1w, n1b), torch.zeros(5, C)]), Wqkv, bqkv)   # padded tokens: zeros after norm1
        qs = hd ** -0.5 * math.log2(math.e)
        tol = 3e-3 * qkv_ref.abs().max()
        assert (from_frag(q, NJ * 8)[:25, :C] - qkv_ref[:, :C] * qs).abs().max() <= tol * qs
        assert (from_frag(k, NJ * 8)[:25, :C] - qkv_ref[:, C:2 * C]).abs().max() <= tol
        # v^T = Wv xn^T: A operand = weight fragments, B operand = the LayerNorm fragments of token tile nt
        vT = torch.zeros(K16, 32)
        for mv in range(KT):
            for nt in range(4):
                vc = zero
                if not biascol:     # rows = channels 16 mv + g (+8), columns = tokens 8 nt + 2t (+1)
                    tokreal = torch.stack([real[8 * nt + 2 * t + i % 2] for i in range(4)], 1)
                    ch = torch.stack([16 * mv + g + 8 * (i // 2) for i in range(4)], 1)
                    vc = torch.where(tokreal, bvf[ch], bvp[ch])
                for kt in range(KT):
                    af = a1[nt // 2][kt]
                    vc = mma(V[mv, kt], torch.cat([af[:, 2 * (nt % 2):2 * (nt % 2) + 2], af[:, 4 + 2 * (nt % 2):6 + 2 * (nt % 2)]], 1), vc)
                for i in range(4):
                    vT[16 * mv + g + 8 * (i // 2), 8 * nt + 2 * t + i % 2] = vc[:, i]
        assert (vT[:C, :25].t() - qkv_ref[:, 2 * C:]).abs().max() <= tol
        # softmax(q k^T + bias) v on the reconstructed matrices (the bias image carries log2(e) and the key mask)
        qm, km = from_frag(q, NJ * 8), from_frag(k, NJ * 8)
        o = torch.zeros(32, K16)
        for h in range(nH):
            bias = torch.zeros(32, 32)
            for mt in range(2):
                for nt in range(4):
                    for i in range(4):
                        bias[16 * mt + g + 8 * (i // 2), 8 * nt + 2 * t + i % 2] = bias_frag[h, mt, nt, :, i]
            s = qm[:, h * hd:(h + 1) * hd] @ km[:, h * hd:(h + 1) * hd].t() + bias
            pr = torch.exp2(s - s.max(-1, keepdim=True).values)
            o[:, h * hd:(h + 1) * hd] = (pr @ vT[h * hd:(h + 1) * hd].t()) / pr.sum(-1, keepdim=True)
        if biascol:
            o[:, C] = 1.0
        ao = a_frags(to_frag(o))
        rows_b = rows.clone()
        if not biascol:
            rows_b[:, :C] += bpj
        x1 = to_frag(rows_b)
        for mt in range(2):
            for j in range(NJ):
                for kt in range(KT):
                    x1[mt][j] = mma(ao[mt][kt], P[kt, j], x1[mt][j])
        x1m = torch.zeros(32, K16)
        x1m[:, :NJ * 8] = from_frag([r[:NJ] for r in x1], NJ * 8)
        a2 = layer_norm_frags(x1m, False)
        y = to_frag(x1m + torch.cat([fpk[:K16]])[None, :])
        for u in range(C // 4):
            for mt in range(2):
                h0, h1 = zero, zero
                if not biascol:
                    h0 = torch.stack([b1f[16 * u + 2 * t + i % 2] for i in range(4)], 1)
                    h1 = torch.stack([b1f[16 * u + 8 + 2 * t + i % 2] for i in range(4)], 1)
                for kt in range(KT):
                    h0, h1 = mma(a2[mt][kt], F1[kt, 2 * u], h0), mma(a2[mt][kt], F1[kt, 2 * u + 1], h1)
                ge = lambda v: torch.nn.functional.gelu(v)
                ah = torch.cat([ge(h0[:, 0:2]), ge(h0[:, 2:4]), ge(h1[:, 0:2]), ge(h1[:, 2:4])], 1)
                for j in range(NJ):
                    y[mt][j] = mma(ah, F2[u, j], y[mt][j])
        out = from_frag([r[:NJ] for r in y], NJ * 8)[:20, :C]
        assert (out - ref).abs().max() <= 5e-3 * ref.abs().max(), (C, (out - ref).abs().max(), ref.abs().max())
