"""Host-side weight packing: fp32 nn.Parameters -> bf16 tile streams in the exact shared-memory image
the tcgen05 kernels consume (UMMA K-major SWIZZLE_128B; see csrc/common.cuh::sw128_offset).

A tile is [R rows x 64 K-elements] bf16 = R*128 bytes; row r lives at r*128 and its eight 16-byte chunks
are stored at chunk index (c ^ (r & 7)).  The kernels copy tiles with 1-D bulk copies (cp.async.bulk), so
no tensor maps are needed and the stream order is the MMA consumption order."""
import torch


def OPERAND_DTYPE():
    from . import ops
    return ops.operand_dtype()


def ceil_to(x, m):
    return (x + m - 1) // m * m


def swizzle_tiles(t: torch.Tensor, dtype=None) -> torch.Tensor:
    """t: [..., R, 64] (any float dtype) -> 16-bit operand dtype, same shape, 16-byte chunks XOR-permuted per row."""
    R = t.shape[-2]
    assert t.shape[-1] == 64
    t = t.to(dtype or OPERAND_DTYPE()).reshape(*t.shape[:-1], 8, 8)
    rows = torch.arange(R, device=t.device)
    idx = torch.arange(8, device=t.device)[None, :] ^ (rows[:, None] & 7)          # [R, 8]: out chunk j <- in chunk j^(r&7)
    idx = idx.view(*([1] * (t.dim() - 3)), R, 8, 1).expand(*t.shape)
    return torch.gather(t, -2, idx).reshape(*t.shape[:-2], 64)


def unswizzle_tiles(t: torch.Tensor) -> torch.Tensor:
    """inverse of swizzle_tiles (XOR is an involution) — used by the CPU layout tests."""
    return swizzle_tiles(t.float(), t.dtype).to(t.dtype)


def choose_chunk(N: int, max_nt: int = 256) -> int:
    """largest divisor d of N with d % 4 == 0 and d <= max_nt (columns per accumulator chunk)."""
    best = 0
    for d in range(4, min(N, max_nt) + 1, 4):
        if N % d == 0:
            best = d
    if best == 0:
        raise ValueError(f"no valid chunk for N={N}")
    return best


def pack_rowgemm(W: torch.Tensor, bias, n_valid: int):
    """W [N, K] (nn.Linear layout) -> (Wp bf16 [nchunks*KB*NT*64], bias_padded f32 [nchunks*NT] | None, NT, nchunks)."""
    N, K = W.shape
    assert N % n_valid == 0 and n_valid % 4 == 0
    nchunks = N // n_valid
    NT = ceil_to(n_valid, 16)
    KB = ceil_to(ceil_to(K, 16), 64) // 64
    Wz = torch.zeros(nchunks, NT, KB * 64, device=W.device, dtype=torch.float32)
    Wz[:, :n_valid, :K] = W.detach().float().view(nchunks, n_valid, K)
    tiles = Wz.view(nchunks, NT, KB, 64).permute(0, 2, 1, 3).contiguous()          # (chunk, kb, NT, 64)
    Wp = swizzle_tiles(tiles).reshape(-1).contiguous()
    bp = None
    if bias is not None:
        bp = torch.zeros(nchunks, NT, device=W.device, dtype=torch.float32)
        bp[:, :n_valid] = bias.detach().float().view(nchunks, n_valid)
        bp = bp.reshape(-1).contiguous()
    return Wp, bp, NT, nchunks


def pack_mlp(W1: torch.Tensor, W2: torch.Tensor, b2: torch.Tensor, HC: int, TR: int):
    """fc1 W1 [4C, C], fc2 W2 [C, 4C] -> tile stream in the order csrc/mlp.cu consumes it:
    G1(0), [G1(j+1), G2(j)] for j = 0..nj-2, G2(nj-1); G1(j) = KB1 tiles [HC x 64], G2(j) = nkk*nT tiles
    [TR x 64] ordered (kk, t).  Returns (Wp bf16 flat, b2 padded to ceil16(C))."""
    Hd, C = W1.shape
    C16 = ceil_to(C, 16)
    KB1 = ceil_to(C16, 64) // 64
    nj = Hd // HC
    nkk = ceil_to(HC, 64) // 64
    nT = C16 // TR
    dev = W1.device
    W1z = torch.zeros(Hd, KB1 * 64, device=dev)
    W1z[:, :C] = W1.detach().float()
    g1 = swizzle_tiles(W1z.view(nj, HC, KB1, 64).permute(0, 2, 1, 3).contiguous())           # [nj, KB1, HC, 64]
    W2z = torch.zeros(nT * TR, nj, nkk * 64, device=dev)
    W2z[:C, :, :HC] = W2.detach().float().view(C, nj, HC)
    g2 = swizzle_tiles(W2z.view(nT, TR, nj, nkk, 64).permute(2, 3, 0, 1, 4).contiguous())    # [nj, nkk, nT, TR, 64]
    parts = [g1[0].reshape(-1)]
    for j in range(nj):
        if j + 1 < nj:
            parts.append(g1[j + 1].reshape(-1))
        parts.append(g2[j].reshape(-1))
    b2p = torch.zeros(C16, device=dev)
    b2p[:C] = b2.detach().float()
    return torch.cat(parts).contiguous(), b2p


def fused_block_geometry(C: int):
    """(K16, NQ, HC, nj, ones_col) of csrc/swin_fused.cu::launch_swin_fused for channel width C."""
    K16 = ceil_to(C, 16)
    ones_col = ceil_to(3 * C, 8)
    NQ = ceil_to(ones_col + 8, 16)
    HC = 64 if (4 * C) % 64 == 0 else 48
    if (4 * C) % HC:
        raise ValueError(f"fused block: hidden width {4 * C} not divisible into {HC}-column chunks")
    return K16, NQ, HC, (4 * C) // HC, ones_col


def rel_pos_bias_fragments(table: torch.Tensor, scale: float) -> torch.Tensor:
    """relative_position_bias_table [81, nH] -> [nH, 2, 4, 32, 4] fp32: for every head the 32x32 (25x25 real) bias
    matrix laid out as the mma.sync m16n8 accumulator fragments of the attention core (m-tile, n-tile, lane, element):
    row = mt*16 + lane//4 + (e//2)*8, key = nt*8 + (lane%4)*2 + e%2.  Key columns 25..31 carry -1e30 (masked)."""
    nH = table.shape[1]
    dev = table.device
    t = torch.arange(25, device=dev)
    R = (t // 5) * 9 + t % 5
    idx = R[:, None] - R[None, :] + 40                                     # [25, 25] (SwinWNet.py:163-173)
    full = torch.zeros(nH, 32, 32, device=dev)
    full[:, :, 25:] = -1e30
    full[:, :25, :25] = table.detach().float().t()[:, idx] * scale         # [nH, 25, 25]
    mt, nt, lane, e = torch.meshgrid(torch.arange(2, device=dev), torch.arange(4, device=dev),
                                     torch.arange(32, device=dev), torch.arange(4, device=dev), indexing="ij")
    row = mt * 16 + lane // 4 + (e // 2) * 8
    key = nt * 8 + (lane % 4) * 2 + e % 2
    return full[:, row, key].contiguous()                                  # [nH, 2, 4, 32, 4]


def pack_fused_block(n1w, n1b, Wqkv, bqkv, table, Wproj, bproj, n2w, n2b, W1, b1, W2, b2, num_heads: int):
    """Parameters of one SwinTransformerBlock (nn.Module layouts) -> (Wpk 16-bit flat, fpk fp32 flat) for
    csrc/swin_fused.cu.

    Wpk = [Wqkv | Wproj | W1 chunk j | W2 chunk j] as [rows x 64] SWIZZLE_128B k-block images (k-block major inside
    each matrix).  Algebraic folds done here in fp32 (exact up to the final 16-bit rounding of the operands):
      * norm1 / norm2 affine into the following linear:  W' = W diag(gamma),  b' = b + W beta  (the kernel only
        computes (x - mean) * rstd); zero-padded window tokens must still produce the PLAIN qkv bias, which is kept
        as a second vector;
      * head_dim^-0.5 * log2(e) into the q rows / q bias and log2(e) into the relative-position bias, so the kernel's
        softmax is a bare exp2;
      * rows [ones_col, ones_col+8) of the padded qkv weight are zero with bias 1: the resulting block of ones behind
        q|k|v turns the softmax denominator into one more column tile of the P V mma.
    fpk = [bqkv' (NQ) | bqkv plain (NQ) | bproj (K16) | b1' (4C) | b2 (K16) | bias fragments (nH*1024)]."""
    C = Wqkv.shape[1]
    K16, NQ, HC, nj, ones_col = fused_block_geometry(C)
    KB = ceil_to(K16, 64) // 64
    dev = Wqkv.device
    LOG2E = 1.4426950408889634
    qs = (C // num_heads) ** -0.5 * LOG2E
    f = lambda t: t.detach().float()

    def kmajor(Wm, rows):          # [N, K<=KB*64] -> [KB, rows, 64] swizzled
        z = torch.zeros(rows, KB * 64, device=dev)
        z[:Wm.shape[0], :Wm.shape[1]] = Wm
        return swizzle_tiles(z.view(rows, KB, 64).permute(1, 0, 2).contiguous()).reshape(-1)

    def padv(v, n, fill=0.0):
        z = torch.full((n,), fill, device=dev)
        z[:v.numel()] = v.reshape(-1)
        return z

    Wq, bq_plain = f(Wqkv).clone(), f(bqkv).clone()
    Wq[:C] *= qs
    bq_plain[:C] *= qs
    bq_fold = bq_plain + Wq @ f(n1b)
    Wq = Wq * f(n1w)[None, :]
    W1f = f(W1) * f(n2w)[None, :]
    b1_fold = f(b1) + f(W1) @ f(n2b)
    W2f = f(W2)
    parts = [kmajor(Wq, NQ), kmajor(f(Wproj), K16)]
    for j in range(nj):
        parts.append(kmajor(W1f[j * HC:(j + 1) * HC], HC))
    for j in range(nj):
        z = torch.zeros(K16, 64, device=dev)
        z[:C, :HC] = W2f[:, j * HC:(j + 1) * HC]
        parts.append(swizzle_tiles(z).reshape(-1))
    Wpk = torch.cat(parts).contiguous()

    def qkv_bias(b):
        z = padv(b, NQ)
        z[ones_col:ones_col + 8] = 1.0
        return z
    fpk = torch.cat([qkv_bias(bq_fold), qkv_bias(bq_plain), padv(f(bproj), K16), padv(b1_fold, 4 * C), padv(f(b2), K16),
                     rel_pos_bias_fragments(f(table), LOG2E).reshape(-1)]).contiguous()
    return Wpk, fpk


def warp_block_geometry(C: int):
    """(K16, KT, NJ) of csrc/swin_warp.cu::WbGeom: padded row width, its 16-column k-steps, 8-column tiles with real channels"""
    K16 = ceil_to(C, 16)
    return K16, K16 // 16, ceil_to(C, 8) // 8


def mma_b_fragments(Wm: torch.Tensor) -> torch.Tensor:
    """Wm [N, K] (row n = output column, N % 8 == 0, K % 16 == 0) -> [K/16, N/8, 32, 4]: the B operand of
    mma.sync.m16n8k16 for k-step kt / column tile nt, per lane the pairs (k, k+1) and (k+8, k+9) with
    k = 16 kt + 2 (lane % 4), n = 8 nt + lane // 4."""
    N, K = Wm.shape
    dev = Wm.device
    kt, nt, lane, e = torch.meshgrid(torch.arange(K // 16, device=dev), torch.arange(N // 8, device=dev),
                                     torch.arange(32, device=dev), torch.arange(4, device=dev), indexing="ij")
    return Wm[nt * 8 + lane // 4, kt * 16 + (lane % 4) * 2 + e % 2 + (e // 2) * 8].contiguous()


def mma_a_fragments(Wm: torch.Tensor) -> torch.Tensor:
    """Wm [M, K] (M, K % 16 == 0) -> [M/16, K/16, 32, 8]: the A operand of mma.sync.m16n8k16, per lane the registers
    a0..a3 = (row g, k), (row g+8, k), (row g, k+8), (row g+8, k+8) with g = lane // 4, k = 16 kt + 2 (lane % 4), two
    consecutive k each."""
    M, K = Wm.shape
    dev = Wm.device
    mt, kt, lane, e = torch.meshgrid(torch.arange(M // 16, device=dev), torch.arange(K // 16, device=dev),
                                     torch.arange(32, device=dev), torch.arange(8, device=dev), indexing="ij")
    r = e // 2
    return Wm[mt * 16 + lane // 4 + (r % 2) * 8, kt * 16 + (lane % 4) * 2 + (r // 2) * 8 + e % 2].contiguous()


def pack_warp_block(n1w, n1b, Wqkv, bqkv, table, Wproj, bproj, n2w, n2b, W1, b1, W2, b2, num_heads: int):
    """Parameters of one SwinTransformerBlock (nn.Module layouts) -> (Wpk 16-bit flat, fpk fp32 flat) for
    csrc/swin_warp.cu (C = 12 / 24 / 48 with 3 heads, C = 48 with 6).

    Wpk = mma.sync fragments of [Wq | Wk | Wv (as A operand: v is produced transposed) | Wproj | W1 | W2], every matrix
    zero-padded to whole 8 / 16 tiles.  Folds done here in fp32: LayerNorm gamma into the columns of the following
    weight; head_dim^-0.5 log2(e) into q, log2(e) into the relative-position bias.  Biases:
      * C = 12 / 24 (spare k columns behind the channels): two extra k columns — column C multiplies a constant 1 (the
        plain bias), column C+1 multiplies 1 for real tokens and 0 for zero-padded window tokens (W beta: the padded tokens
        of the reference are zeros AFTER norm1, SwinWNet.py:242,254);
      * C = 48 (48 = 3 x 16, no spare column): fp32 vectors that initialise the accumulators — bq' bk' bv' (beta folded in,
        real tokens), bq bk bv (zero-padded tokens), bproj, b1' — appended to fpk.
    fpk = [b2 (K16) | bias fragments (nH * 1024) | (C = 48:) the 11 C bias floats]."""
    C = Wqkv.shape[1]
    if (C, num_heads) not in ((12, 3), (24, 3), (48, 3), (48, 6)):
        raise ValueError(f"pack_warp_block: C={C}, num_heads={num_heads} not supported (12/3, 24/3, 48/3, 48/6)")
    K16, KT, NJ = warp_block_geometry(C)
    biascol = K16 >= C + 2
    dev = Wqkv.device
    LOG2E = 1.4426950408889634
    qs = (C // num_heads) ** -0.5 * LOG2E
    f = lambda t: t.detach().float()

    def aug(Wm, b, gamma, beta, rows, split_beta):
        z = torch.zeros(rows, K16, device=dev)
        n = Wm.shape[0]
        z[:n, :C] = Wm if gamma is None else Wm * gamma[None, :]
        if biascol:
            z[:n, C] = b
            if beta is not None:
                if split_beta:
                    z[:n, C + 1] = Wm @ beta
                else:
                    z[:n, C] += Wm @ beta
        return z

    g1, be1, g2, be2 = f(n1w), f(n1b), f(n2w), f(n2b)
    Wq, Wk, Wv = f(Wqkv)[:C] * qs, f(Wqkv)[C:2 * C], f(Wqkv)[2 * C:]
    bq, bk, bv = f(bqkv)[:C] * qs, f(bqkv)[C:2 * C], f(bqkv)[2 * C:]
    W2p = torch.zeros(NJ * 8, 4 * C, device=dev)
    W2p[:C] = f(W2)
    parts = [mma_b_fragments(aug(Wq, bq, g1, be1, NJ * 8, True)), mma_b_fragments(aug(Wk, bk, g1, be1, NJ * 8, True)),
             mma_a_fragments(aug(Wv, bv, g1, be1, K16, True)), mma_b_fragments(aug(f(Wproj), f(bproj), None, None, NJ * 8, False)),
             mma_b_fragments(aug(f(W1), f(b1), g2, be2, 4 * C, False)), mma_b_fragments(W2p)]
    Wpk = torch.cat([t.reshape(-1) for t in parts])
    dt = OPERAND_DTYPE()
    if dt == torch.float16:
        Wpk = Wpk.clamp(-65504.0, 65504.0)
    Wpk = Wpk.to(dt).contiguous()
    b2p = torch.zeros(K16, device=dev)
    b2p[:C] = f(b2)
    fparts = [b2p, rel_pos_bias_fragments(f(table), LOG2E).reshape(-1)]
    if not biascol:
        fparts += [bq + Wq @ be1, bk + Wk @ be1, bv + Wv @ be1, bq, bk, bv, f(bproj), f(b1) + f(W1) @ be2]
    fpk = torch.cat(fparts).contiguous()
    return Wpk, fpk


def pack_fused_attn_stream(n1w, n1b, Wqkv, bqkv, table, Wproj, bproj, num_heads: int):
    """C = 96 attention half for csrc/swin_fused.cu::swin_attn_stream_kernel: the same folds as pack_fused_block (norm1
    affine, q scale * log2 e, ones block at column 288, bias fragment images), with two differences:
      * the qkv bias rides in the GEMM: k columns 96 / 97 of the packed Wqkv hold the folded bias (applied to valid
        tokens) / the plain bias (applied to zero-padded window tokens); the kernel puts the matching indicator
        columns into its A tile;
      * the weights are emitted as the six [rows x 64] SWIZZLE_128B tiles the kernel streams per token tile, in
        consumption order: Wqkv rows [0,160) k-block 0, k-block 1; rows [160,304) k-block 0, k-block 1; Wproj kb 0, 1.
    fpk = [bproj (96) | bias fragments (nH*1024)]."""
    C = Wqkv.shape[1]
    assert C == 96 and Wqkv.shape[0] == 288
    NQ, NQ0, ones_col = 304, 160, 288
    dev = Wqkv.device
    LOG2E = 1.4426950408889634
    qs = (C // num_heads) ** -0.5 * LOG2E
    f = lambda t: t.detach().float()
    Wq, bq_plain = f(Wqkv).clone(), f(bqkv).clone()
    Wq[:C] *= qs
    bq_plain[:C] *= qs
    bq_fold = bq_plain + Wq @ f(n1b)
    Wq = Wq * f(n1w)[None, :]
    Wz = torch.zeros(NQ, 128, device=dev)
    Wz[:3 * C, :C] = Wq
    Wz[:3 * C, C] = bq_fold
    Wz[:3 * C, C + 1] = bq_plain
    Wz[ones_col:ones_col + 8, C:C + 2] = 1.0
    Pz = torch.zeros(C, 128, device=dev)
    Pz[:, :C] = f(Wproj)
    tiles = [Wz[:NQ0, :64], Wz[:NQ0, 64:], Wz[NQ0:, :64], Wz[NQ0:, 64:], Pz[:, :64], Pz[:, 64:]]
    Wpk = torch.cat([swizzle_tiles(t.contiguous()).reshape(-1) for t in tiles]).contiguous()
    fpk = torch.cat([f(bproj).reshape(-1), rel_pos_bias_fragments(f(table), LOG2E).reshape(-1)]).contiguous()
    return Wpk, fpk
