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
