"""Seeded synthetic PLINK bed generators.

Balding-Nichols genotypes: ``K_pop`` populations, ancestral p ~ U(0.05, 0.95),
F_ST = ``fst``, genotypes Binomial(2, p_pop). Codes follow the PLINK bed contract the
reference decodes (src/Common.hpp:49-59, src/FilePlink.cpp:39-47): 00 = 2 copies of A1,
10 = 1 copy, 11 = 0 copies, 01 = missing; sample 4q+r sits in bits 2r..2r+1 of byte q of
its SNP; file header 6c 1b 01 (src/FilePlink.hpp:24-27).

numpy path: tests / small fixtures (writes .bed/.bim/.fam triplets).
torch path: bench-sized packed matrices generated directly in HBM.
"""
from __future__ import annotations

import os

import numpy as np

BED_MAGIC = bytes([0x6C, 0x1B, 0x01])
# copies of A1 (0,1,2) -> 2-bit bed code
_COPIES_TO_CODE = np.array([3, 2, 0], dtype=np.uint8)


def bytes_per_snp(n_samples: int) -> int:
    return (n_samples + 3) >> 2  # src/FilePlink.hpp:19


def pack_codes(codes: np.ndarray, pad_code: int = 0) -> np.ndarray:
    """codes: (M, N) uint8 in 0..3 -> packed (M, ceil(N/4)) uint8, LSB first."""
    M, N = codes.shape
    bpr = bytes_per_snp(N)
    padded = np.full((M, bpr * 4), pad_code, dtype=np.uint8)
    padded[:, :N] = codes
    q = padded.reshape(M, bpr, 4)
    return (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8)


def unpack_codes(packed: np.ndarray, n_samples: int) -> np.ndarray:
    """packed (M, bpr) uint8 -> codes (M, N) uint8."""
    M = packed.shape[0]
    out = np.empty((M, packed.shape[1], 4), dtype=np.uint8)
    for r in range(4):
        out[:, :, r] = (packed >> (2 * r)) & 3
    return out.reshape(M, -1)[:, :n_samples]


def balding_nichols_codes(n_samples, n_snps, k_pop=4, fst=0.1, miss=0.0, seed=1, chunk=8192):
    """Yield (snp_start, codes[(chunk, N) uint8]) in SNP chunks."""
    rng = np.random.default_rng(seed)
    pop = (np.arange(n_samples) * k_pop // n_samples).astype(np.int64)
    for s in range(0, n_snps, chunk):
        m = min(chunk, n_snps - s)
        p_anc = rng.uniform(0.05, 0.95, size=m)
        a = p_anc * (1 - fst) / fst
        b = (1 - p_anc) * (1 - fst) / fst
        p_pop = rng.beta(a[None, :].repeat(k_pop, 0), b[None, :].repeat(k_pop, 0))  # (K, m)
        p_ind = p_pop[pop, :].T  # (m, N)
        copies = rng.binomial(2, p_ind).astype(np.uint8)
        codes = _COPIES_TO_CODE[copies]
        if miss > 0:
            codes[rng.random(size=codes.shape) < miss] = 1
        yield s, codes


def sample_pops(n_samples, k_pop):
    return (np.arange(n_samples) * k_pop // n_samples).astype(np.int64)


def write_bed(prefix, n_samples, n_snps, k_pop=4, fst=0.1, miss=0.0, seed=1, chunk=8192,
              pad_code=0, bp_step=100):
    """Write <prefix>.bed/.bim/.fam; returns the packed matrix (M, bpr) uint8."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    bpr = bytes_per_snp(n_samples)
    packed = np.empty((n_snps, bpr), dtype=np.uint8)
    for s, codes in balding_nichols_codes(n_samples, n_snps, k_pop, fst, miss, seed, chunk):
        packed[s:s + codes.shape[0]] = pack_codes(codes, pad_code)
    write_bed_from_packed(prefix, packed, n_samples, k_pop=k_pop, bp_step=bp_step)
    return packed


def write_bed_from_packed(prefix, packed, n_samples, k_pop=1, bp_step=100):
    n_snps = packed.shape[0]
    with open(prefix + ".bed", "wb") as f:
        f.write(BED_MAGIC)
        f.write(np.ascontiguousarray(packed).tobytes())
    pops = sample_pops(n_samples, k_pop)
    with open(prefix + ".fam", "w") as f:
        for i in range(n_samples):
            f.write(f"pop{pops[i]} s{i} 0 0 0 -9\n")
    # chromosomes 1..22 in equal contiguous runs, positions bp_step apart within each
    per_chr = (n_snps + 21) // 22
    with open(prefix + ".bim", "w") as f:
        for j in range(n_snps):
            c = j // per_chr + 1
            pos = (j % per_chr + 1) * bp_step
            f.write(f"{c}\trs{j}\t0\t{pos}\tA\tC\n")


def read_bed(prefix):
    """Return (packed (M, bpr) uint8, N, M)."""
    with open(prefix + ".fam") as f:
        n = sum(1 for _ in f)
    with open(prefix + ".bim") as f:
        m = sum(1 for _ in f)
    raw = np.fromfile(prefix + ".bed", dtype=np.uint8)
    if raw[:3].tobytes() != BED_MAGIC:
        raise ValueError("Incorrect magic number in plink bed file.")
    bpr = bytes_per_snp(n)
    return raw[3:3 + m * bpr].reshape(m, bpr).copy(), n, m


def torch_packed(n_samples, n_snps, k_pop=4, fst=0.1, miss=0.0, seed=1, device="cuda", chunk=1 << 16):
    """Packed (M, bpr) uint8 tensor generated on ``device`` (bench-sized inputs).

    Same population model as the numpy path but with torch's RNG, so the two streams
    are not bit-identical to each other; each is deterministic for its seed.
    """
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    bpr = bytes_per_snp(n_samples)
    npad = bpr * 4
    out = torch.empty((n_snps, bpr), dtype=torch.uint8, device=device)
    pop = (torch.arange(npad, device=device) * k_pop // max(n_samples, 1)).clamp_(max=k_pop - 1)
    valid = (torch.arange(npad, device=device) < n_samples)
    lut = torch.tensor([3, 2, 0], dtype=torch.uint8, device=device)
    for s in range(0, n_snps, chunk):
        m = min(chunk, n_snps - s)
        p_anc = torch.rand(m, generator=g, device=device) * 0.9 + 0.05
        # normal approximation of the Balding-Nichols beta, clipped into (0.01, 0.99)
        sd = torch.sqrt(p_anc * (1 - p_anc) * fst)
        p_pop = (p_anc[:, None] + sd[:, None] * torch.randn(m, k_pop, generator=g, device=device)).clamp_(0.01, 0.99)
        p_ind = p_pop[:, pop]  # (m, npad)
        u1 = torch.rand(m, npad, generator=g, device=device)
        u2 = torch.rand(m, npad, generator=g, device=device)
        copies = (u1 < p_ind).to(torch.uint8) + (u2 < p_ind).to(torch.uint8)
        codes = lut[copies.long()]
        if miss > 0:
            codes = torch.where(torch.rand(m, npad, generator=g, device=device) < miss,
                                torch.ones_like(codes), codes)
        codes = codes * valid.to(torch.uint8)  # padding bits 0
        q = codes.view(m, bpr, 4)
        out[s:s + m] = q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)
    return out


def torch_packed_tile(n_total, samp0, n_loc, snp0, m, k_pop=4, fst=0.1, miss=0.0, seed=1, device="cuda"):
    """Packed (m, ceil(n_loc/4)) uint8 tile of ONE fixed synthetic matrix: SNPs [snp0, snp0+m) x samples
    [samp0, samp0+n_loc) of an n_total-sample job (samp0 % 4 == 0). The allele frequencies of a SNP
    depend on (seed, SNP) only and the genotype draws on (seed, SNP chunk start, samp0), so the matrix
    is the same however it is cut into tiles of these boundaries — what a strong-scaling run needs:
    every world size sees the same bed. Callers keep snp0 and samp0 on a fixed grid."""
    import torch

    assert samp0 % 4 == 0
    g1 = torch.Generator(device=device)
    g1.manual_seed((seed * 1_000_003 + snp0) & 0x7FFFFFFFFFFF)
    g2 = torch.Generator(device=device)
    g2.manual_seed(((seed * 1_000_003 + snp0) * 7919 + samp0 + 1) & 0x7FFFFFFFFFFF)
    bpr = bytes_per_snp(n_loc)
    npad = bpr * 4
    gidx = torch.arange(npad, device=device) + samp0
    pop = (gidx * k_pop // max(n_total, 1)).clamp_(max=k_pop - 1)
    valid = (torch.arange(npad, device=device) < n_loc)
    lut = torch.tensor([3, 2, 0], dtype=torch.uint8, device=device)
    p_anc = torch.rand(m, generator=g1, device=device) * 0.9 + 0.05
    sd = torch.sqrt(p_anc * (1 - p_anc) * fst)
    p_pop = (p_anc[:, None] + sd[:, None] * torch.randn(m, k_pop, generator=g1, device=device)).clamp_(0.01, 0.99)
    p_ind = p_pop[:, pop]  # (m, npad)
    u1 = torch.rand(m, npad, generator=g2, device=device)
    copies = (u1 < p_ind).to(torch.uint8)
    u1 = torch.rand(m, npad, generator=g2, device=device)
    copies += (u1 < p_ind).to(torch.uint8)
    codes = lut[copies.long()]
    if miss > 0:
        codes = torch.where(torch.rand(m, npad, generator=g2, device=device) < miss, torch.ones_like(codes), codes)
    codes = codes * valid.to(torch.uint8)  # padding bits 0
    q = codes.view(m, bpr, 4)
    return q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)
