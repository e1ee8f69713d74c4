"""Host-side mirror of PCAone's operator interface for the randomized-SVD path, on top of the
C-ABI (include/pcaone_b200.h). Same names and argument meaning as the reference so the parity
tests read like the reference's own call sites:

    Param                      src/Cmd.hpp:16-98 (the fields the hot path reads; derivations
                               of src/Cmd.cpp:141-238: oversamples=max(os,k), out_of_core=memory>0,
                               perm = winSVD && !no_shuffle)
    FileBed / Data.prepare     src/FilePlink.hpp:8-47, src/Data.cpp:14-85
    RsvdOpData                 src/Halko.hpp:6-42  (U, S, V, Omg, setFlags, initOmg, computeGandH,
                               computeUSV)
    NormalRsvdOpData / FancyRsvdOpData   src/Halko.hpp:44-93
    run_pca_with_halko         src/Halko.cpp:271-345
    permute_plink              src/FilePlink.cpp:303-408 (index map; rows are permuted in memory)

All arithmetic runs in the CUDA library; this module only plans blocks, owns host buffers
and converts a non-zero C status into the RuntimeError the reference would throw
(src/Logger.hpp:85-94).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .synth import BED_MAGIC, bytes_per_snp


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _vp(a):
    """void* of a numpy array or a torch tensor (host, pinned or device)."""
    if a is None:
        return None
    if _is_torch(a):
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


def _f(shape):
    return np.zeros(shape, dtype=np.float64, order="F")


@dataclass
class Param:
    """The hot-path subset of the reference's Param with the CLI spellings in comments."""
    filein: str = ""            # -b/--bfile
    fileout: str = "pcaone"     # -o
    k: int = 10                 # -k
    svd: int = 2                # -d/--svd (1: sSVD, 2: winSVD; 0: IRAM — only its block plan and operator)
    memory: float = 0.0         # -m GB; > 0 => out_of_core
    maxp: int = 20              # --maxp
    oversamples: int = 10       # --oversamples
    bands: int = 64             # -w/--batches
    tol: float = 1e-4           # --tol-rsvd
    seed: int = 112             # --seed
    rand: int = 1               # --rand
    no_shuffle: bool = False    # -S
    emu: bool = False           # --emu
    maf: float = 0.0            # --maf
    maxiter: int = 100          # --maxiter
    tolem: float = 1e-5         # --tol-em
    scale: int = -9             # -C
    haploid: bool = False       # --haploid
    ld: bool = False            # -D/--ld
    buffer: int = 2             # --buffer
    device: int = 0
    precision: int = _lib.PREC_FP64
    # derived (src/Cmd.cpp:141-238)
    out_of_core: bool = field(init=False, default=False)
    perm: bool = field(init=False, default=False)
    ploidy: int = field(init=False, default=2)

    def __post_init__(self):
        if self.svd not in (0, 1, 2):
            raise ValueError("only --svd 1 (sSVD), 2 (winSVD) and the operator of 0 (IRAM) are on the GPU path")
        if self.bands < 4 or self.bands % 2 != 0:
            raise ValueError("the -w/--batches must be a power of 2 and the minimun is 4.")
        self.oversamples = max(self.oversamples, self.k)
        self.out_of_core = self.memory > 0
        self.perm = self.svd == 2 and not self.no_shuffle
        self.ploidy = 1 if self.haploid else 2

    @property
    def l(self):
        return self.k + self.oversamples


def ooc_block_plan(N, M, l, memory_gb, winsvd, bands, iram=False):
    """Data::prepare, out-of-core branch (src/Data.cpp:42-84). `iram`: --svd 0, whose working set
    is only the N x blocksize block (src/Data.cpp:42-44)."""
    if iram:
        blocksize = int(math.ceil(memory_gb * 134217728 / N))
    else:
        m = float(3 * N * l + 2 * M * l + 5 * M) / 134217728
        if memory_gb > 1.1 * m:
            m = 0.0
        blocksize = int(math.ceil(((m + memory_gb) * 134217728 - 3 * N * l - 2 * M * l - 5 * M) / N))
    nblocks = int(math.ceil(M / blocksize))
    band_factor = 1
    if nblocks == 1:
        raise RuntimeError("only one block exists. please remove -m option")
    if winsvd:
        if nblocks < bands:
            blocksize = int(math.ceil(M / bands))
        else:
            band_factor = int(math.ceil(nblocks / bands))
            blocksize = int(math.ceil(M / (bands * band_factor)))
        nblocks = int(math.ceil(M / blocksize))
    start = np.arange(nblocks, dtype=np.uint64) * np.uint64(blocksize)
    stop = np.minimum(start + np.uint64(blocksize - 1), np.uint64(M - 1))
    return blocksize, nblocks, band_factor, start, stop


def permute_plink_indices(M, N, bands, gb=2):
    """PermMat of permute_plink (src/FilePlink.cpp:303-408): indices[new] = original."""
    bpr = bytes_per_snp(N)
    two = int(math.floor(1073741824.0 * gb / bpr))
    two = min(two, M)
    bufsize = two // bands
    two = bufsize * bands
    if two == 0:
        raise RuntimeError("permute_plink: fewer SNPs than bands")
    nblocks = (M + two - 1) // two
    modr2 = M % two
    modr = M % bands
    bandsize = (M + bands - 1) // bands
    bandidx = [i * bandsize if (modr == 0 or i < modr) else modr * bandsize + (bandsize - 1) * (i - modr)
               for i in range(bands)]
    indices = np.full(M, -1, dtype=np.int64)
    bufidx = bufsize
    for i in range(nblocks):
        if i == nblocks - 1 and modr2 != 0:
            two2 = M - (nblocks - 1) * two
            bufsize = (two2 + bands - 1) // bands
            modr2 = two2 % bands
        j = np.arange(bufsize - 1, dtype=np.int64)
        for b in range(bands):
            indices[i * bufidx + bandidx[b] + j] = i * two + j * bands + b
            jl = bufsize - 1
            if i != nblocks - 1 or b < modr2 or modr2 == 0:
                indices[i * bufidx + bandidx[b] + jl] = i * two + jl * bands + b
    return indices


class FileBed:
    """`Data` for PLINK bed input. In-core: the packed matrix goes to HBM once (read_all).
    Out-of-core: blocks are streamed from the .bed file (or a host array) every pass."""

    def __init__(self, params: Param, packed: np.ndarray | None = None, nsamples: int | None = None):
        self.params = params
        self.perm = None
        self.start = self.stop = None
        self.nblocks, self.blocksize, self.bandFactor = 1, 0, 1
        if packed is None:
            with open(params.filein + ".fam") as f:
                self.nsamples = sum(1 for _ in f)
            with open(params.filein + ".bim") as f:
                self.nsnps = sum(1 for _ in f)
            with open(params.filein + ".bed", "rb") as f:
                if f.read(3) != BED_MAGIC:
                    raise RuntimeError("Incorrect magic number in plink bed file.")
            self.packed = None
        else:
            # numpy array, or a torch uint8 tensor (CUDA: uploaded device-to-device; pinned
            # host: streamed by cudaMemcpyAsync in out-of-core mode)
            self.packed = packed.contiguous() if _is_torch(packed) else np.ascontiguousarray(packed, dtype=np.uint8)
            self.nsamples = int(nsamples)
            self.nsnps = int(self.packed.shape[0])
            if self.packed.shape[1] != bytes_per_snp(self.nsamples):
                raise RuntimeError("packed matrix does not have ceil(N/4) bytes per SNP")

    def _load_packed(self):
        if self.packed is None:
            raw = np.fromfile(self.params.filein + ".bed", dtype=np.uint8)
            self.packed = raw[3:].reshape(self.nsnps, bytes_per_snp(self.nsamples))
        return self.packed

    def prepare(self):
        p = self.params
        if not p.out_of_core:
            self._load_packed()
            return
        self.blocksize, self.nblocks, self.bandFactor, self.start, self.stop = ooc_block_plan(
            self.nsamples, self.nsnps, p.l, p.memory, p.svd == 2, p.bands, iram=p.svd == 0)
        if p.perm:
            # permute_plink writes <out>.perm.bed; here the rows are permuted in host memory
            self.perm = permute_plink_indices(self.nsnps, self.nsamples, p.bands, p.buffer)
            self.packed = np.ascontiguousarray(np.asarray(self._load_packed())[self.perm])


class FileBgen:
    """`Data` for BGEN input, in-core (src/FileBgen.cpp:15-72). The container parsing stays on the
    host side of the boundary: this class takes what `var.minor_allele_dosage()` yields for every
    variant — `dosages`, float32 [nsnps][nsamples], NaN = missing — and applies the `--maf` filter
    of FileBgen.cpp:42-45 once the device has computed the allele frequencies. Decode, mean
    imputation, centring and scaling are fused into the GEMM operand load on the device."""

    def __init__(self, params: Param, dosages):
        self.params = params
        if params.out_of_core:
            raise RuntimeError("FileBgen: only the in-core path (read_all) is mirrored")
        if params.precision != _lib.PREC_FP64:
            raise RuntimeError("FileBgen: dosages run on the FP64 kernels (precision = PREC_FP64)")
        self.dosages = dosages.contiguous() if _is_torch(dosages) else np.ascontiguousarray(dosages, dtype=np.float32)
        self.nsnps, self.nsamples = int(self.dosages.shape[0]), int(self.dosages.shape[1])
        self.packed = None
        self.perm = None
        self.start = self.stop = None
        self.nblocks, self.blocksize, self.bandFactor = 1, 0, 1
        self.keep = None

    def prepare(self):
        """Variant selection of FileBgen.cpp:42-45: a variant is kept iff af > params.maf (so even
        with the default --maf 0 an all-zero variant is dropped). The allele frequencies come from
        the device (k_dosage_af through a throw-away context); only the index selection is host
        logic."""
        L = _lib.load()
        p = self.params
        cfg = _lib.Config(nsamples=self.nsamples, nsnps=self.nsnps, nsnps_total=self.nsnps, k=1, oversamples=0, svd=1,
                          bands=p.bands, maxp=1, tol=0.0, ploidy=p.ploidy, scale=p.scale, emu=0, out_of_core=0,
                          precision=_lib.PREC_FP64, device=p.device, rank=0, world=1, maxiter=0, tolem=0.0)
        h = C.c_void_p()
        if L.pcaone_create(C.byref(cfg), C.byref(h)):
            raise RuntimeError(L.pcaone_last_error(None).decode())
        try:
            on_dev = _is_torch(self.dosages) and self.dosages.is_cuda
            af = np.zeros(self.nsnps)
            for call in (lambda: L.pcaone_upload_dosage(h, _vp(self.dosages), self.nsnps, int(on_dev)),
                         lambda: L.pcaone_allele_freq(h), lambda: L.pcaone_get_F(h, _vp(af))):
                if call():
                    raise RuntimeError(L.pcaone_last_error(h).decode())
        finally:
            L.pcaone_destroy(h)
        keep = np.flatnonzero(af > p.maf)
        if len(keep) == 0:
            raise RuntimeError("the number of SNPs after filtering is 0!")
        if len(keep) != self.nsnps:
            self.keep = keep
            if _is_torch(self.dosages):
                import torch
                self.dosages = self.dosages[torch.as_tensor(keep, device=self.dosages.device)].contiguous()
            else:
                self.dosages = np.ascontiguousarray(self.dosages[keep])
            self.nsnps = len(keep)


class FileBeagle:
    """`Data` for Beagle genotype-likelihood input, in-core (src/FileBeagle.cpp:14-68). The gz text
    parsing stays on the host side of the boundary: this class takes the matrix the reference's
    parse_beagle_file fills, P (2 * nsamples x nsnps, P[2i, j] / P[2i+1, j] = likelihoods of genotypes
    0 / 1). prepare() runs the allele-frequency EM on the device (emMAF_with_GL, Utils.cpp:745-775)
    and applies the `--maf` filter of Data::filter_snps_resize_F (Data.cpp:88-105)."""

    def __init__(self, params: Param, P, tolmaf: float = 1e-6):
        self.params = params
        if params.out_of_core:
            raise RuntimeError("doesn't support out-of-core PCAngsd algorithm")     # FileBeagle.cpp:77
        if params.precision != _lib.PREC_FP64:
            raise RuntimeError("FileBeagle: genotype likelihoods run on the FP64 kernels (precision = PREC_FP64)")
        self.P = np.asfortranarray(P, dtype=np.float64)
        if self.P.shape[0] % 2:
            raise RuntimeError("P must have two rows per sample")
        self.nsamples, self.nsnps = self.P.shape[0] // 2, int(self.P.shape[1])
        self.tolmaf = tolmaf
        self.packed = None
        self.perm = None
        self.start = self.stop = None
        self.nblocks, self.blocksize, self.bandFactor = 1, 0, 1
        self.keep = None
        self.F = None
        self.maf_iters = 0

    def prepare(self):
        L = _lib.load()
        p = self.params
        cfg = _lib.Config(nsamples=self.nsamples, nsnps=self.nsnps, nsnps_total=self.nsnps, k=1, oversamples=0, svd=1,
                          bands=p.bands, maxp=1, tol=0.0, ploidy=p.ploidy, scale=p.scale, emu=0, out_of_core=0,
                          precision=_lib.PREC_FP64, device=p.device, rank=0, world=1, maxiter=0, tolem=0.0)
        h = C.c_void_p()
        if L.pcaone_create(C.byref(cfg), C.byref(h)):
            raise RuntimeError(L.pcaone_last_error(None).decode())
        try:
            F = np.zeros(self.nsnps)
            it = C.c_int(0)
            for call in (lambda: L.pcaone_upload_gl(h, _vp(self.P), self.nsnps, 0),
                         lambda: L.pcaone_gl_em_maf(h, int(p.maxiter), C.c_double(self.tolmaf), C.byref(it)),
                         lambda: L.pcaone_get_F(h, _vp(F))):
                if call():
                    raise RuntimeError(L.pcaone_last_error(h).decode())
        finally:
            L.pcaone_destroy(h)
        self.maf_iters = it.value
        self.F = F
        if p.maf > 0:   # Data::filter_snps_resize_F: keep MAF(F) > maf
            if not 0 < p.maf <= 0.5:
                raise RuntimeError("--maf has to be between (0, 0.5)")
            keep = np.flatnonzero(np.minimum(F, 1 - F) > p.maf)
            if len(keep) < 1:
                raise RuntimeError("no SNPs left after filtering!")
            self.keep = keep
            self.P = np.asfortranarray(self.P[:, keep])
            self.F = F[keep]
            self.nsnps = len(keep)


class RsvdOpData:
    """Abstract op (src/Halko.hpp:6-42). Subclasses pick the computeGandH variant."""
    svd = None

    def __init__(self, data: FileBed, k: int, os_: int = 10, *, rank=0, world=1, nsnps_total=None,
                 allreduce=None, allreduce2=None, library_comm=False, shard_samples=False, nsamples_total=None, sample_offset=0):
        """Multi-GPU jobs (one op per GPU): `rank` / `world`, and either SNP sharding (data holds this
        rank's SNPs of every window, `nsnps_total` = M of the job) or `shard_samples` (data holds this
        rank's samples [sample_offset, sample_offset + data.nsamples) of ALL SNPs). `library_comm`: the
        collectives run inside the library (NCCL, id exchanged through torch.distributed); else through
        the `allreduce` host hook."""
        L = _lib.load()
        self.L = L
        self.data = data
        p = data.params
        self.nk, self.os = int(k), int(os_)
        self.update = False
        self.standardize = False
        self.U = self.S = self.V = None
        cfg = _lib.Config(nsamples=data.nsamples, nsnps=data.nsnps, nsnps_total=nsnps_total or data.nsnps,
                          k=self.nk, oversamples=self.os, svd=self.svd, bands=p.bands, maxp=p.maxp, tol=p.tol,
                          ploidy=p.ploidy, scale=p.scale, emu=int(p.emu), out_of_core=int(p.out_of_core),
                          precision=p.precision, device=p.device, rank=rank, world=world, maxiter=p.maxiter,
                          tolem=p.tolem, shard_samples=int(bool(shard_samples)),
                          nsamples_total=int(nsamples_total or data.nsamples), sample_offset=int(sample_offset))
        self.shard_samples = bool(shard_samples) and world > 1
        self.nsamples_total = int(nsamples_total or data.nsamples) if self.shard_samples else data.nsamples
        self.sample_offset = int(sample_offset) if self.shard_samples else 0
        h = C.c_void_p()
        if L.pcaone_create(C.byref(cfg), C.byref(h)):
            raise RuntimeError(L.pcaone_last_error(None).decode())
        self.h = h
        self._keep = []
        if library_comm and world > 1:
            from . import dist as _pdist
            _pdist.init_library_comm(L, self.h, rank, world, peer_mailboxes=self.shard_samples)
        if allreduce is not None:
            cb = _lib.ALLREDUCE_FN(allreduce)
            self._keep.append(cb)
            self._chk(L.pcaone_set_allreduce(self.h, cb, None))
        if allreduce2 is not None:
            cb = _lib.ALLREDUCE2_FN(allreduce2)
            self._keep.append(cb)
            self._chk(L.pcaone_set_allreduce2(self.h, cb, None))
        if getattr(data, "P", None) is not None:
            if data.F is None:
                raise RuntimeError("FileBeagle: call prepare() first")
            self._chk(L.pcaone_upload_gl(self.h, _vp(data.P), data.nsnps, 0))
            self._chk(L.pcaone_set_F(self.h, _vp(np.ascontiguousarray(data.F))))
        elif getattr(data, "dosages", None) is not None:
            on_dev = _is_torch(data.dosages) and data.dosages.is_cuda
            self._chk(L.pcaone_upload_dosage(self.h, _vp(data.dosages), data.nsnps, int(on_dev)))
            self._chk(L.pcaone_allele_freq(self.h))
        elif p.out_of_core:
            if data.packed is not None:
                self._chk(L.pcaone_set_host_source(self.h, _vp(data.packed), data.nsnps))
            else:
                self._chk(L.pcaone_open_bed(self.h, (p.filein + ".bed").encode(), 0))
            self._chk(L.pcaone_set_blocks(self.h, _vp(data.start), _vp(data.stop), data.nblocks, data.bandFactor))
        else:
            on_dev = _is_torch(data.packed) and data.packed.is_cuda
            self._chk(L.pcaone_upload_bed(self.h, _vp(data.packed), data.nsnps, int(on_dev)))
            if data.start is not None:  # explicit windows for a resident shard (SNP-sharded winSVD)
                self._chk(L.pcaone_set_blocks(self.h, _vp(data.start), _vp(data.stop), len(data.start),
                                              data.bandFactor))
            self._chk(L.pcaone_allele_freq(self.h))
        self._permuted = False
        self.initOmg()

    # -- plumbing
    def _chk(self, rc):
        if rc:
            raise RuntimeError(self.L.pcaone_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.pcaone_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference interface
    def rows(self):
        return self.data.nsnps

    def cols(self):
        return self.data.nsamples

    def ranks(self):
        return self.nk

    def oversamples(self):
        return self.os

    def size(self):
        return self.nk + self.os

    def setFlags(self, is_update, is_standardize):
        self.update, self.standardize = bool(is_update), bool(is_standardize)
        self._chk(self.L.pcaone_set_flags(self.h, int(self.update), int(self.standardize)))

    def initOmg(self):
        p = self.data.params
        if self.shard_samples:
            # the job's Omega is N_total x l (one RNG stream, column-major); this rank keeps its rows
            full = _f((self.nsamples_total, self.size()))
            self.L.pcaone_init_omega(self.nsamples_total, self.size(), p.seed, int(p.rand), _vp(full))
            self.Omg = np.asfortranarray(full[self.sample_offset:self.sample_offset + self.cols()])
            del full
        else:
            self.Omg = _f((self.cols(), self.size()))
            self.L.pcaone_init_omega(self.cols(), self.size(), p.seed, int(p.rand), _vp(self.Omg))
        self._chk(self.L.pcaone_set_omega(self.h, _vp(self.Omg)))

    def setOmg(self, Omg):
        self.Omg = np.asfortranarray(Omg, dtype=np.float64)
        self._chk(self.L.pcaone_set_omega(self.h, _vp(self.Omg)))

    def _maybe_permute(self):
        """in-core winSVD permutes the columns of G at pi == 0 of the first pass
        (src/Halko.cpp:183-186, src/RSVD.hpp:61-71)."""
        p = self.data.params
        if self.svd == _lib.SVD_WINSVD and p.perm and not p.out_of_core and not self._permuted:
            idx = np.zeros(self.rows(), dtype=np.uint32)
            self.L.pcaone_shuffle_indices(self.rows(), _vp(idx))
            self._chk(self.L.pcaone_permute_resident(self.h, _vp(idx)))
            self.data.perm = idx.astype(np.int64)
            self._permuted = True

    def getH(self, out=None):
        H = out if out is not None else _f((self.cols(), self.size()))
        self._chk(self.L.pcaone_get_GH(self.h, None, _vp(H)))
        return H

    def smallStage(self):
        """dense stage of one computeUSV epoch (src/Halko.cpp:55-70)."""
        self._chk(self.L.pcaone_small_stage(self.h))

    def computeGandH(self, pi, want=True):
        """One power-iteration pass; returns (G M x l, H N x l) col-major when want."""
        if pi == 0:
            self._maybe_permute()
        self._chk(self.L.pcaone_compute_gandh(self.h, int(pi)))
        if not want:
            return None
        G, H = _f((self.rows(), self.size())), _f((self.cols(), self.size()))
        self._chk(self.L.pcaone_get_GH(self.h, _vp(G), _vp(H)))
        return G, H

    def computeUSV(self, p, tol):
        self._maybe_permute()
        diff, ep = C.c_double(0), C.c_int(0)
        self._chk(self.L.pcaone_compute_usv(self.h, int(p), C.c_double(tol), C.byref(diff), C.byref(ep)))
        self.diff, self.epochs = diff.value, ep.value
        self._fetch_usv()

    def _fetch_usv(self):
        self.U, self.S, self.V = _f((self.cols(), self.nk)), np.zeros(self.nk), _f((self.rows(), self.nk))
        self._chk(self.L.pcaone_get_usv(self.h, _vp(self.U), _vp(self.S), _vp(self.V)))

    def runEM(self):
        self._maybe_permute()
        it = C.c_int(0)
        self._chk(self.L.pcaone_run_em(self.h, C.byref(it)))
        self._fetch_usv()
        return it.value

    # -- Data-side views used by the parity tests
    def F(self):
        out = np.zeros(self.rows())
        self._chk(self.L.pcaone_get_F(self.h, _vp(out)))
        return out

    def lookup(self):
        out = _f((4, self.rows()))
        self._chk(self.L.pcaone_get_lookup(self.h, _vp(out)))
        return out

    def missing_count(self):
        n = C.c_uint64(0)
        self._chk(self.L.pcaone_missing_count(self.h, C.byref(n)))
        return n.value

    def read_block(self, start, stop, standardize, update=False):
        """FileBed::read_block_initial / read_block_update output (N x B, col-major)."""
        out = _f((self.cols(), stop - start + 1))
        self._chk(self.L.pcaone_decode_block(self.h, int(start), int(stop), int(standardize), int(update), _vp(out)))
        return out

    def setF(self, F):
        """External allele frequencies (projection: the reference panel's .mbim column 7,
        Data::prepare for --project)."""
        F = np.ascontiguousarray(F, dtype=np.float64)
        self._chk(self.L.pcaone_set_F(self.h, _vp(F)))

    def sampleCovariance(self):
        """X X^T (N x N) for the current flags: `data->G * data->G.transpose()` of Main.cpp:187 / Halko.cpp:323,
        FP64 GEMM panels on the device."""
        K = _f((self.cols(), self.cols()))
        self._chk(self.L.pcaone_sample_covariance(self.h, _vp(K)))
        return K

    def grm(self):
        """PCAngsd GRM step (Halko.cpp:320-326) after runEM on a FileBeagle source: (C, Dc)."""
        Cm, Dc = _f((self.cols(), self.cols())), np.zeros(self.cols())
        self._chk(self.L.pcaone_gl_grm(self.h, _vp(Cm), _vp(Dc)))
        return Cm, Dc

    def symSVD(self, A):
        """SVD of a symmetric matrix on the device (one-sided Jacobi): (U, S) with S descending."""
        A = np.asfortranarray(A, dtype=np.float64)
        n = A.shape[0]
        U, S = _f((n, n)), np.zeros(n)
        sw = C.c_int(0)
        self._chk(self.L.pcaone_sym_svd(self.h, _vp(A), n, _vp(U), _vp(S), C.byref(sw)))
        self.jacobi_sweeps = sw.value
        return U, S

    def exactPCA(self):
        """`--svd 3` (Main.cpp:180-217): K = G G^T / nsnps on the standardised genotypes, its eigen-decomposition,
        V = G^T U / sqrt(eval nsnps), flip_UV by the largest |U| entry. Sets U, S, V and returns the eigenvalues."""
        self.setFlags(False, True)
        N, M = self.cols(), self.rows()
        ncomp = min(self.nk, N, M)
        Uall, Sall = self.symSVD(self.sampleCovariance() / M)
        evals = np.maximum(0.0, Sall[:ncomp])
        U = np.asfortranarray(Uall[:, :ncomp])
        svals = np.sqrt(evals * M)
        V = self.xtTimes(U)
        V[:, svals > 0] /= svals[svals > 0]
        x = np.abs(U).argmax(axis=0)
        sgn = np.where(U[x, np.arange(ncomp)] < 0, -1.0, 1.0)
        self.U, self.S, self.V = np.asfortranarray(U * sgn), svals, np.asfortranarray(V * sgn)
        return evals

    def xtTimes(self, A, want_sqnorm=False):
        """X^T A (M x c) [+ per-SNP squared norms]: Selection.cpp:16-34."""
        A = np.asfortranarray(A, dtype=np.float64)
        out = _f((self.rows(), A.shape[1]))
        nrm = np.zeros(self.rows()) if want_sqnorm else None
        self._chk(self.L.pcaone_xt_times(self.h, _vp(A), A.shape[1], _vp(out), _vp(nrm)))
        return (out, nrm) if want_sqnorm else out

    def xTimes(self, B):
        """X B (N x c): Projection.cpp:236-241."""
        B = np.asfortranarray(B, dtype=np.float64)
        out = _f((self.cols(), B.shape[1]))
        self._chk(self.L.pcaone_x_times(self.h, _vp(B), B.shape[1], _vp(out)))
        return out

    def maskTimes(self, B):
        """C B (N x c) with C the missing-call indicator (the reference's data->C): Projection.cpp:159-179."""
        B = np.asfortranarray(B, dtype=np.float64)
        out = _f((self.cols(), B.shape[1]))
        self._chk(self.L.pcaone_mask_times(self.h, _vp(B), B.shape[1], _vp(out)))
        return out

    def setUSV(self, U, S, V):
        U = np.asfortranarray(U, dtype=np.float64)
        V = np.asfortranarray(V, dtype=np.float64)
        S = np.ascontiguousarray(S, dtype=np.float64)
        self._chk(self.L.pcaone_set_usv(self.h, _vp(U), _vp(S), _vp(V)))

    def omega(self):
        out = _f((self.cols(), self.size()))
        self._chk(self.L.pcaone_get_omega(self.h, _vp(out)))
        return out

    def timers(self, reset=False):
        t = _lib.Timers()
        self._chk(self.L.pcaone_get_timers(self.h, C.byref(t), int(reset)))
        return t

    def enable_timing(self, on=True):
        self._chk(self.L.pcaone_enable_timing(self.h, int(on)))

    def sync(self):
        self._chk(self.L.pcaone_sync(self.h))


class NormalRsvdOpData(RsvdOpData):
    """sSVD / "Halko" (src/Halko.cpp:99-153)."""
    svd = _lib.SVD_SSVD


class FancyRsvdOpData(RsvdOpData):
    """winSVD / "PCAone Alg2" (src/Halko.cpp:155-269)."""
    svd = _lib.SVD_WINSVD


class ArnoldiOpData(NormalRsvdOpData):
    """src/Arnoldi.hpp:6-34: the operator Spectra's SymEigsSolver iterates, y = G G^T x accumulated
    over the blocks of the plan (src/Arnoldi.cpp:18-46). Only the operator is mirrored — the IRAM
    driver itself (Spectra) is host code that stays in PCAone. Built on a context with l = 1."""

    def __init__(self, data, **kw):
        super().__init__(data, 1, 0, **kw)
        self.nops = 1

    def perform_op(self, x_in, y_out=None):
        x = np.ascontiguousarray(x_in, dtype=np.float64)
        if x.shape != (self.cols(),):
            raise ValueError("x must have nsamples entries")
        y = np.zeros(self.cols()) if y_out is None else y_out
        self._chk(self.L.pcaone_perform_op(self.h, _vp(x), _vp(y)))
        self.nops += 1
        return y


def run_pca_with_halko(data: FileBed, params: Param, **kw):
    """src/Halko.cpp:271-345 without the file writers: returns the op with U, S, V set
    (eigenvalues are S**2 / nsnps as written at :339)."""
    if params.svd == 0:
        raise RuntimeError("--svd 0: only ArnoldiOpData.perform_op is on the GPU; the IRAM driver (Spectra) is host code")
    cls = FancyRsvdOpData if params.svd == 2 else NormalRsvdOpData
    rsvd = cls(data, params.k, params.oversamples, **kw)
    if getattr(data, "P", None) is not None:      # Beagle input => PCAngsd EM (Cmd.cpp:227, Halko.cpp:290-311)
        rsvd.em_iters = rsvd.runEM()
    elif not params.emu:
        rsvd.setFlags(False, not params.ld)
        rsvd.computeUSV(params.maxp, params.tol)
    else:
        rsvd.em_iters = rsvd.runEM()
    return rsvd


def mev(X, Y, device=0):
    """mev(X, Y) (src/Utils.cpp:194-200) evaluated on the device."""
    raise NotImplementedError("use RsvdOpData / pcaone_mev through an op context")
