"""CPU: host-side logic of the product (block planning, permutation maps, Param derivations),
the C-ABI library's exports, and the loud failure without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden
from pcaone_b200 import _lib, halko, synth


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pcaone_b200.h")).read()
    declared = set(re.findall(r"\b(pcaone_[a-z_0-9A-Z]+)\s*\(", hdr))
    declared -= {"pcaone_allreduce_fn", "pcaone_read_block_fn"}
    L = _lib.load()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert L.pcaone_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib.load()
    cfg = _lib.Config(nsamples=10, nsnps=100, k=2, oversamples=2, svd=1, ploidy=2, scale=-9, world=1)
    h = C.c_void_p()
    assert L.pcaone_create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CPU fallback" in L.pcaone_last_error(None)
    data = halko.FileBed(halko.Param(k=2, svd=1), packed=np.zeros((100, 3), np.uint8), nsamples=10)
    data.prepare()
    with pytest.raises(RuntimeError):
        halko.NormalRsvdOpData(data, 2, 2)


def test_param_derivations():
    p = halko.Param(k=3)
    assert p.oversamples == 10 and p.l == 13 and p.perm and not p.out_of_core
    p = halko.Param(k=40, svd=1, memory=2.0)
    assert p.oversamples == 40 and p.l == 80 and not p.perm and p.out_of_core
    with pytest.raises(ValueError):
        halko.Param(bands=3)


def test_block_plan_matches_reference_golden():
    g = golden("ssvd_small")
    N, M, k = int(g["N"]), int(g["M"]), int(g["k"])
    w = golden("winsvd_ooc_small")
    bs, nb, bf, start, stop = halko.ooc_block_plan(N, M, k + 10, float(w["memory"]), True, int(w["bands"]))
    assert [bs, nb, bf] == list(w["plan"])
    assert np.array_equal(start, w["start"]) and np.array_equal(stop, w["stop"])
    s = golden("ssvd_ooc_small")
    bs, nb, bf, start, stop = halko.ooc_block_plan(N, M, k + 10, float(s["memory"]), False, 64)
    assert [bs, nb, bf] == list(s["plan"])
    assert np.array_equal(start, s["start"]) and np.array_equal(stop, s["stop"])
    with pytest.raises(RuntimeError):
        halko.ooc_block_plan(N, M, k + 10, 10.0, False, 64)


def test_permute_plink_indices_match_reference_golden():
    g = golden("ssvd_small")
    w = golden("winsvd_ooc_small")
    assert np.array_equal(halko.permute_plink_indices(int(g["M"]), int(g["N"]), int(w["bands"])), w["perm"])
    assert np.array_equal(halko.permute_plink_indices(203, 13, 8), golden("helpers")["plink_perm"])


def test_host_random_streams_match_reference_golden():
    h = golden("helpers")
    L = _lib.load()
    a = np.zeros((7, 3), order="F")
    L.pcaone_init_omega(7, 3, 9, 1, a.ctypes.data_as(C.c_void_p))
    assert np.array_equal(a, h["omega_7x3_seed9"])
    L.pcaone_init_omega(7, 3, 9, 0, a.ctypes.data_as(C.c_void_p))
    assert np.array_equal(a, h["omega_uniform"])
    idx = np.zeros(10, dtype=np.uint32)
    L.pcaone_shuffle_indices(10, idx.ctypes.data_as(C.c_void_p))
    assert np.array_equal(idx, h["shuffle10"])
    g = golden("ssvd_small")
    om = np.zeros((int(g["N"]), 13), order="F")
    L.pcaone_init_omega(int(g["N"]), 13, 112, 1, om.ctypes.data_as(C.c_void_p))
    assert np.array_equal(om, g["omega"])
    idx = np.zeros(int(g["M"]), dtype=np.uint32)
    L.pcaone_shuffle_indices(int(g["M"]), idx.ctypes.data_as(C.c_void_p))
    assert np.array_equal(idx, golden("winsvd_small")["perm"])


def test_synth_bed_roundtrip(tmp_path):
    prefix = str(tmp_path / "x")
    packed = synth.write_bed(prefix, 23, 40, k_pop=3, seed=4, miss=0.05)
    p2, n, m = synth.read_bed(prefix)
    assert (n, m) == (23, 40) and np.array_equal(packed, p2)
    codes = synth.unpack_codes(packed, 23)
    assert np.array_equal(synth.pack_codes(codes), packed)
    bad = open(prefix + ".bed", "r+b")
    bad.write(b"\x00")
    bad.close()
    with pytest.raises(ValueError):
        synth.read_bed(prefix)
    with pytest.raises(RuntimeError):
        halko.FileBed(halko.Param(filein=prefix, k=2, svd=1))


@pytest.mark.parametrize("bit_depth,compression,layout", [(8, 1, 2), (16, 1, 2), (5, 1, 2), (8, 2, 2), (8, 0, 2), (16, 1, 1)])
def test_bgen_reader_vs_reference_library(tmp_path, bit_depth, compression, layout):
    """The front-end's own BGEN reader (host/bgen.cpp: header, zlib / zstd blocks, layout 1 and 2 packing, minor-allele
    rule) against the dosages the reference's vendored bgen library returns for the same file, written with that
    library's writer. Bit for bit except the 8-bit fast path of the library, whose tail samples come from a table of
    7-decimal literals (2.4e-7)."""
    import subprocess
    from oracle import ref
    dump = os.path.join(ROOT, "pcaone_b200", "bin", "bgen_dump")
    if not ref.available() or not os.path.exists(dump):
        pytest.skip("needs oracle/_ref and pcaone_b200/bin/bgen_dump")
    rng = np.random.default_rng(11)
    N, M = 157, 260
    f = rng.uniform(0.05, 0.95, M)                      # both orientations of the minor allele
    g = rng.binomial(2, f[:, None], size=(M, N))
    P = np.zeros((M, N, 3))
    for q in range(3):
        P[:, :, q] = np.where(g == q, 0.85, 0.075)
    P = 0.7 * P + 0.3 * rng.dirichlet([8, 8, 8], size=(M, N))
    P /= P.sum(-1, keepdims=True)
    P[rng.random((M, N)) < 0.03] = np.nan
    P[5] = [1.0, 0.0, 0.0]                              # a monomorphic variant: af = 0 after the minor-allele swap, dropped
    path = str(tmp_path / "t.bgen")
    ref.write_bgen(path, P, bit_depth=bit_depth, compression=compression, layout=layout)
    want = ref.bgen_dosages(path, N, M)
    af = np.nanmean(want / 2.0, axis=1)
    want = want[af > 0]
    r = subprocess.run([dump, "--bgen", path, "-S", "-o", str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    raw = open(str(tmp_path / "o.dosages"), "rb").read()
    m, n = (int(x) for x in np.frombuffer(raw[:16], dtype=np.uint64))
    got = np.frombuffer(raw[16:], dtype=np.float32).reshape(m, n)
    assert (m, n) == want.shape and m == M - 1
    assert np.array_equal(np.isnan(got), np.isnan(want))
    tol = 3e-7 if (bit_depth == 8 and layout == 2) else 0.0
    assert np.nanmax(np.abs(got - want)) <= tol


@pytest.mark.parametrize("scale", [0, 1, 2])
def test_csv_reader_vs_reference(tmp_path, scale):
    """The front-end's CSV reader (host/csv.cpp: zstd stream, parsing, -C normalisation, per-feature standardisation)
    against the data->G the reference's FileCsv::read_all builds from the same file (FileCsv.cpp:10-62)."""
    import subprocess
    import pyarrow as pa
    from oracle import ref
    dump = os.path.join(ROOT, "pcaone_b200", "bin", "csv_dump")
    if not ref.available() or not os.path.exists(dump):
        pytest.skip("needs oracle/_ref and pcaone_b200/bin/csv_dump")
    rng = np.random.default_rng(5 + scale)
    N, M = 77, 310
    cnt = rng.poisson(rng.gamma(2.0, 2.0, size=(M, 1)) * np.exp(rng.normal(0, 0.5, size=(1, N))))
    cnt[7] = 3                                             # a constant feature: sd = 0, left centred only
    path = str(tmp_path / "c.csv.zst")
    text = "".join(",".join(str(int(x)) for x in row) + "\n" for row in cnt)
    open(path, "wb").write(pa.Codec("zstd").compress(text.encode(), asbytes=True))
    r = ref.Ref(f"PCAone --csv {path} -k 3 -d 1 -C {scale} -o {tmp_path}/r -n 2", threads=2)
    want = r.dataG()
    r.close()
    q = subprocess.run([dump, "--csv", path, "-C", str(scale), "-S", "-o", str(tmp_path / "o")], capture_output=True, text=True)
    assert q.returncode == 0, q.stderr[-500:]
    raw = open(str(tmp_path / "o.matrix"), "rb").read()
    n, m = (int(x) for x in np.frombuffer(raw[:16], dtype=np.uint64))
    got = np.frombuffer(raw[16:], dtype=np.float64).reshape(m, n).T
    assert (n, m) == (N, M) == want.shape
    assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
