"""-m gpu: the HBM cache of streamed tiles (pcaone_b200/csrc/sources.cu cache_plan). An out-of-core
source is read from the host on the first pass only when its re-tiled operands fit in HBM; the cache
is one tiling of the whole SNP axis, so blocks that start or end inside a 64-SNP k-block / 128-row
tile share chunks with their neighbours. Checked: cached == streamed every pass == resident == the
numpy oracle, partial caching, invalidation when the host hands over new data."""
import os

import numpy as np
import pytest

from conftest import assert_usv_close
from oracle import pcaone_oracle as orc
from pcaone_b200 import _lib, halko, synth

pytestmark = pytest.mark.gpu


def _bed(N, M, seed, miss=0.0):
    return np.concatenate([synth.pack_codes(c) for _, c in
                           synth.balding_nichols_codes(N, M, k_pop=7, seed=seed, miss=miss)])


def _run(packed, N, svd, *, memory, bands=16, k=5, maxp=7, env=None, keep=False):
    old = {}
    for key, val in (env or {}).items():
        old[key] = os.environ.get(key)
        os.environ[key] = val
    try:
        p = halko.Param(k=k, svd=svd, bands=bands, maxp=maxp, tol=0.0, no_shuffle=True, memory=memory,
                        precision=_lib.PREC_INT8X3)
        d = halko.FileBed(p, packed=packed, nsamples=N)
        d.prepare()
        cls = halko.FancyRsvdOpData if svd == 2 else halko.NormalRsvdOpData
        op = cls(d, p.k, p.oversamples)
        op.setFlags(False, True)
        op.computeUSV(maxp, 0.0)
        out = dict(U=op.U, S=op.S, V=op.V, F=op.F(), hits=int(op.timers().cache_hits), nblocks=d.nblocks, op=op, d=d)
        if not keep:
            op.close()
        return out
    finally:
        for key, val in old.items():
            if val is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = val


@pytest.mark.parametrize("svd", [1, 2])
@pytest.mark.parametrize("miss", [0.0, 0.03])
def test_cached_equals_streamed_equals_resident(svd, miss):
    N, M = 1003, 9001   # ragged: blocks start / end inside k-blocks and row tiles, padded last byte
    packed = _bed(N, M, 21, miss)
    resident = _run(packed, N, svd, memory=0.0)
    streamed = _run(packed, N, svd, memory=0.004, env={"PCAONE_TILE_CACHE": "0"})
    cached = _run(packed, N, svd, memory=0.004)
    assert streamed["hits"] == 0 and cached["nblocks"] > 8
    assert cached["hits"] >= (7 - 1) * cached["nblocks"]   # every block of every pass after the first
    assert np.array_equal(cached["F"], streamed["F"]) and np.array_equal(cached["F"], resident["F"])
    # cached blocks with no Omega update between them run as ONE range. Without missing calls that only
    # regroups exact integer sums (1e-11); with missing calls the H-pass mask operand D = (f - 1) W~ is
    # rounded to 23 bits against the RANGE's column scale, so regrouping moves the imputed
    # zeros by 2^-24 of that scale: documented bound 1e-9 on the eigenvalues
    tol = 1e-11 if miss == 0.0 else 5e-9
    assert_usv_close(cached["U"], cached["S"], cached["V"], streamed["U"], streamed["S"], streamed["V"], eig_rtol=tol,
                     min_corr=1 - 10 * tol)
    # against the numpy restatement of the reference on the same plan
    od = orc.OracleData(packed, N)
    p = halko.Param(k=5, svd=svd, bands=16, maxp=7, tol=0.0, no_shuffle=True, memory=0.004)
    d = halko.FileBed(p, packed=packed, nsamples=N)
    d.prepare()
    Om = np.zeros((N, p.l), order="F")
    _lib.load().pcaone_init_omega(N, p.l, p.seed, 1, Om.ctypes.data)
    blocks = list(zip(d.start.astype(int).tolist(), d.stop.astype(int).tolist()))
    oo = orc.OracleRsvd(od, 5, winsvd=svd == 2, bands=16, omega=Om, windows=blocks, out_of_core=True,
                        band_factor=d.bandFactor)
    oo.set_flags(False, True)
    U, S, V = oo.compute_usv(7, 0.0)
    otol = 1e-9 if miss == 0.0 else 1e-8   # (missing calls: the mask operand is rounded to 23 bits)
    assert_usv_close(cached["U"], cached["S"], cached["V"], U, S, V, eig_rtol=otol, min_corr=1 - otol)


def test_partial_cache_and_invalidation():
    N, M = 1003, 9001
    packed = _bed(N, M, 22)
    full = _run(packed, N, 2, memory=0.004)
    # room for roughly a third of the tiles: the first blocks come from HBM, the rest keep streaming
    total_mb = 2.2 * packed.size / 2 ** 20
    part = _run(packed, N, 2, memory=0.004, env={"PCAONE_TILE_CACHE_MB": str(max(1, int(total_mb / 3)))}, keep=True)
    assert 0 < part["hits"] < full["hits"]
    assert_usv_close(part["U"], part["S"], part["V"], full["U"], full["S"], full["V"], eig_rtol=1e-11, min_corr=1 - 1e-10)
    # the host hands over a different bed behind the same plan: nothing of the old tiles may survive
    op = part["op"]
    other = np.ascontiguousarray(_bed(N, M, 23))
    op._chk(op.L.pcaone_set_host_source(op.h, other.ctypes.data, M))
    op.computeUSV(7, 0.0)
    fresh = _run(other, N, 2, memory=0.004)
    assert np.array_equal(op.F(), fresh["F"])
    assert_usv_close(op.U, op.S, op.V, fresh["U"], fresh["S"], fresh["V"], eig_rtol=1e-11, min_corr=1 - 1e-10)
    op.close()
