"""-m gpu: EMU update passes on the int8 route (pcaone_b200/csrc/emu_fix.cuh). On an update pass every missing
call holds clamp(U S V^T) (FilePlink.cpp:246-259); the int8 route computes the mean-imputed block on the tensor
cores and adds the missing calls' terms in FP64. Checked here:

  * one pass against the numpy restatement of read_block_update: G = X^T Omega and H = X G~ to the 23-bit
    operand rounding of the int8 route (1000x below what leaving the fill out would cost);
  * whole EM runs against the FP64 DMMA route of the same library, sSVD and winSVD, resident / cached / streamed
    tiles, several register widths of the correction kernels (k <= 6 ... <= 56), ragged N and M.
The live-reference comparison of the same route is tests/test_gpu_scale.py::test_configs3_shape_emu_vs_reference.
"""
import os

import numpy as np
import pytest

from conftest import assert_usv_close
from oracle import pcaone_oracle as orc
from pcaone_b200 import _lib, halko, synth

pytestmark = pytest.mark.gpu


def _bed(N, M, seed, miss, k_pop=7):
    return np.concatenate([synth.pack_codes(c) for _, c in
                           synth.balding_nichols_codes(N, M, k_pop=k_pop, seed=seed, miss=miss)])


def X0h(od, M, standardize):
    return od.block(0, M - 1, standardize)


def _mk(packed, N, env=None, **kw):
    old = {}
    for key, val in (env or {}).items():
        old[key] = os.environ.get(key)
        os.environ[key] = val
    try:
        p = halko.Param(**kw)
        d = halko.FileBed(p, packed=packed, nsamples=N)
        d.prepare()
        cls = halko.FancyRsvdOpData if p.svd == 2 else halko.NormalRsvdOpData
        return cls(d, p.k, p.oversamples), d, p
    finally:
        for key, val in old.items():
            if val is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = val


@pytest.mark.parametrize("standardize", [False, True])
@pytest.mark.parametrize("N,M,k", [(1003, 9001, 5), (300, 2000, 20)])
def test_update_pass_vs_numpy_restatement(N, M, k, standardize):
    packed = _bed(N, M, 31, 0.08)
    op, d, p = _mk(packed, N, k=k, svd=1, emu=True, precision=_lib.PREC_INT8X3)
    od = orc.OracleData(packed, N)
    assert np.array_equal(op.F(), od.F)
    # a fill that clips on both sides: rank-k factors scaled so that |U S V^T| reaches past [-F, 1 - F]
    rng = np.random.default_rng(5)
    U = np.linalg.qr(rng.standard_normal((N, k)))[0]
    V = np.linalg.qr(rng.standard_normal((M, k)))[0]
    S = np.sqrt(N * M) * 0.35 / np.sqrt(np.arange(1, k + 1))
    fill = (U * S) @ V.T
    miss = (od.codes == 1).T
    clipped = ((fill < -od.F[None, :]) | (fill > 1 - od.F[None, :])) & miss
    assert 0.02 < clipped.sum() / miss.sum() < 0.9, "the case must exercise both the clamp and the plain fill"
    op.setUSV(U, S, V)
    op.setFlags(True, standardize)
    G, H = op.computeGandH(0)
    tm = op.timers()
    assert int(tm.tc_emu_ranges) >= 1 and int(tm.fp64_ranges) == 0
    X = od.block(0, M - 1, standardize, usv=(U, S, V), emu=True)      # N x M, read_block_update
    Gr = X.T @ op.Omg
    assert np.abs(G - Gr).max() <= 3e-6 * np.abs(Gr).max()             # Omega rounded to 23 bits per column scale
    Hr = X @ G
    # H = X G~ in the G~ the pass returns; with missing calls the mask operand D = (f - 1) W~ of the mean-imputed
    # part is itself rounded to 23 bits, so 2^-24 of the column scale, not FP64 accuracy
    eh = np.abs(H - Hr).max() / np.abs(Hr).max()
    print(f"N={N} M={M} k={k} standardize={standardize}: G err {np.abs(G - Gr).max() / np.abs(Gr).max():.2e}, H err {eh:.2e}")
    assert eh <= 2e-6
    assert np.abs(X0h(od, M, standardize) @ G - Hr).max() > 1e-3 * np.abs(Hr).max()
    # the fill is what moves the products: the same pass without it is far away
    X0 = od.block(0, M - 1, standardize)
    assert np.abs(X0.T @ op.Omg - Gr).max() > 1e-3 * np.abs(Gr).max()
    op.close()


CASES = [
    # svd, bands, k, N, M, memory (0 = resident), env
    (1, 64, 5, 1003, 9001, 0.0, None),
    (2, 16, 5, 1003, 9001, 0.0, None),
    (2, 16, 5, 1003, 9001, 0.004, None),                              # streamed once, then the HBM tile cache
    (2, 16, 5, 1003, 9001, 0.004, {"PCAONE_TILE_CACHE": "0"}),       # streamed every pass (per-buffer tiles)
    (1, 64, 13, 600, 4000, 0.0, None),                                # KR = 16, l = 26: two column tiles of 16
    (1, 64, 20, 640, 5000, 0.0, None),                                # KR = 24
    (2, 8, 36, 520, 4100, 0.0, None),                                 # KR = 56, l = 72: three column tiles
]


@pytest.mark.parametrize("svd,bands,k,N,M,memory,env", CASES)
def test_em_run_int8_route_equals_fp64_route(svd, bands, k, N, M, memory, env):
    packed = _bed(N, M, 32 + k, 0.06, k_pop=min(k, 8) + 2)
    out = {}
    for name, prec in (("int8", _lib.PREC_INT8X3), ("fp64", _lib.PREC_FP64)):
        op, d, p = _mk(packed, N, env=env if name == "int8" else None, k=k, svd=svd, bands=bands, emu=True, maxiter=3,
                       maxp=6, tol=0.0, no_shuffle=True, memory=memory, precision=prec)
        it = op.runEM()
        tm = op.timers()
        out[name] = (op.U, op.S, op.V, it, int(tm.tc_emu_ranges), int(tm.fp64_ranges), int(tm.tc_ranges))
        op.close()
    Ui, Si, Vi, iti, emu_i, fp_i, tc_i = out["int8"]
    Uf, Sf, Vf, itf, emu_f, fp_f, tc_f = out["fp64"]
    assert iti == itf
    assert emu_i > 0 and fp_i == 0, "every product of the int8 run is on the tensor-core route"
    assert emu_f == 0 and tc_f == 0
    assert_usv_close(Ui, Si, Vi, Uf, Sf, Vf, eig_rtol=1e-7, min_corr=1 - 1e-7)


def test_switch_restores_the_fp64_kernels():
    N, M, k = 400, 3000, 4
    packed = _bed(N, M, 40, 0.05)
    op, d, p = _mk(packed, N, env={"PCAONE_EMU_TC": "0"}, k=k, svd=1, emu=True, maxiter=2, maxp=4, tol=0.0,
                   precision=_lib.PREC_INT8X3)
    op.runEM()
    tm = op.timers()
    assert int(tm.tc_emu_ranges) == 0 and int(tm.fp64_ranges) > 0 and int(tm.tc_ranges) > 0
    op.close()
