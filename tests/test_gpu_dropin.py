"""-m gpu: the compiled drop-in. integration/HalkoGpu.hpp (GpuFileBed : Data, Gpu{Normal,Fancy}RsvdOpData :
RsvdOpData) is built against the UNMODIFIED reference sources (oracle/Makefile, target refgpu) and the
reference's own host code — Param, Data::prepare, permute_plink, RsvdOpData::initOmg / computeUSV /
computeU (src/Halko.cpp:15-97: Householder QR x2, fullPivHouseholderQr solve, JacobiSVD, the MEV stopping
rule) — runs on G / H served by libpcaone_b200.so through the C-ABI. Outputs must equal the same command
line run through the reference's CPU ops (oracle/ref_shim.cpp)."""
import numpy as np
import pytest

from conftest import assert_usv_close
from oracle import ref
from pcaone_b200 import _lib, synth

pytestmark = [pytest.mark.gpu, pytest.mark.ref]


@pytest.fixture(scope="module")
def bed(tmp_path_factory):
    d = tmp_path_factory.mktemp("dropin")
    prefix = str(d / "g")
    synth.write_bed(prefix, 611, 7013, k_pop=6, seed=31)
    return prefix, str(d)


CASES = [
    # (extra flags, precision, epoch loop on the device?)
    ("-d 1", _lib.PREC_FP64, False),                    # sSVD in-core, reference computeUSV on the host
    ("-d 2 -w 8", _lib.PREC_INT8X3, False),             # winSVD in-core incl. permute_matrix (Halko.cpp:183-186)
    ("-d 2 -w 8 -m 0.001", _lib.PREC_INT8X3, False),    # winSVD out-of-core incl. permute_plink + the block plan
    ("-d 1 -m 0.001", _lib.PREC_FP64, False),           # sSVD out-of-core
    ("-d 2 -w 8 -m 0.001", _lib.PREC_INT8X3, True),     # same, whole epoch loop on the device
]


@pytest.mark.parametrize("flags,prec,on_device", CASES)
def test_reference_host_code_on_the_c_abi(bed, flags, prec, on_device):
    if not (ref.available() and ref.gpu_available()):
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    prefix, out = bed
    cmd = f"PCAone -b {prefix} -k 5 {flags} --maxp 8 -n 4"
    r = ref.Ref(cmd + f" -o {out}/cpu", threads=4)
    r.new_op()
    Ur, Sr, Vr = r.compute_usv(8, 1e-4)
    ep = r.last_epochs()
    r.close()
    U, S, V, _ = ref.gpu_run(cmd + f" -o {out}/gpu", prec, on_device)
    tol = 1e-9 if prec == _lib.PREC_FP64 else 1e-8
    assert_usv_close(U, S, V, Ur, Sr, Vr, eig_rtol=tol, min_corr=1 - 1e-8)
    assert ep >= 2
