// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Driver that runs the UNMODIFIED reference host code — Param (src/Cmd.cpp), Data::prepare
// (src/Data.cpp:14-85), permute_plink (src/FilePlink.cpp:303), RsvdOpData::computeUSV / initOmg /
// computeU (src/Halko.cpp:15-97) — on top of integration/HalkoGpu.hpp, i.e. with computeGandH served by
// libpcaone_b200.so through the C-ABI. tests/test_gpu_dropin.py compares what it returns with the
// same command run through the reference's own CPU ops (oracle/ref_shim.cpp).
#define _DECLARE_TOOLBOX_HERE
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "Cmd.hpp"
#include "Common.hpp"
#include "Data.hpp"
#include "FilePlink.hpp"
#include "Halko.hpp"
#include "HalkoGpu.hpp"
#include "Utils.hpp"

namespace {
thread_local std::string g_err;
}

extern "C" {

const char* refgpu_last_error() { return g_err.c_str(); }

// cmdline: a PCAone command line (PLINK input). precision: PCAONE_PREC_*. on_device != 0: the epoch
// loop runs on the device (computeUSVonDevice) instead of the reference's computeUSV on the host.
// Outputs (may be NULL): U N x k, S k, V M x k (column-major), dims = {N, M, k}. V rows are in the
// order the run used (permuted when the reference permutes); perm_out (M ints) receives the
// permutation indices or is left untouched when there is none.
int refgpu_run(const char* cmdline, int precision, int on_device, double* U, double* S, double* V, long long* dims,
               int* perm_out) {
  try {
    std::vector<std::string> toks;
    std::istringstream is(cmdline);
    for (std::string t; is >> t;) toks.push_back(t);
    std::vector<char*> argv;
    for (auto& t : toks) argv.push_back(const_cast<char*>(t.c_str()));
    Param params((int)argv.size(), argv.data());
    if (cao.cao.is_open()) cao.cao.close();
    cao.cao.open(params.fileout + ".log");
    cao.is_screen = false;
    if (params.file_t != FileType::PLINK) throw std::runtime_error("refgpu_run: PLINK input only");
    PermMat perm;
    bool have_perm = false;
    if (params.perm && params.out_of_core) {   // Main.cpp:125-130
      perm = permute_plink(params.filein, params.fileout, params.buffer, params.bands);
      have_perm = true;
    }
    GpuFileBed data(params);
    if (have_perm) data.perm = perm;
    data.prepare();                            // Main.cpp:168 — the reference's own block plan
    RsvdOpData* op = params.svd_t == SvdType::PCAoneAlg2
                         ? (RsvdOpData*)new GpuFancyRsvdOpData(&data, params.k, params.oversamples, precision)
                         : (RsvdOpData*)new GpuNormalRsvdOpData(&data, params.k, params.oversamples, precision);
    op->setFlags(false, params.ld ? false : true);   // Halko.cpp:283-288
    if (on_device)
      static_cast<GpuRsvdOpData*>(op)->computeUSVonDevice(params.maxp, params.tol);
    else
      op->computeUSV(params.maxp, params.tol);        // reference code, Halko.cpp:46-97
    if (dims) {
      dims[0] = data.nsamples;
      dims[1] = data.nsnps;
      dims[2] = params.k;
    }
    if (U) std::memcpy(U, op->U.data(), sizeof(double) * op->U.size());
    if (S) std::memcpy(S, op->S.data(), sizeof(double) * op->S.size());
    if (V) std::memcpy(V, op->V.data(), sizeof(double) * op->V.size());
    if (perm_out && data.perm.indices().size() == (Eigen::Index)data.nsnps)
      for (Eigen::Index i = 0; i < (Eigen::Index)data.nsnps; ++i) perm_out[i] = data.perm.indices()(i);
    delete op;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  } catch (...) {
    g_err = "unknown exception";
    return 2;
  }
}

}  // extern "C"
