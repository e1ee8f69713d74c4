// pcaone_b200 host — the command-line front-end (`PCAone-b200`). Same flow as
// /root/reference/src/Main.cpp:49-224 for the paths SURVEY §8 puts in scope: PLINK bed input,
// --svd 1/2, in-core or -m out-of-core, --emu, --ld, --print-r2; everything numerical is a call
// into libpcaone_b200.so. There is no CPU fallback: without a CUDA device the first call fails.
#include <thread>

#include "beagle.hpp"
#include "bgen.hpp"
#include "csv.hpp"
#include "halko.hpp"
#include "ld.hpp"

namespace pcaone_host {
Logger cao;
Timer tick;
}  // namespace pcaone_host

using namespace pcaone_host;

static int bye() {
  cao.print(tick.date(), "total elapsed wall time:", tick.abstime(), " seconds");
  cao.print(tick.date(), "have a nice day. bye!");
  return 0;
}

int main(int argc, char* argv[]) {
  Param params(argc, argv);
  cao.file.open(params.fileout + ".log");
  if (params.verbose > 0) cao.is_screen = true;
  cao.print(params.ss.str());
  cao.print(tick.date(), "program started");
  try {
    if (pcaone_device_count() == 0) cao.error("no CUDA device is visible: pcaone_b200 has no CPU fallback");
    // LD from a bed (Main.cpp:78-97): windows from the .bim, r2 on the device
    if (params.print_r2 || params.ld_r2 > 0 || !params.clump.empty()) {
      params.memory = 0, params.out_of_core = false;  // Main.cpp:84
      params.perm = false;
      if (params.file_t == FileType::BINARY) {  // Main.cpp:149-150: -B file.residuals
        FileBin data(params);
        run_ld_stuff(&data, params);
        return bye();
      }
      FileBed data(params);
      run_ld_stuff(&data, params);
      return bye();
    }
    if (params.project > 0) {  // Main.cpp:95-107
      params.perm = false;
      FileBed data(params);
      data.prepare();
      run_projection(&data, params);
      return bye();
    }
    if (params.file_t == FileType::BINARY) cao.error("-B (binary residuals) is an LD input: give --print-r2, --ld-r2 or --clump");
    if (params.file_t == FileType::BEAGLE) {  // Main.cpp:117-120 + Halko.cpp:290-311 (PCAngsd EM)
      FileBeagle data(params);
      data.tolmaf = params.tolmaf;
      data.prepare();
      run_pca_with_halko(&data, params);
      if (params.pcangsd) run_pcangsd_grm(&data, params, data.samples);
      cao.print(tick.date(), "total elapsed reading time: ", data.readtime, " seconds");
      return bye();
    }
    if (params.file_t == FileType::CSV) {  // Main.cpp:160-161: a dense matrix, in core
      FileCsv data(params);
      data.prepare();
      run_pca_with_halko(&data, params);
      cao.print(tick.date(), "total elapsed reading time: ", data.readtime, " seconds");
      return bye();
    }
    if (params.file_t == FileType::BGEN) {  // Main.cpp:113-116: dosages, in core
      FileBgen data(params);
      data.prepare();
      run_pca_with_halko(&data, params);
      cao.print(tick.date(), "total elapsed reading time: ", data.readtime, " seconds");
      return bye();
    }
    if (params.svd_t == SvdType::FULL) {
      if (params.gpus > 1 || params.out_of_core) cao.error("--svd 3 runs in core on one GPU");
      params.perm = false;
      FileBed data(params);
      data.prepare();
      run_pca_full(&data, params);
    } else if (params.gpus > 1) {
      run_pca_sharded(params);
    } else {
      FileBed data(params);
      if (!data.perm.empty() && params.out_of_core)
        cao.print(tick.date(), "SNPs are permuted logically (the permute_plink map); no .perm.bed copy is written");
      data.prepare();
      run_pca_with_halko(&data, params);
      cao.print(tick.date(), "total elapsed reading time: ", data.readtime, " seconds");
    }
    make_plink2_eigenvec_file(params.k, params.fileout + ".eigvecs2", params.fileout + ".eigvecs",
                              params.filein + ".fam");
  } catch (const std::exception& e) {
    return 1;  // cao.error already logged the message (the reference aborts here)
  }
  return bye();
}
