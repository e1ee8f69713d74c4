// pcaone_b200 host — command line of the GPU front-end.
// Mirrors the reference's `Param` (/root/reference/src/Cmd.hpp:16-98, Cmd.cpp:14-238): same
// flag spellings, defaults and derived fields for everything the randomized-SVD path reads.
// Flags of subsystems outside the hot path (IRAM, projection, selection, inbreeding,
// clumping) are recognised so that a reference command line parses, and rejected with a
// clear message instead of being silently ignored.
#pragma once
#include <sstream>
#include <string>

#include "common.hpp"

namespace pcaone_host {

enum class FileType { PLINK, CSV, BEAGLE, BINARY, BGEN, PGEN, NONE };
enum class SvdType { IRAM, PCAoneAlg1, PCAoneAlg2, FULL };

constexpr int SCALE_STANDARDIZE_GENETIC = -9;

class Param {
 public:
  Param(int argc, char** argv);

  FileType file_t = FileType::NONE;
  SvdType svd_t = SvdType::PCAoneAlg2;
  std::string fileU, fileS, fileE, fileV;
  std::string filein;
  std::string fileout = "pcaone";
  double memory = 0;  // -m, GB; 0 = in-core
  uint nsamples = 0, nsnps = 0;
  uint k = 10;
  uint maxp = 20;
  uint threads = 12;  // accepted for compatibility; host threads are not on the GPU path
  uint bands = 64;
  bool genetic = true;
  bool dopca = true;
  int project = 0;  // --project 1 | 2 (Projection.cpp:188-246)
  bool perm = false;
  uint maxiter = 100;
  double tolem = 1e-5;
  double tolmaf = 1e-6;
  double maf = 0.0;
  uint oversamples = 10;
  double tol = 1e-4;
  uint buffer = 2;
  uint rand = 1;
  bool print_r2 = false;
  std::string filebim;
  int ld_stats = 0;
  double ld_r2 = 0;
  bool ld = false;
  uint ld_bp = 1000000;
  std::string clump;  // comma separated assoc-like files (Cmd.hpp:62-68)
  std::string assoc_colnames = "CHR,BP,P";
  double clump_p1 = 0.0001;
  double clump_p2 = 0.01;
  double clump_r2 = 0.5;
  uint clump_bp = 250000;
  uint verbose = 1;
  int scale = SCALE_STANDARDIZE_GENETIC;
  bool printv = false;
  bool missme = false;
  bool noshuffle = false;
  bool emu = false;
  bool pcangsd = false;
  bool mev = true;
  bool out_of_core = false;
  int ploidy = 2;
  int seed = 112;
  bool center = true;

  // GPU-side additions (not in the reference)
  int device = 0;            // --device: CUDA ordinal of the first GPU
  int gpus = 1;              // --gpus: SNP-shard the job over this many GPUs of the box (NCCL)
  std::string precision = "int8x3";  // --precision fp64|int8x2|int8x3|int8x4
  int precision_code() const;

  std::ostringstream ss;  // "Options in effect" banner, like Cmd.cpp:126-129
};

}  // namespace pcaone_host
