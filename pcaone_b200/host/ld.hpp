// pcaone_b200 host — LD front-end: --print-r2 and --ld-r2 pruning (ld_prune_big, LD.cpp:240-268). Window planning and the .ld.gz text
// format follow /root/reference/src/LD.cpp:79-103 (get_snp_pos_bim), :154-168
// (divide_pos_by_window) and :450-473 (ld_r2_big); the correlations themselves are one banded
// tile Gram on the device (pcaone_ld_r2).
#pragma once
#include <string>
#include <vector>

#include "data.hpp"

namespace pcaone_host {

struct SNPld {
  std::vector<int> pos;      // base-pair position of every SNP
  std::vector<int> end_pos;  // index of the last SNP of every chromosome
  std::vector<std::string> chr;
  std::vector<int> ws, we;   // lead SNP of each window / number of SNPs in it (lead included)
  std::vector<double> af;    // 7th column of a .mbim, if present (LD.cpp:90): MAF rule of the pruning
};

void get_snp_pos_bim(SNPld& snp, const std::string& filebim);
void divide_pos_by_window(SNPld& snp, int ld_window_bp);
std::vector<std::string> read_variant_labels(const std::string& filebim);  // "CHR\tBP\tSNP"
void run_ld_stuff(Data* data, const Param& params);

}  // namespace pcaone_host
