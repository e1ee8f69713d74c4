#include "bgen.hpp"

#include <zlib.h>

#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>

// libzstd is on the box as a shared object without headers: the two entry points the reader needs
extern "C" {
size_t ZSTD_decompress(void* dst, size_t dstCapacity, const void* src, size_t compressedSize);
unsigned ZSTD_isError(size_t code);
}

namespace pcaone_host {

namespace {

struct Cursor {  // little-endian reads from a byte buffer
  const uint8_t* p;
  const uint8_t* end;
  template <class T>
  T get() {
    if (p + sizeof(T) > end) cao.error("BGEN: truncated block");
    T v;
    std::memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  void skip(size_t n) {
    if (p + n > end) cao.error("BGEN: truncated block");
    p += n;
  }
};

template <class T>
T fget(std::ifstream& f) {
  T v;
  f.read(reinterpret_cast<char*>(&v), sizeof(T));
  if (!f) cao.error("BGEN: unexpected end of file");
  return v;
}

void fskip(std::ifstream& f, uint64 n) {
  f.seekg((std::streamoff)n, std::ios::cur);
  if (!f) cao.error("BGEN: unexpected end of file");
}

std::vector<uint8_t> decompress(const std::vector<uint8_t>& src, size_t out_len, uint32_t compression) {
  std::vector<uint8_t> out(out_len);
  if (compression == 1) {
    uLongf n = (uLongf)out_len;
    if (uncompress(out.data(), &n, src.data(), (uLong)src.size()) != Z_OK || n != out_len)
      cao.error("BGEN: zlib decompression of a variant block failed");
  } else if (compression == 2) {
    const size_t n = ZSTD_decompress(out.data(), out_len, src.data(), src.size());
    if (ZSTD_isError(n) || n != out_len) cao.error("BGEN: zstd decompression of a variant block failed");
  } else {
    cao.error("BGEN: unknown compression flag");
  }
  return out;
}

// dosage of the FIRST allele for every sample (float), missing -> flagged; probabilities of an unphased diploid
// biallelic variant: P(AA), P(AB) stored, P(BB) implied (layout 2) or stored third (layout 1)
void first_allele_dosage(const std::vector<uint8_t>& raw, uint32_t layout, uint64 N, std::vector<float>& dose,
                         std::vector<uint8_t>& missing) {
  dose.assign(N, 0.f);
  missing.assign(N, 0);
  Cursor c{raw.data(), raw.data() + raw.size()};
  if (layout == 1) {
    const float factor = 1.0f / 32768.0f;
    for (uint64 i = 0; i < N; ++i) {
      const uint32_t hom = c.get<uint16_t>(), het = c.get<uint16_t>(), alt = c.get<uint16_t>();
      dose[i] = (float)(hom * 2u + het) * factor;
      if (hom == 0 && het == 0 && alt == 0) missing[i] = 1;
    }
    return;
  }
  if (c.get<uint32_t>() != N) cao.error("BGEN: a variant block has a different number of samples");
  if (c.get<uint16_t>() != 2) cao.error("BGEN: only biallelic variants give an allele dosage");
  const uint8_t pmin = c.get<uint8_t>(), pmax = c.get<uint8_t>();
  if (pmin != 2 || pmax != 2) cao.error("BGEN: only diploid samples are supported");
  const uint8_t* ploidy = c.p;
  c.skip(N);
  if (c.get<uint8_t>() != 0) cao.error("BGEN: phased probabilities are not supported");
  const uint32_t B = c.get<uint8_t>();
  if (B < 1 || B > 32) cao.error("BGEN: bits per probability out of range");
  const uint64 maxval = (1ull << B) - 1;
  const float factor = 1.0f / (float)maxval;
  const uint64 need_bits = 2ull * B * N;
  if ((uint64)(c.end - c.p) * 8 < need_bits) cao.error("BGEN: truncated probability data");
  uint64 bit = 0;
  auto take = [&](void) -> uint64 {  // B bits, least significant first
    uint64 v = 0;
    const uint64 byte = bit >> 3, sh = bit & 7;
    for (uint32_t k = 0; k < 5 && c.p + byte + k < c.end; ++k) v |= (uint64)c.p[byte + k] << (8 * k);
    bit += B;
    return (v >> sh) & maxval;
  };
  for (uint64 i = 0; i < N; ++i) {
    const uint64 hom = take(), het = take();
    dose[i] = (float)(hom * 2 + het) * factor;
    if (ploidy[i] & 0x80) missing[i] = 1;
  }
}

// which allele is the minor one: the sampled running frequency of the first allele with a 5-sigma early stop
// (the rule of the reference's bgen library; missing samples enter with the dosage their stored bits give)
bool first_allele_is_major(const std::vector<float>& dose) {
  const uint32_t N = (uint32_t)dose.size(), batch = 100;
  const uint32_t increment = std::max<uint32_t>(N / batch, 1u);
  double total = 0, freq = 0;
  for (uint32_t s = 0; s < increment; ++s) {
    for (uint32_t n = s; n < N; n += increment) total += dose[n];
    const double checked = (double)batch * (s + 1);
    freq = total / (checked * 2);
    const double delta = 5.0 * std::sqrt(freq * (1 - freq) / checked);
    if (!((freq - delta < 0.5) && (freq + delta > 0.5))) break;
  }
  return freq > 0.5;
}

}  // namespace

FileBgen::FileBgen(const Param& p) : Data(p) {
  cao.warn("BGEN support is very limited. Please convert BGEN to PGEN instead!");
  cao.print(tick.date(), "start parsing BGEN format");
  tick.clock();
  std::ifstream f(params.filein, std::ios::binary);
  if (!f.is_open()) cao.error("can not open " + params.filein);
  const uint32_t offset = fget<uint32_t>(f);
  const uint32_t header_len = fget<uint32_t>(f);
  nvariants_file = fget<uint32_t>(f);
  nsamples = fget<uint32_t>(f);
  char magic[4];
  f.read(magic, 4);
  if (std::memcmp(magic, "bgen", 4) != 0 && std::memcmp(magic, "\0\0\0\0", 4) != 0) cao.error("not a BGEN file: " + params.filein);
  if (header_len < 20) cao.error("BGEN: bad header length");
  fskip(f, header_len - 20);
  const uint32_t flags = fget<uint32_t>(f);
  const uint32_t compression = flags & 3u, layout = (flags >> 2) & 0xFu;
  if (layout != 1 && layout != 2) cao.error("BGEN: unsupported layout");
  cao.print(tick.date(), "N(#samples) =", nsamples, ", M(#SNPs) =", nvariants_file);
  cao.print(tick.date(), "the layout is", layout, ", compressed by", compression == 2 ? "zstd" : compression == 1 ? "zlib" : "none");
  f.seekg((std::streamoff)offset + 4, std::ios::beg);
  // ---- every variant: identifiers skipped, probabilities -> minor-allele dosage, af filter (FileBgen.cpp:22-45)
  std::vector<float> dose;
  std::vector<uint8_t> missing, comp, raw;
  dosages.reserve((size_t)nvariants_file * nsamples);
  uint64 kept = 0;
  for (uint64 j = 0; j < nvariants_file; ++j) {
    if (layout == 1 && fget<uint32_t>(f) != nsamples) cao.error("BGEN: a variant block has a different number of samples");
    fskip(f, fget<uint16_t>(f));                           // variant id
    fskip(f, fget<uint16_t>(f));                           // rsid
    fskip(f, fget<uint16_t>(f));                           // chromosome
    fskip(f, 4);                                           // position
    const uint32_t nalleles = layout == 2 ? fget<uint16_t>(f) : 2;
    for (uint32_t a = 0; a < nalleles; ++a) fskip(f, fget<uint32_t>(f));
    if (nalleles != 2) cao.error("BGEN: only biallelic variants give an allele dosage");
    if (layout == 2) {
      const uint32_t C = fget<uint32_t>(f);
      if (compression) {
        const uint32_t D = fget<uint32_t>(f);
        comp.resize(C - 4);
        f.read(reinterpret_cast<char*>(comp.data()), comp.size());
        raw = decompress(comp, D, compression);
      } else {
        raw.resize(C);
        f.read(reinterpret_cast<char*>(raw.data()), raw.size());
      }
    } else {
      if (compression) {
        comp.resize(fget<uint32_t>(f));
        f.read(reinterpret_cast<char*>(comp.data()), comp.size());
        raw = decompress(comp, 6 * nsamples, 1);
      } else {
        raw.resize(6 * nsamples);
        f.read(reinterpret_cast<char*>(raw.data()), raw.size());
      }
    }
    if (!f) cao.error("BGEN: unexpected end of file");
    first_allele_dosage(raw, layout, nsamples, dose, missing);
    if (first_allele_is_major(dose))
      for (auto& d : dose) d = 2.0f - d;
    double gs = 0;
    uint64 gc = 0;
    for (uint64 i = 0; i < nsamples; ++i) {
      if (missing[i]) {
        dose[i] = std::numeric_limits<float>::quiet_NaN();
      } else {
        gs += dose[i] / 2.0;
        ++gc;
      }
    }
    const double af = gc ? gs / (double)gc : 0.0;
    if (!(af > params.maf)) continue;  // FileBgen.cpp:42-45
    dosages.insert(dosages.end(), dose.begin(), dose.end());
    ++kept;
  }
  if (kept == 0) cao.error("the number of SNPs after filtering is 0!");
  cao.print(tick.date(), "number of SNPs after filtering by MAF >", params.maf, ":", kept);
  nsnps = kept;
  readtime += tick.reltime();
  // in-core winSVD shuffles the SNP order (Halko.cpp:183-186): fixed here, applied to the rows before the upload
  if (p.perm && p.svd_t == SvdType::PCAoneAlg2) {
    perm.resize(nsnps);
    pcaone_shuffle_indices(nsnps, perm.data());
    std::vector<float> shuffled(dosages.size());
    for (uint64 l = 0; l < nsnps; ++l)
      std::copy(dosages.begin() + (size_t)perm[l] * nsamples, dosages.begin() + (size_t)(perm[l] + 1) * nsamples,
                shuffled.begin() + (size_t)l * nsamples);
    dosages.swap(shuffled);
  }
}

void FileBgen::read_all() {
  check(pcaone_upload_dosage(ctx, dosages.data(), nsnps, 0));
  check(pcaone_allele_freq(ctx));
  F.resize(nsnps);
  check(pcaone_get_F(ctx, F.data()));
  uint64 nmiss = 0;
  check(pcaone_missing_count(ctx, &nmiss));
  p_miss = (double)nmiss / ((double)nsnps * (double)nsamples);
}

void FileBgen::read_block_initial(uint64 start_idx, uint64 stop_idx, bool standardize) {
  const uint64 B = stop_idx - start_idx + 1;
  if (G.rows() != nsamples || G.cols() != B) G.resize(nsamples, B);
  check(pcaone_decode_block(ctx, start_idx, stop_idx, standardize ? 1 : 0, 0, G.data()));
}

}  // namespace pcaone_host
