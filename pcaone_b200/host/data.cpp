// pcaone_b200 host — Data / FileBed (see data.hpp). File layout facts follow
// /root/reference/src/FilePlink.hpp:19-27 (bpr = ceil(N/4), magic 6c 1b 01) and the writers
// follow Data.cpp:110-291; the arithmetic (allele frequency, decode, EMU fill, residuals) runs
// in the CUDA library.
#include "data.hpp"

#include <sys/mman.h>
#include <sys/stat.h>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <iomanip>
#include <numeric>

namespace pcaone_host {

uint64 count_lines(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in.is_open()) cao.error("can not open " + path);
  uint64 n = 0;
  char last = '\n';
  std::vector<char> buf(1 << 20);
  while (in) {
    in.read(buf.data(), buf.size());
    const std::streamsize got = in.gcount();
    for (std::streamsize i = 0; i < got; ++i) n += buf[i] == '\n';
    if (got > 0) last = buf[got - 1];
  }
  if (last != '\n') ++n;  // unterminated final line
  return n;
}

BlockPlan ooc_block_plan(uint64 N, uint64 M, uint l, double memory_gb, bool winsvd, uint bands) {
  // Data.cpp:42-84. 2^27 doubles per GB; 3Nl + 2Ml + 5M doubles are the reference's own
  // working set and are subtracted from the budget exactly as it does, so that the block
  // boundaries (which winSVD results depend on) are the reference's.
  BlockPlan bp;
  const double gb = 134217728.0;
  double m = (double)(3 * N * l + 2 * M * l + 5 * M) / gb;
  if (memory_gb > 1.1 * m)
    m = 0;
  else
    cao.warn("minimum RAM required is " + std::to_string(m) + " GB. trying to allocate more RAM.");
  bp.blocksize = (uint64)std::ceil(((m + memory_gb) * gb - 3.0 * N * l - 2.0 * M * l - 5.0 * M) / (double)N);
  bp.nblocks = (uint)((M + bp.blocksize - 1) / bp.blocksize);
  if (bp.nblocks == 1) cao.error("only one block exists. please remove -m / --memory option instead.");
  if (winsvd) {
    if (bp.nblocks < bands) {
      bp.blocksize = (M + bands - 1) / bands;
    } else {
      bp.bandFactor = (bp.nblocks + bands - 1) / bands;
      bp.blocksize = (M + (uint64)bands * bp.bandFactor - 1) / ((uint64)bands * bp.bandFactor);
    }
    bp.nblocks = (uint)((M + bp.blocksize - 1) / bp.blocksize);
  }
  bp.start.resize(bp.nblocks);
  bp.stop.resize(bp.nblocks);
  for (uint i = 0; i < bp.nblocks; ++i) {
    bp.start[i] = (uint64)i * bp.blocksize;
    bp.stop[i] = std::min(bp.start[i] + bp.blocksize - 1, M - 1);
  }
  return bp;
}

std::vector<uint32_t> permute_plink_indices(uint64 M, uint64 N, uint bands, uint gb) {
  // the index map of permute_plink (FilePlink.cpp:303-408): a read buffer of `two` SNPs (a
  // multiple of `bands`) is dealt round-robin into the bands; SNP `i*two + j*bands + b` lands at
  // `i*bufsize + bandidx[b] + j`. Only the map is needed: rows are gathered at read time.
  const uint64 bpr = (N + 3) >> 2;
  uint64 two = (uint64)std::floor(1073741824.0 * gb / (double)bpr);
  two = std::min(two, M);
  uint64 bufsize = two / bands;
  two = bufsize * bands;
  if (two == 0) cao.error("permute_plink: fewer SNPs than batches");
  const uint64 nblocks = (M + two - 1) / two;
  uint64 modr2 = M % two;
  const uint64 modr = M % bands;
  const uint64 bandsize = (M + bands - 1) / bands;
  std::vector<uint64> bandidx(bands);
  for (uint64 i = 0; i < bands; ++i)
    bandidx[i] = (modr == 0 || i < modr) ? i * bandsize : modr * bandsize + (bandsize - 1) * (i - modr);
  std::vector<uint32_t> indices(M, 0);
  const uint64 bufidx = bufsize;
  for (uint64 i = 0; i < nblocks; ++i) {
    if (i == nblocks - 1 && modr2 != 0) {
      const uint64 two2 = M - (nblocks - 1) * two;
      bufsize = (two2 + bands - 1) / bands;
      modr2 = two2 % bands;
    }
    for (uint64 b = 0; b < bands; ++b) {
      const uint64 base = i * bufidx + bandidx[b];
      for (uint64 j = 0; j + 1 < bufsize; ++j) indices[base + j] = (uint32_t)(i * two + j * bands + b);
      const uint64 jl = bufsize - 1;
      if (i != nblocks - 1 || b < modr2 || modr2 == 0) indices[base + jl] = (uint32_t)(i * two + jl * bands + b);
    }
  }
  return indices;
}

// ---------------------------------------------------------------------------- Data
Data::~Data() {
  if (ctx) pcaone_destroy(ctx);
}

void Data::check(int rc) const {
  if (rc) cao.error(std::string(pcaone_last_error(ctx)));
}

void Data::create_context() {
  pcaone_config cfg{};
  cfg.nsamples = nsamples;
  cfg.nsnps = nsnps_local;
  cfg.nsnps_total = nsnps;
  cfg.k = params.k;
  cfg.oversamples = params.oversamples;
  cfg.svd = params.svd_t == SvdType::PCAoneAlg1 ? PCAONE_SVD_SSVD : PCAONE_SVD_WINSVD;
  cfg.bands = params.bands;
  cfg.maxp = params.maxp;
  cfg.tol = params.tol;
  cfg.ploidy = params.ploidy;
  cfg.scale = params.scale;
  cfg.emu = params.emu ? 1 : 0;
  cfg.out_of_core = params.out_of_core ? 1 : 0;
  cfg.precision = params.precision_code();
  cfg.device = params.device + shard.rank;
  cfg.rank = shard.rank;
  cfg.world = shard.world;
  cfg.maxiter = params.maxiter;
  cfg.tolem = params.tolem;
  svd_code = cfg.svd;
  if (pcaone_create(&cfg, &ctx)) cao.error(std::string(pcaone_last_error(nullptr)));
  if (pcaone_precision(ctx) != cfg.precision)
    cao.warn("--precision int8x* holds (k + oversamples) x slices <= 256 columns; this run uses the FP64 tensor-core kernels");
}

static std::pair<uint64, uint64> shard_range(uint64 n, int rank, int world) {
  const uint64 base = n / world, rem = n % world;
  const uint64 s = rank * base + std::min<uint64>(rank, rem);
  return {s, s + base + ((uint64)rank < rem ? 1 : 0)};
}

void Data::prepare() {
  const bool winsvd = params.svd_t == SvdType::PCAoneAlg2;
  const uint l = params.k + params.oversamples;
  // ---- block / window plan on the WHOLE job (Data.cpp:42-84, Halko.cpp:180-194)
  std::vector<uint64> gs, ge;
  if (params.out_of_core) {
    BlockPlan bp = ooc_block_plan(nsamples, nsnps, l, params.memory, winsvd, params.bands);
    blocksize = bp.blocksize;
    nblocks = bp.nblocks;
    bandFactor = bp.bandFactor;
    gs = bp.start;
    ge = bp.stop;
    cao.print(tick.date(), "initial setting by -m/--memory: blocksize =", blocksize, ", nblocks =", nblocks,
              ", factor =", bandFactor);
  } else if (winsvd && shard.world > 1) {
    const uint64 bs = (nsnps + params.bands - 1) / params.bands;
    for (uint64 b = 0; b < params.bands; ++b) {
      const uint64 s = b * bs, e = std::min((b + 1) * bs, nsnps);
      if (s < e) {
        gs.push_back(s);
        ge.push_back(e - 1);
      }
    }
  }
  // ---- this context's rows: 1/world of every block (or one contiguous slice)
  shard.snps.clear();
  shard.start.clear();
  shard.stop.clear();
  if (gs.empty()) {
    auto r = shard_range(nsnps, shard.rank, shard.world);
    shard.snps.resize(r.second - r.first);
    std::iota(shard.snps.begin(), shard.snps.end(), r.first);
  } else {
    for (size_t b = 0; b < gs.size(); ++b) {
      auto r = shard_range(ge[b] - gs[b] + 1, shard.rank, shard.world);
      if (r.second == r.first) cao.error("a block has fewer SNPs than GPUs; use fewer --gpus");
      shard.start.push_back(shard.snps.size());
      for (uint64 j = r.first; j < r.second; ++j) shard.snps.push_back(gs[b] + j);
      shard.stop.push_back(shard.snps.size() - 1);
    }
  }
  nsnps_local = shard.snps.size();
  start = gs;
  stop = ge;
  create_context();
  if (!params.out_of_core) {
    read_all();
    if (!shard.start.empty())
      check(pcaone_set_blocks(ctx, shard.start.data(), shard.stop.data(), (uint32_t)shard.start.size(), bandFactor));
  } else {
    attach_stream_source();
    check(pcaone_set_blocks(ctx, shard.start.data(), shard.stop.data(), (uint32_t)shard.start.size(), bandFactor));
  }
}

static void write_row(std::ostream& os, const Mat2D& A, uint64 i) {
  for (uint64 j = 0; j < A.cols(); ++j) {
    if (j) os << '\t';
    os << A(i, j);
  }
  os << '\n';
}

void Data::write_eigs_files(const Mat1D& E, const Mat1D& S, const Mat2D& U, const Mat2D& V) {
  // Data.cpp:211-240; Eigen::IOFormat(6, DontAlignCols, "\t", "\n") == stream precision 6
  std::ofstream outs(params.fileout + ".sigvals"), oute(params.fileout + ".eigvals"),
      outu(params.fileout + ".eigvecs");
  outs << std::setprecision(6);
  oute << std::setprecision(6);
  outu << std::setprecision(6);
  if (outs.is_open()) {
    outs << '#' << U.rows() << ',' << V.rows() << '\n';
    for (uint64 i = 0; i < S.size(); ++i) outs << S(i) << '\n';
  }
  if (oute.is_open())
    for (uint64 i = 0; i < E.size(); ++i) oute << E(i) << '\n';
  if (outu.is_open())
    for (uint64 i = 0; i < U.rows(); ++i) write_row(outu, U, i);
  if (params.printv) {
    save_snps_in_mbim();
    std::ofstream outv(params.fileout + ".loadings");
    if (!outv.is_open()) cao.error("can not open " + params.fileout + ".loadings");
    outv << std::setprecision(6);
    if (!perm.empty() && V.rows() == nsnps) {
      std::vector<uint64> original_to_logical(nsnps);
      for (uint64 j = 0; j < nsnps; ++j) original_to_logical[perm[j]] = j;
      for (uint64 o = 0; o < nsnps; ++o) write_row(outv, V, original_to_logical[o]);
    } else {
      for (uint64 i = 0; i < V.rows(); ++i) write_row(outv, V, i);
    }
  }
  cao.print(tick.date(), "eigen vectors and values saved");
}

void Data::save_snps_in_mbim() {
  // Data.cpp:110-170: the .bim line of every SNP plus its allele frequency, in ORIGINAL order
  std::ifstream bim(params.filein + ".bim");
  if (!bim.is_open()) {
    cao.warn(params.filein + ".bim/.pvar not found; skipping mbim output");
    return;
  }
  if (F.size() != nsnps) return;  // a sharded rank other than the writer
  std::ofstream out(params.fileout + ".mbim");
  std::vector<uint64> original_to_logical;
  if (!perm.empty()) {
    original_to_logical.resize(nsnps);
    for (uint64 j = 0; j < nsnps; ++j) original_to_logical[perm[j]] = j;
  }
  std::string line;
  for (uint64 o = 0; o < nsnps && std::getline(bim, line); ++o)
    out << line << "\t" << F(perm.empty() ? o : original_to_logical[o]) << "\n";
  if (!perm.empty()) cao.print(tick.date(), "save matched sites in .mbim file and permutation mode is", params.perm);
}

void Data::write_residuals(const Mat1D& S, const Mat2D& U, const Mat2D& VT) {
  // Data.cpp:242-291: [uint32 M][uint32 N] then one float32 column per SNP at its ORIGINAL
  // position. Blocks are decoded (centred, not standardised) on the device; ld-stats 0 removes
  // U S V^T first. The reference streams the same blocks through read_block_initial.
  if (shard.world > 1) cao.error("--ld with --gpus > 1 is not supported yet");
  cao.print(tick.date(), params.ld_stats == 1 ? "ld-stats=1: calculate standardized genotype matrix!"
                                              : "ld-stats=0: calculate the ancestry adjusted LD matrix!");
  std::FILE* fp = std::fopen((params.fileout + ".residuals").c_str(), "wb");
  if (!fp) cao.error("can not open " + params.fileout + ".residuals");
  const uint32_t m32 = (uint32_t)nsnps, n32 = (uint32_t)nsamples;
  std::fwrite(&m32, 4, 1, fp);
  std::fwrite(&n32, 4, 1, fp);
  const uint64 bytes_per_snp = nsamples * 4, magic = 8;
  // the float32 rows come from the device (pcaone_residuals_block: decode, G -= U S V^T from the
  // context's U, S, V, column-centre, cast) — resident or streamed through the block plan's buffers
  (void)S;
  (void)U;
  (void)VT;
  const uint64 B = std::max<uint64>(1, std::min<uint64>(nsnps, (256ull << 20) / (nsamples * 4)));
  std::vector<float> fg(B * nsamples);
  for (uint64 s = 0; s < nsnps; s += B) {
    const uint64 e = std::min(nsnps, s + B) - 1;
    check(pcaone_residuals_block(ctx, s, e, params.ld_stats, fg.data()));
    for (uint64 ib = 0; ib <= e - s; ++ib) {
      const uint64 orig = perm.empty() ? s + ib : perm[s + ib];
      fseeko(fp, (off_t)(magic + orig * bytes_per_snp), SEEK_SET);
      std::fwrite(fg.data() + ib * nsamples, 4, nsamples, fp);
    }
  }
  std::fclose(fp);
  save_snps_in_mbim();
  cao.print(tick.date(), "the LD matrix and SNPs info are saved");
}

// ---------------------------------------------------------------------------- FileBed
FileBed::FileBed(const Param& p, int rank, int world) : Data(p) {
  shard.rank = rank;
  shard.world = world;
  cao.print(tick.date(), "start parsing PLINK format");
  nsamples = p.nsamples ? p.nsamples : count_lines(p.filein + ".fam");
  nsnps = p.nsnps ? p.nsnps : count_lines(p.filein + ".bim");
  cao.print(tick.date(), "N (# samples):", nsamples, ", M (# SNPs):", nsnps);
  bed_bytes_per_snp = (nsamples + 3) >> 2;
  fd = ::open((p.filein + ".bed").c_str(), O_RDONLY);
  if (fd < 0) cao.error("Cannot open bed file.");
  unsigned char hdr[3];
  if (::pread(fd, hdr, 3, 0) != 3 || hdr[0] != 0x6c || hdr[1] != 0x1b || hdr[2] != 0x01)
    cao.error("Incorrect magic number in plink bed file.");
  struct stat st;
  if (fstat(fd, &st) != 0 || (uint64)st.st_size < 3 + bed_bytes_per_snp * nsnps)
    cao.error("the bed file is shorter than 3 + ceil(N/4) * M bytes; check .fam / .bim");
  // winSVD shuffles the SNP order (Halko.cpp:183-186 in-core, permute_plink out-of-core). The
  // map is fixed here and rows are gathered at read time, so no .perm.bed copy is written.
  if (p.perm && p.svd_t == SvdType::PCAoneAlg2) {
    if (p.out_of_core) {
      perm = permute_plink_indices(nsnps, nsamples, p.bands, p.buffer);
    } else {
      perm.resize(nsnps);
      pcaone_shuffle_indices(nsnps, perm.data());  // the reference's unseeded std::shuffle stream
    }
  }
}

FileBed::~FileBed() {
  if (ctx) {
    pcaone_destroy(ctx);  // before the pinned buffer it may still be reading from
    ctx = nullptr;
  }
  if (pinned) pcaone_free_pinned(pinned);
  if (fd >= 0) ::close(fd);
}

// copy local rows [s, e] (logical order of this shard) into dst; contiguous runs of original
// rows become one pread
static void gather_rows(const std::vector<uint64>& snps, const std::vector<uint32_t>& perm,
                        uint64 s, uint64 e, uint64 bpr, uint8_t* dst, int fd) {
  uint64 r = s;
  while (r <= e) {
    const uint64 o0 = perm.empty() ? snps[r] : perm[snps[r]];
    uint64 run = 1;
    while (r + run <= e && (perm.empty() ? snps[r + run] : perm[snps[r + run]]) == o0 + run) ++run;
    uint64 off = 3 + o0 * bpr, left = run * bpr;
    uint8_t* d = dst + (r - s) * bpr;
    while (left) {
      const ssize_t got = ::pread(fd, d, left, (off_t)off);
      if (got <= 0) throw std::runtime_error("short read from the bed file");
      d += got;
      off += got;
      left -= got;
    }
    r += run;
  }
}

int FileBed::read_block_cb(void* user, uint64_t s, uint64_t e, uint8_t* dst) {
  auto* self = static_cast<FileBed*>(user);
  try {
    Timer t;
    gather_rows(self->shard.snps, self->perm, s, e, self->bed_bytes_per_snp, dst, self->fd);
    self->readtime += t.reltime();
    return 0;
  } catch (const std::exception&) {
    return 1;
  }
}

void FileBed::attach_stream_source() { check(pcaone_set_reader_source(ctx, &FileBed::read_block_cb, this)); }

void FileBed::check_file_offset_first_var() {
  // the reference rewinds its ifstream to byte 3 (FilePlink.cpp:17-24); reads here are
  // positional (pread), so there is no file position to restore — only the header to trust
  unsigned char hdr[3];
  if (::pread(fd, hdr, 3, 0) != 3 || hdr[0] != 0x6c || hdr[1] != 0x1b || hdr[2] != 0x01)
    cao.error("Incorrect magic number in plink bed file.");
}

void FileBed::read_all() {
  // FilePlink.cpp:26-120 without the dense decode: packed rows -> pinned host -> HBM, then
  // allele frequencies (bit-exact integer counts + one division) on the device.
  tick.clock();
  const size_t bytes = nsnps_local * bed_bytes_per_snp;
  void* p = nullptr;
  if (pcaone_alloc_pinned(&p, bytes)) cao.error("cannot allocate page-locked host memory for the bed");
  pinned = (uint8_t*)p;
  gather_rows(shard.snps, perm, 0, nsnps_local - 1, bed_bytes_per_snp, pinned, fd);
  readtime += tick.reltime();
  check(pcaone_upload_bed(ctx, pinned, nsnps_local, 0));
  check(pcaone_allele_freq(ctx));
  uint64_t nmiss = 0;
  check(pcaone_missing_count(ctx, &nmiss));
  p_miss = (double)nmiss / ((double)nsamples * (double)nsnps_local);
  F.resize(nsnps_local);
  check(pcaone_get_F(ctx, F.data()));
  uint64 mono = 0;
  for (uint64 j = 0; j < nsnps_local; ++j) mono += (F(j) == 0.0 || F(j) == 1.0);
  if (mono) cao.warn("sites with MAF=0 found! remove them first! count:", mono);
  if (params.missme)
    cao.print(tick.date(), "the proportion of missingness is", p_miss);
  pcaone_free_pinned(pinned);  // the shard is resident in HBM now
  pinned = nullptr;
}

void FileBed::read_block_initial(uint64 start_idx, uint64 stop_idx, bool standardize) {
  // FilePlink.cpp:122-218: the decoded, centred (and optionally scaled) N x B block in G
  const uint64 B = stop_idx - start_idx + 1;
  if (G.rows() != nsamples || G.cols() != B) G.resize(nsamples, B);
  check(pcaone_decode_block(ctx, start_idx, stop_idx, standardize ? 1 : 0, 0, G.data()));
}

void FileBed::read_block_update(uint64 start_idx, uint64 stop_idx, const Mat2D& U, const Mat1D& svals,
                                const Mat2D& VT, bool standardize) {
  // FilePlink.cpp:220-298: as above with the EMU fill of missing calls from U diag(S) VT
  const uint64 B = stop_idx - start_idx + 1;
  if (G.rows() != nsamples || G.cols() != B) G.resize(nsamples, B);
  Mat2D V(VT.cols(), VT.rows());
  for (uint64 i = 0; i < VT.rows(); ++i)
    for (uint64 j = 0; j < VT.cols(); ++j) V(j, i) = VT(i, j);
  check(pcaone_set_usv(ctx, U.data(), svals.data(), V.data()));
  check(pcaone_decode_block(ctx, start_idx, stop_idx, standardize ? 1 : 0, 1, G.data()));
}

// ---------------------------------------------------------------------------- FileBin
FileBin::FileBin(const Param& p) : Data(p) {
  cao.print(tick.date(), "start parsing binary format (LD residuals)");
  const int fd2 = ::open(p.filein.c_str(), O_RDONLY);
  if (fd2 < 0) cao.error("Cannot open binary file.");
  uint32_t hdr[2];
  if (::pread(fd2, hdr, 8, 0) != 8) cao.error("binary file is too short");
  nsnps = hdr[0];
  nsamples = hdr[1];
  nsnps_local = nsnps;
  map_bytes = 8 + (size_t)nsnps * nsamples * 4;
  struct stat st;
  if (fstat(fd2, &st) != 0 || (size_t)st.st_size < map_bytes) cao.error("binary file is shorter than its header says");
  map = mmap(nullptr, map_bytes, PROT_READ, MAP_PRIVATE, fd2, 0);
  ::close(fd2);
  if (map == MAP_FAILED) cao.error("cannot map the binary file");
  rows = reinterpret_cast<const float*>(reinterpret_cast<const char*>(map) + 8);
  cao.print(tick.date(), "shape of input matrix (features x samples) is", nsnps, " x", nsamples);
}

FileBin::~FileBin() {
  if (map && map != MAP_FAILED) munmap(map, map_bytes);
}

void FileBin::prepare_ld() {
  shard.snps.clear();
  create_context();
}

}  // namespace pcaone_host
