/* pcaone_b200 — C-ABI of the B200-native randomized-SVD hot path of PCAone.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / Eigen
 * types. Every entry point names the reference interface it replaces
 * (file:line relative to Zilong-Li/PCAone v0.7.2). INTEGRATION.md shows the
 * `GpuNormalRsvdOpData / GpuFancyRsvdOpData : RsvdOpData` subclasses a PCAone maintainer
 * would add on top of these calls.
 *
 * Conventions
 *  - every call returns 0 on success, non-zero on failure; pcaone_last_error(ctx) gives
 *    the message (the reference throws std::runtime_error from cao.error, Logger.hpp:85-94;
 *    the host shim converts a non-zero status into that throw).
 *  - host matrices are column-major FP64 exactly as Eigen::MatrixXd::data() lays them out
 *    (Common.hpp:30-31): G is M x l, H/Omega are N x l, U is N x k, V is M x k.
 *  - packed genotypes are PLINK bed SNP-major rows of bpr = ceil(N/4) bytes, sample 4q+r in
 *    bits 2r..2r+1 of byte q (FilePlink.cpp:39-47), WITHOUT the 3-byte file header.
 *  - all calls are made from one host thread per context (the reference is single-threaded
 *    at this level). One context drives one GPU.
 *  - there is NO CPU fallback: creation fails when no CUDA device is usable.
 */
#ifndef PCAONE_B200_H_
#define PCAONE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pcaone_ctx pcaone_ctx;

enum { PCAONE_SVD_SSVD = 1, PCAONE_SVD_WINSVD = 2 };          /* --svd 1 / 2 (Cmd.cpp:41-45) */
/* GEMM arithmetic. FP64: DMMA tensor cores. INT8Xs: error-free Ozaki scheme on
 * the tcgen05 int8 tensor cores — Omega / G are rounded once to s signed 8-bit slices per entry
 * (8s-1 bits against the column maximum) and the products are then exact integers; ranges with
 * missing genotypes run each product as a (non-missing count, missing mask) pair on the same
 * kernels (mean imputation); EMU update passes add the fills of the missing calls in FP64 on top of
 * those products (FilePlink.cpp:246-259). */
enum { PCAONE_PREC_FP64 = 0, PCAONE_PREC_INT8X2 = 2, PCAONE_PREC_INT8X3 = 3, PCAONE_PREC_INT8X4 = 4 };
enum { PCAONE_SRC_RESIDENT = 0, PCAONE_SRC_HOST = 1, PCAONE_SRC_FILE = 2, PCAONE_SRC_DENSE = 3, PCAONE_SRC_DOSAGE = 4, PCAONE_SRC_GL = 5 };

/* Mirrors the fields of `Param` (Cmd.hpp:16-98) that the hot path reads. */
typedef struct pcaone_config {
  uint64_t nsamples;     /* N                                                        */
  uint64_t nsnps;        /* M owned by this context (the local SNP shard)             */
  uint64_t nsnps_total;  /* M of the whole job (== nsnps on one GPU); eigvals use it  */
  uint32_t k;            /* -k                                                       */
  uint32_t oversamples;  /* already max(oversamples,k) as Cmd.cpp:216 derives it      */
  uint32_t svd;          /* PCAONE_SVD_*                                             */
  uint32_t bands;        /* -w/--batches, 64                                         */
  uint32_t maxp;         /* --maxp 20                                                */
  double   tol;          /* --tol-rsvd 1e-4                                          */
  int32_t  ploidy;       /* 2, or 1 with --haploid                                   */
  int32_t  scale;        /* -C; -9 = standardize by sqrt(ploidy f(1-f))              */
  int32_t  emu;          /* --emu                                                    */
  int32_t  out_of_core;  /* -m > 0: walk start[]/stop[] blocks, no flipOmg in sSVD   */
  int32_t  precision;    /* PCAONE_PREC_*                                            */
  int32_t  device;       /* CUDA ordinal                                             */
  int32_t  rank, world;  /* sharded multi-GPU job: this context's rank / #ranks      */
  uint32_t maxiter;      /* --maxiter 100 (EM)                                       */
  double   tolem;        /* --tol-em 1e-5                                            */
  /* How the matrix is split over the ranks of a multi-GPU job. 0: by SNPs — every rank
   * owns nsnps of the nsnps_total SNPs and all nsamples samples; the N x l partial H is summed over
   * the ranks before each Omega update. 1: by samples — every rank owns nsamples of the
   * nsamples_total samples (a byte-column range of every bed row, sample_offset % 4 == 0) and ALL
   * SNPs; the exact int64 partial sums of G_b = X_b^T Omega (window x l) are summed instead, H and
   * Omega stay row-sharded, and the orthonormalisation exchanges only l x l Gram matrices. The
   * exchange per Omega update is N x l doubles in mode 0, (window SNPs) x l in mode 1: hosts pick
   * mode 1 when the windows are shorter than N (configs[2]: 7.8k-SNP windows, N = 500k). In mode 1 every matrix with
   * one row per sample (Omega, H, U) holds THIS RANK's nsamples rows in pcaone_set_omega / pcaone_get_omega /
   * pcaone_get_GH / pcaone_get_usv; G, V, S and F are the whole job's and identical on every rank. */
  int32_t  shard_samples;
  uint64_t nsamples_total; /* N of the whole job (0 = nsamples)                        */
  uint64_t sample_offset;  /* first sample of this rank (shard_samples)                */
} pcaone_config;

/* Collective hook for SNP-sharded jobs: sum `count` doubles at device pointer `buf`
 * across ranks, ordered after prior work on `stream` (a cudaStream_t) and before later
 * work on it. The Python host installs torch.distributed.all_reduce (NCCL) here. */
typedef int (*pcaone_allreduce_fn)(void* user, void* buf, uint64_t count, void* stream);

/* Typed variant: `kind` says what to reduce — the sample-sharded mode also sums exact int64 partial
 * products and uint32 genotype counts and takes maxima of IEEE bit patterns. */
enum { PCAONE_RED_F64_SUM = 0, PCAONE_RED_I64_SUM = 1, PCAONE_RED_U64_MAX = 2, PCAONE_RED_U32_SUM = 3 };
typedef int (*pcaone_allreduce2_fn)(void* user, void* buf, uint64_t count, int kind, void* stream);

/* Block source for out-of-core passes: fill dst (pinned) with the packed rows of SNPs
 * start..stop inclusive (bpr bytes each). Replaces the ifstream.read of
 * FileBed::read_block_initial (FilePlink.cpp:125-136). */
typedef int (*pcaone_read_block_fn)(void* user, uint64_t start, uint64_t stop, uint8_t* dst);

/* ---- lifetime ------------------------------------------------------------------- */
int  pcaone_create(const pcaone_config* cfg, pcaone_ctx** out);
void pcaone_destroy(pcaone_ctx* ctx);
const char* pcaone_last_error(const pcaone_ctx* ctx); /* ctx may be NULL: creation errors */
int  pcaone_abi_version(void);
/* The arithmetic the context actually runs (PCAONE_PREC_*): an INT8Xs request whose s * (k + oversamples)
 * exceeds the 256 columns of one UMMA falls back to the FP64 DMMA kernels at creation. */
int  pcaone_precision(const pcaone_ctx* ctx);
void* pcaone_stream(pcaone_ctx* ctx);                  /* the cudaStream_t all work runs on */
int  pcaone_sync(pcaone_ctx* ctx);
int  pcaone_set_allreduce(pcaone_ctx* ctx, pcaone_allreduce_fn fn, void* user);
/* In-library collectives (NCCL over NVLink, enqueued on the context's stream with no host
 * round trip): rank 0 calls pcaone_comm_unique_id, the host hands the 128 bytes to every rank
 * (one process per GPU), each calls pcaone_comm_init. A host that drives several GPUs from one
 * process creates the communicators itself (ncclCommInitAll) and attaches them. With a
 * communicator attached the allreduce hook above is not used. */
int  pcaone_set_allreduce2(pcaone_ctx* ctx, pcaone_allreduce2_fn fn, void* user);
int  pcaone_comm_unique_id(uint8_t* out128);
int  pcaone_comm_init(pcaone_ctx* ctx, const uint8_t* id128, int rank, int world);
int  pcaone_comm_attach(pcaone_ctx* ctx, void* nccl_comm);
/* Peer-memory mailboxes for sample-sharded jobs with one process per GPU: every rank exports a 64-byte CUDA IPC
 * handle of its mailbox, the host gathers the world's handles (rank order) and hands them to every rank. The
 * row-sharded Omega update then runs as ONE cooperative launch whose three small exchanges (two l x l Gram
 * matrices; flipOmg sums + Householder signs + column maxima) are NVLink stores into the peers' mailboxes inside the
 * kernel, instead of three NCCL launches between four kernel launches. PCAONE_PEER_EXCHANGE=0 keeps the NCCL form. */
int  pcaone_comm_peer_export(pcaone_ctx* ctx, uint8_t* out64);
int  pcaone_comm_peer_import(pcaone_ctx* ctx, const uint8_t* handles, int nranks);
/* Page-locked host buffers for the packed bed (FileBed::inbed, FilePlink.hpp:43-46): a host that
 * reads the .bed into one of these gets full-rate cudaMemcpyAsync in upload / streaming. */
int  pcaone_alloc_pinned(void** out, size_t bytes);
void pcaone_free_pinned(void* p);
int  pcaone_device_count(void);                        /* usable CUDA devices; 0 = none */

/* ---- genotype sources (Data::prepare, Data.cpp:14-85) ----------------------------- */
/* FileBed::read_all (FilePlink.cpp:26-120): keep the packed shard resident in HBM.
 * `packed` holds nsnps rows of bpr bytes; host or device pointer (device_ptr != 0). */
int pcaone_upload_bed(pcaone_ctx* ctx, const uint8_t* packed, uint64_t nsnps, int device_ptr);
/* Out-of-core from host memory: `packed` (nsnps x bpr, ideally pinned) stays on the host
 * and is streamed block by block through double-buffered cudaMemcpyAsync every pass. */
int pcaone_set_host_source(pcaone_ctx* ctx, const uint8_t* packed, uint64_t nsnps);
/* Same with rows `row_stride` bytes apart: `packed` points at this context's first byte of SNP 0
 * inside a wider bed (a sample shard reads bytes [sample_offset/4, +ceil(nsamples/4)) of every row).
 * Either call also tells the context that the data behind the plan is new: the HBM tile cache
 * (below) is refilled on the next pass.
 * Streamed blocks on the int8 route keep their re-tiled operands in HBM after the first pass when
 * they fit (2 x the packed bytes), so an out-of-core job whose bed fits a 180 GB B200 twice reads
 * the host link once; blocks that do not fit keep streaming every pass. */
int pcaone_set_host_source2(pcaone_ctx* ctx, const uint8_t* packed, uint64_t nsnps, uint64_t row_stride);
/* Out-of-core through a reader callback (file-backed FileBed). */
int pcaone_set_reader_source(pcaone_ctx* ctx, pcaone_read_block_fn fn, void* user);
/* Out-of-core straight from a .bed file (checks the magic 6c 1b 01, FilePlink.hpp:24-27);
 * snp_offset = first SNP of this context's shard inside the file. */
int pcaone_open_bed(pcaone_ctx* ctx, const char* bed_path, uint64_t snp_offset);
/* Block plan: start[]/stop[] inclusive SNP ranges of Data::prepare (Data.cpp:78-84),
 * band_factor as Data.cpp:70. In-core winSVD derives its own windows (Halko.cpp:180-194)
 * when this is not called. */
int pcaone_set_blocks(pcaone_ctx* ctx, const uint64_t* start, const uint64_t* stop, uint32_t nblocks,
                      uint32_t band_factor);
/* permute_matrix (RSVD.hpp:61-78) for resident shards: new row j = old row indices[j]
 * (packed rows, F). */
int pcaone_permute_resident(pcaone_ctx* ctx, const uint32_t* indices);

/* ---- decode / allele frequency (bit-exact, FilePlink.cpp:37-60,165-204) ------------ */
int pcaone_allele_freq(pcaone_ctx* ctx);                 /* resident / host source: all SNPs */
int pcaone_get_F(pcaone_ctx* ctx, double* F);            /* nsnps doubles                    */
int pcaone_set_F(pcaone_ctx* ctx, const double* F);      /* projection-style external AF     */
int pcaone_get_lookup(pcaone_ctx* ctx, double* lut4xM);  /* centered_geno_lookup, 4 x M      */
int pcaone_get_scale(pcaone_ctx* ctx, double* s);        /* sqrt(ploidy)/sqrt(F(1-F)) or 1   */
int pcaone_missing_count(pcaone_ctx* ctx, uint64_t* n);  /* Data::C.count() (FilePlink.cpp:112) */
/* FileBed::read_block_initial / read_block_update (FilePlink.cpp:122-298): the dense
 * N x (stop-start+1) block as the reference leaves it in data->G; update != 0 applies the
 * EMU fill from the U,S,V installed with pcaone_set_usv. */
int pcaone_decode_block(pcaone_ctx* ctx, uint64_t start, uint64_t stop, int standardize, int update,
                        double* out);

/* ---- RsvdOpData state (Halko.hpp:6-42) --------------------------------------------- */
int pcaone_set_flags(pcaone_ctx* ctx, int update, int standardize);   /* setFlags, Halko.hpp:32 */
int pcaone_set_omega(pcaone_ctx* ctx, const double* Omg);             /* initOmg result, N x l  */
int pcaone_get_omega(pcaone_ctx* ctx, double* Omg);
int pcaone_set_usv(pcaone_ctx* ctx, const double* U, const double* S, const double* V);
int pcaone_get_usv(pcaone_ctx* ctx, double* U, double* S, double* V); /* any may be NULL        */
int pcaone_get_GH(pcaone_ctx* ctx, double* G, double* H);             /* any may be NULL        */
int pcaone_set_H(pcaone_ctx* ctx, const double* H);

/* NormalRsvdOpData::computeGandH (Halko.cpp:99-153) / FancyRsvdOpData::computeGandH
 * (Halko.cpp:155-269): one power-iteration pass for epoch pi; G (M x l) and H (N x l) stay
 * on the device (read them with pcaone_get_GH). Dispatches on cfg.svd. */
int pcaone_compute_gandh(pcaone_ctx* ctx, int pi);
/* The dense stage of RsvdOpData::computeUSV for one epoch (Halko.cpp:55-70): QR(G) twice,
 * B = R^-T H^T, SVD(B); leaves Q2 in G, Ucur, sigma, U_B on the device. */
int pcaone_small_stage(pcaone_ctx* ctx);
/* RsvdOpData::computeUSV (Halko.cpp:46-97): the whole epoch loop incl. MEV stopping and the
 * winSVD minimum-epoch rule; diff_out/epochs_out may be NULL. */
int pcaone_compute_usv(pcaone_ctx* ctx, int maxp, double tol, double* diff_out, int* epochs_out);
/* EM driver of run_pca_with_halko (Halko.cpp:290-319) incl. flip_UV (Utils.cpp:118-154). */
int pcaone_run_em(pcaone_ctx* ctx, int* iters_out);
/* Omega = thinQ(H); flipOmg (Halko.cpp:121-123, RSVD.hpp:80-89) as a stand-alone call. */
int pcaone_orth_omega(pcaone_ctx* ctx, int flip);
/* mev(X, Y) (Utils.cpp:194-200) on host matrices rows x cols (col-major). */
int pcaone_mev(pcaone_ctx* ctx, const double* X, const double* Y, uint64_t rows, uint32_t cols, double* out);

/* ---- host helpers that reproduce the reference's libstdc++ random streams --------------- */
/* RsvdOpData::initOmg (Halko.cpp:15-23, RSVD.hpp:20-59): N x l, column-major, seeded
 * std::default_random_engine; gaussian != 0 -> N(0,1) (--rand 1), else U(-1,1). */
int pcaone_init_omega(uint64_t rows, uint32_t cols, int seed, int gaussian, double* out);
/* permute_matrix (RSVD.hpp:61-71): std::shuffle of 0..n-1 with the unseeded default engine. */
int pcaone_shuffle_indices(uint64_t n, uint32_t* out);

/* ---- LD r2 (LD.cpp:48-51, 450-473) ---------------------------------------------------- */
/* Windows ws[w] (lead SNP) / we[w] (#SNPs incl. lead) as divide_pos_by_window produces
 * (LD.cpp:154-168). G source: standardized-or-residual genotypes, N x M col-major doubles
 * on the host (G != NULL), or the resident packed shard centred by F (G == NULL).
 * r2_out receives sum_w (we[w]-1) values in the reference's output order. */
int pcaone_ld_r2(pcaone_ctx* ctx, const double* G, uint64_t nsnps, const int32_t* ws, const int32_t* we,
                 uint64_t nwin, double* r2_out);

/* The same tiles on the other operands of the reference's LD path, without the 8 N M-byte host matrix:
 *   PCAONE_LD_DENSE_F64       data = N x M column-centred doubles on the host (pcaone_ld_r2 with G != NULL)
 *   PCAONE_LD_PACKED          the resident packed shard centred by F          (pcaone_ld_r2 with G == NULL)
 *   PCAONE_LD_RESID_F32       data = the float32 rows of a `.residuals` file (`-B`): cast to double and
 *                             centred per SNP as FileBin::read_all does (FileBinary.cpp:21-30); 4 N M host bytes,
 *                             streamed chunk by chunk
 *   PCAONE_LD_PACKED_RESID    the resident packed shard turned into those residuals ON THE DEVICE: centred
 *                             decode, `G -= U S V^T` with the context's U, S, V when ld_stats == 0
 *                             (Data::write_residuals, Data.cpp:242-291), column-centred, rounded through float32
 *                             and centred again — bit for bit what `--ld` writes and `-B` reads back, minus the file
 *   PCAONE_LD_PACKED_PROJECT  the resident packed shard with `(I - U U^T) G` applied, data = U of `--USV`
 *                             (nsamples x ncols doubles, column-major; LD.cpp:491-496)
 * r2_out (may be NULL with keep_out) and af / r2_tol / keep_out (pruning; keep_out may be NULL) as above. */
enum { PCAONE_LD_DENSE_F64 = 0, PCAONE_LD_PACKED = 1, PCAONE_LD_RESID_F32 = 2, PCAONE_LD_PACKED_RESID = 3,
       PCAONE_LD_PACKED_PROJECT = 4 };
typedef struct pcaone_ld_source {
  int32_t kind;       /* PCAONE_LD_* */
  int32_t ld_stats;   /* PACKED_RESID: 0 = ancestry-adjusted (subtract U S V^T), 1 = standardized-genotype LD */
  const void* data;   /* host operand of the kinds that have one */
  uint32_t ncols;     /* PACKED_PROJECT: columns of U */
} pcaone_ld_source;
int pcaone_ld_r2_ex(pcaone_ctx* ctx, const pcaone_ld_source* src, uint64_t nsnps, const int32_t* ws, const int32_t* we,
                    uint64_t nwin, double* r2_out, const double* af, double r2_tol, uint8_t* keep_out);
/* Data::write_residuals (Data.cpp:242-291) for SNPs start..stop of the resident shard: out receives the
 * float32 rows of `<out>.residuals` ([stop - start + 1][nsamples], SNP-major); the host writes the 8-byte
 * header, seeks by the permutation (Data.cpp:264-267) and saves the .mbim. */
int pcaone_residuals_block(pcaone_ctx* ctx, uint64_t start, uint64_t stop, int ld_stats, float* out);

/* ---- IRAM operator: ArnoldiOpData::perform_op (Arnoldi.cpp:18-46, Arnoldi.hpp:6-34) ---------
 * y = sum over blocks G_b (G_b^T x): x_in, y_out are nsamples doubles on the host (what Spectra's
 * SymEigsSolver hands to perform_op). Uses the context's source (resident, streamed blocks,
 * dosages), allele frequencies and the pcaone_set_flags state (update => EMU fill from
 * pcaone_set_usv); a context with k = 1, oversamples = 0 makes it a GEMV-shaped pass. */
int pcaone_perform_op(pcaone_ctx* ctx, const double* x_in, double* y_out);

/* ---- downstream consumers of U, S, V: the two half products on their own ----------
 * pcaone_xt_times: out (nsnps x ncols) = X^T A for A = nsamples x ncols, plus sqnorm[j] = sum_i x_ij^2
 *   (NULL to skip) — `V.row(j) = U^T G.col(j); y_norm2(j) = G.col(j).squaredNorm()` of run_selection
 *   (Selection.cpp:16-34); the statistics on top (galinsky / pcadapt) stay host code.
 * pcaone_x_times:  out (nsamples x ncols) = X B for B = nsnps x ncols — `U = G * V` of
 *   run_projection option 1 (Projection.cpp:236-241; the caller scales V by 1 / S, and installs the
 *   reference panel's allele frequencies with pcaone_set_F as Data::prepare does for projection).
 * Column-major host matrices, ncols <= k + oversamples, X decoded under the pcaone_set_flags state
 * (FP64 kernels on every source). */
int pcaone_xt_times(pcaone_ctx* ctx, const double* A, uint32_t ncols, double* out, double* sqnorm);
int pcaone_x_times(pcaone_ctx* ctx, const double* B, uint32_t ncols, double* out);
/* out (nsamples x ncols) = C B with C the missing-call indicator of the packed source (C(i, j) = 1 iff the call of
 * sample i at SNP j is missing: the `data->C` of the reference) and B nsnps x ncols — the per-sample correction of
 * the normal equations of `--project 2` (solve_projection_scores, Projection.cpp:159-179: rows of V at the missing
 * calls are left out of each sample's least-squares problem). Same kernels as pcaone_x_times, decode table {0,1,0,0}. */
int pcaone_mask_times(pcaone_ctx* ctx, const double* B, uint32_t ncols, double* out);

/* ---- BGEN-style dosages (FileBgen::read_all / read_block_initial, FileBgen.cpp:15-168) ----------
 * The host keeps the container parsing (`var.minor_allele_dosage`, FileBgen.cpp:26) and hands over
 * what that call yields: one row of nsamples floats per variant, NaN = missing, SNP-major
 * [nsnps][nsamples]. The device keeps the floats (4 bytes per genotype) and fuses
 * value = NaN ? 0 : (d / 2 - F_j) [* sqrt(ploidy) / sqrt(F_j (1 - F_j))] into the operand load of the
 * same FP64 tensor-core products; pcaone_allele_freq computes F_j = mean(d / 2) over non-missing
 * (FileBgen.cpp:27-41), pcaone_decode_block returns the dense block for parity. After this call the
 * context behaves like a resident genotype shard (sSVD / winSVD, pcaone_permute_resident,
 * pcaone_set_blocks). precision must be PCAONE_PREC_FP64; --emu is rejected for this source. */
int pcaone_upload_dosage(pcaone_ctx* ctx, const float* dosage, uint64_t nsnps, int device_ptr);

/* ---- Beagle genotype likelihoods, PCAngsd (FileBeagle::read_all FileBeagle.cpp:14-68, emMAF_with_GL
 * Utils.cpp:745-775, Data::fit_with_pi Data.cpp:296-316, EM loop Halko.cpp:290-311) -------------------
 * The host keeps the gz text parsing (parse_beagle_file) and hands over the reference's own matrix
 * P: 2*nsamples x nsnps doubles, column-major, P(2i, j) / P(2i+1, j) = likelihoods of genotypes 0 / 1.
 * pcaone_gl_em_maf runs the allele-frequency EM from F = 0.25 (iters_out: EM steps taken).
 * The expected genotypes E = (p1 + 2 p2) / (p0 + p1 + p2) - 2 F are rebuilt on the device at pi == 0 of
 * every computeUSV — from F, or on update passes from the individual allele frequencies of the current
 * U, S, V — and feed the dense FP64 products; pcaone_run_em (with emu = 0) is the PCAngsd EM loop,
 * pcaone_decode_block returns E blocks. precision = PCAONE_PREC_FP64, single GPU. The GRM step after
 * the loop (pcangsd_standardize_E + the N x N covariance, Halko.cpp:320-334) is pcaone_gl_grm + pcaone_sym_svd. */
int pcaone_upload_gl(pcaone_ctx* ctx, const double* P, uint64_t nsnps, int device_ptr);
int pcaone_gl_em_maf(pcaone_ctx* ctx, uint32_t maxiter, double tolmaf, int* iters_out);

/* ---- generic dense matrix: RsvdOpOnePass / RsvdOnePass / RsvdOne (RSVD.hpp:92-362) ----------
 * The same passes on a dense FP64 matrix A (rows x cols, column-major, as Eigen hands it over)
 * instead of packed genotypes. The context is created with nsnps = max(rows, cols),
 * nsamples = min(rows, cols) (a wide matrix is used transposed, RSVD.hpp:113-121), k, oversamples
 * (here NOT forced to >= k: size = k + os, RSVD.hpp:109), precision = PCAONE_PREC_FP64,
 * out_of_core = 1 and svd = PCAONE_SVD_WINSVD when windows > 0. Omega (nsamples x (k + os)) comes
 * from the host RNG through pcaone_set_omega (pcaone_init_omega(n, size, 1, gaussian) reproduces
 * the default-seeded engine of RSVD.hpp:123). pcaone_dense_rsvd runs computeGandH(G, H, p[, windows])
 * and computeUSV; pcaone_get_usv then returns U = nsamples x k (svd.matrixV()), S, and
 * V = nsnps x k (G * svd.matrixU()): RsvdOne::matrixU() is V here when rows >= cols, else U.
 * finder: 1 = QR (implemented); 2 = LU range finder (RSVD.hpp:150-153) is rejected. */
int pcaone_upload_dense(pcaone_ctx* ctx, const double* A, uint64_t rows, uint64_t cols);
/* The in-core `data->G` of a non-genetic input (FileCsv::read_all, FileCsv.cpp:10-62): G is nsamples x nsnps doubles,
 * column-major (Eigen), already normalised / standardised by the host; never transposed, whatever the shape. The
 * context is created with that nsamples / nsnps and precision = PCAONE_PREC_FP64; computeGandH / computeUSV then run
 * on it like on any source (no allele frequencies, pcaone_set_flags(update = 0, standardize = 0)). */
int pcaone_upload_dense_data(pcaone_ctx* ctx, const double* G);
int pcaone_dense_rsvd(pcaone_ctx* ctx, uint32_t p, uint32_t windows, int finder);

/* LD pruning (ld_prune_big, LD.cpp:240-268 — SURVEY 8f-2): the same banded r2 tiles, kept on the
 * device and consumed by a greedy kernel that walks the windows in lead order. af: per-SNP allele
 * frequency (7th column of the .mbim) or NULL = always prune the partner; keep_out: nsnps bytes,
 * 1 = kept (what write_pruned_snp_ids splits into .ld.prune.in / .ld.prune.out). */
int pcaone_ld_prune(pcaone_ctx* ctx, const double* G, uint64_t nsnps, const int32_t* ws, const int32_t* we,
                    uint64_t nwin, const double* af, double r2_tol, uint8_t* keep_out);

/* ---- exact PCA (`--svd 3`, Main.cpp:180-217) and the PCAngsd GRM step (Halko.cpp:320-334) -----------
 * pcaone_sample_covariance: K_out (nsamples x nsamples doubles, column-major, host) = X X^T for the context's
 * source and pcaone_set_flags state — the reference's `data->G * data->G.transpose()` (the caller divides by
 * nsnps, Main.cpp:187 / Halko.cpp:324). Computed as panels of k + oversamples unit vectors through the operator
 * X (X^T .) on the FP64 kernels, whatever the context's GEMM precision; SNP-sharded contexts sum the panels.
 * pcaone_sym_svd: SVD of a symmetric n x n host matrix on the device (one-sided Jacobi): S_out descending
 * (= eigenvalues of a positive semi-definite A, what SelfAdjointEigenSolver returns in ascending order), U_out
 * n x n column-major (= the eigenvectors; JacobiSVD's matrixU up to sign). sweeps_out may be NULL. */
int pcaone_sample_covariance(pcaone_ctx* ctx, double* K_out);
/* The PCAngsd GRM step (Halko.cpp:320-326) after pcaone_run_em on a genotype-likelihood source: the expected
 * genotypes are re-standardised with the individual allele frequencies of the final U, S, V
 * (pcangsd_standardize_E, Data.cpp:364-407), C_out (nsamples x nsamples, column-major, host) = E E^T / nsnps with
 * its diagonal replaced by Dc / nsnps; Dc_out (nsamples) may be NULL. pcaone_sym_svd(C) then gives the
 * eigenvectors the reference writes to .eigvecs2 (JacobiSVD's matrixU, up to sign). */
int pcaone_gl_grm(pcaone_ctx* ctx, double* C_out, double* Dc_out);
int pcaone_sym_svd(pcaone_ctx* ctx, const double* A, uint64_t n, double* U_out, double* S_out, int* sweeps_out);

/* ---- measurement ------------------------------------------------------------------- */
typedef struct pcaone_timers {
  double gemm_g_ms, gemm_h_ms, orth_ms, small_ms, h2d_ms, allreduce_ms, decode_ms;
  uint64_t gemm_g_launches, gemm_h_launches, kernel_launches, h2d_bytes, d2h_bytes;
  uint64_t omega_updates;
  uint64_t tc_ranges, fp64_ranges; /* SNP ranges whose GEMMs ran on the int8 / FP64 tensor-core kernels */
  double tc_g_ms, tc_h_ms;         /* k_tc_gemm alone (inside gemm_g_ms / gemm_h_ms), G pass / H pass */
  double ld_ms;                    /* k_ld_tiles (pcaone_ld_r2) */
  uint64_t ld_tiles, ld_pairs;     /* 128 x 128 Gram tiles computed / r2 values produced */
  uint64_t tc_miss_ranges;         /* of tc_ranges: ranges with missing calls (count + mask GEMM pairs) */
  uint64_t cache_hits;             /* streamed blocks served from the HBM tile cache instead of the host */
  uint64_t tc_emu_ranges;          /* of tc_miss_ranges: EMU update passes (FP64 correction over the missing calls) */
  double emu_fix_ms;               /* k_emu_fix_g / k_emu_fix_h (inside gemm_g_ms / gemm_h_ms) */
} pcaone_timers;
int pcaone_get_timers(pcaone_ctx* ctx, pcaone_timers* out, int reset);
int pcaone_enable_timing(pcaone_ctx* ctx, int on); /* CUDA-event timing around the GEMM kernels */

#ifdef __cplusplus
}
#endif
#endif /* PCAONE_B200_H_ */
