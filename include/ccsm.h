/* ccsm.h -- C ABI of libccsm.so, the B200 (sm_100a) implementation of ccsmeth's per-site
 * methylation-call inference path.
 *
 * The reference (PengNi/ccsmeth, pure Python) has no FFI layer: the boundary this library replaces
 * is the nn.Module protocol used at three call sites --
 *     ccsmeth/call_modifications.py:315-369   model construction + checkpoint load (_call_mods_q)
 *     ccsmeth/call_modifications.py:201-214   model(16 tensors) -> (logits, probs)  (_call_mods2s)
 *     ccsmeth/call_mods_freq_bam.py:317-342, 301-302   AggrAttRNN load + model(offsets, histos)
 * and the forwards behind them --
 *     ccsmeth/models.py:89-150    ModelAttRNN.forward   (attbigru2s)
 *     ccsmeth/models.py:673-694   AggrAttRNN.forward
 *     ccsmeth/utils/attention.py:48-70   Attention.forward
 * Each entry point below names the reference lines it stands in for.  The Python classes in
 * ccsmeth_b200/models.py bind these symbols with ctypes (see INTEGRATION.md for the stub a
 * reference maintainer would add).
 *
 * Conventions
 *   - plain C types only; tensors cross as raw pointers + sizes (torch tensors: data_ptr()).
 *   - every function returns 0 on success or a negative CCSM_E* code; the message is available
 *     from ccsm_last_error() (thread-local, never NULL).  No exceptions / exit() cross the ABI.
 *   - "device" pointers are CUDA device pointers on the model's device; `stream` is a cudaStream_t
 *     passed as void* (NULL = legacy default stream).  Device entry points are asynchronous on
 *     that stream and never call cudaDeviceSynchronize.
 *   - a handle is bound to one device and is not thread-safe (the reference runs one model per
 *     process per GPU, call_modifications.py:572-578).
 *   - the caller owns all I/O buffers; the library owns packed weights and workspace.
 */
#ifndef CCSM_H_
#define CCSM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCSM_ABI_VERSION 1

/* error codes */
#define CCSM_OK            0
#define CCSM_EINVAL       -1   /* bad argument / unsupported configuration */
#define CCSM_ESTATE       -2   /* call order violated (e.g. forward before finalize) */
#define CCSM_ECUDA        -3   /* a CUDA runtime call failed (message has the CUDA error string) */
#define CCSM_ENOMEM       -4
#define CCSM_EUNSUPPORTED -5   /* device is not sm_100 / feature not built */
#define CCSM_EKEY         -6   /* unknown state_dict key or wrong shape */

/* model kinds */
#define CCSM_KIND_ATT2S 0      /* ModelAttRNN(model_type="attbigru2s"), models.py:17-150 */
#define CCSM_KIND_AGGR  1      /* AggrAttRNN(model_type="attbigru"),    models.py:625-694 */

/* arithmetic of the GEMM-shaped work (gate/softmax math is always fp32) */
#define CCSM_PREC_FP32    0    /* fp32 FFMA kernels: reference-exact to ~1e-6 */
#define CCSM_PREC_BF16X3  1    /* tcgen05, bf16 hi/lo split, 3 passes, fp32 accumulate (parity mode) */
#define CCSM_PREC_BF16    2    /* tcgen05, single-pass bf16 in / fp32 accumulate (throughput mode) */
#define CCSM_PREC_FP16X3  3    /* tcgen05, fp16 hi/lo split, 3 passes (parity mode, ~fp32-exact) */
#define CCSM_PREC_FP16    4    /* tcgen05, single-pass fp16 in / fp32 accumulate */
#define CCSM_PREC_FP16C8  5    /* tcgen05, fp16 main pass + two e4m3 (kind::f8f6f4) correction passes at half cost:
                                  2 pass-equivalents, max |dprob| ~3e-5 (parity mode, default) */

/* feature flags (reference models.py:35-47, CLI --is_npass/--is_stds/--is_sn/--is_map) */
#define CCSM_FEAT_NPASS 1
#define CCSM_FEAT_STDS  2
#define CCSM_FEAT_SN    4
#define CCSM_FEAT_MAP   8
/* cell type, carried in feat_flags: ModelAttRNN(model_type="attbilstm2s") (models.py:48-51) -- nn.LSTM instead of
 * nn.GRU: state_dict keys are the same with 4*hidden gate rows (i, f, g, o), the initial state is (h0, c0).  No
 * checkpoint ships for it; it runs in CCSM_PREC_FP32 only and through ccsm_forward_att2s_lstm. */
#define CCSM_CELL_LSTM  16
/* model class, carried in feat_flags: ModelAttRNN2 (model_type "attbigru2s2" / "attbilstm2s2", models.py:221-382) --
 * the kinetics are embedded as integers (ipd_embed / pw_embed: 953 x 8, npass_embed: 31 x 4 on clamp(npass, 1, 30))
 * and the head is classifier = Linear(4H, 4H) + ReLU + Linear(4H, classes).  state_dict keys: seq_embed.weight,
 * ipd_embed.weight, pw_embed.weight, npass_embed.weight, rnn.*, _att3.*, classifier.0.*, classifier.3.*.
 * CCSM_PREC_FP32 only; --is_stds / --is_sn / --is_map are not implemented for it. */
#define CCSM_MODEL_2S2  32
/* ModelTransEnc (model_type "transencoder2s", models.py:451-620): the ModelAttRNN2 embeddings, SrcEmbed (three
 * Conv1d(k=3) + BatchNorm1d + ReLU + MaxPool1d(k=3) stages), a learned positional embedding, `num_layers` post-norm
 * nn.TransformerEncoderLayer's (hidden = d_model, heads = (feat_flags >> 8) & 255, dim_feedforward taken from the
 * linear1 weight), mean over positions, the two-layer classifier.  state_dict keys as in the reference (BatchNorm
 * running_mean / running_var included; num_batches_tracked is not a float tensor and is not passed).  CCSM_PREC_FP32. */
#define CCSM_MODEL_TRANSENC 64
/* CCSM_KIND_AGGR only: AggrAttRNN(model_type="attbilstm") (models.py:640-643), ORed onto the bin count in feat_flags
 * (bins live in bits 0..7).  Runs on the layer-by-layer fp32 kernels, initial state (h0, c0). */
#define CCSM_AGGR_LSTM 0x100

typedef struct ccsm_model ccsm_model;

typedef struct ccsm_config {
  int32_t kind;         /* CCSM_KIND_* */
  int32_t seq_len;      /* 21 (att2s) / 11 (aggr) */
  int32_t num_layers;   /* 3 / 1 */
  int32_t hidden;       /* 256 / 32 */
  int32_t num_classes;  /* 2 / 1 */
  int32_t n_vocab;      /* 5 (att2s); ignored for aggr */
  int32_t n_embed;      /* 8 (att2s); ignored for aggr */
  int32_t feat_flags;   /* CCSM_FEAT_* (att2s); for aggr: bin count (20) */
  int32_t precision;    /* CCSM_PREC_* */
  int32_t device;       /* CUDA device ordinal */
} ccsm_config;

/* One strand's 8 forward tensors, in the reference's positional order
 * (models.py:89-90; call_modifications.py:201-208).  float32, row-major.  Dead slots may be NULL. */
typedef struct ccsm_strand {
  const float* kmer;       /* (n, L) base codes 0..4 stored as floats; truncated like .int() */
  const float* kpass;      /* (n, L)   used iff CCSM_FEAT_NPASS */
  const float* ipd_means;  /* (n, L) */
  const float* ipd_stds;   /* (n, L)   used iff CCSM_FEAT_STDS */
  const float* pw_means;   /* (n, L) */
  const float* pw_stds;    /* (n, L)   used iff CCSM_FEAT_STDS */
  const float* sns;        /* (n, 4)   used iff CCSM_FEAT_SN */
  const float* maps;       /* (n, L)   used iff CCSM_FEAT_MAP */
} ccsm_strand;

int         ccsm_abi_version(void);
const char* ccsm_last_error(void);

/* Number of CUDA kernels this library has launched in the calling process (for bench accounting). */
int64_t     ccsm_kernel_launches(void);

/* Replaces model construction, reference call_modifications.py:316-323 / call_mods_freq_bam.py:317-321. */
int  ccsm_create(ccsm_model** out, const ccsm_config* cfg);
void ccsm_destroy(ccsm_model* m);

/* Replaces load_state_dict, reference call_modifications.py:343-358.  `key` is a reference state_dict
 * key ("embed.weight", "rnn.weight_ih_l0_reverse", "_att3.Ua.weight", "fc1.bias", ...); a leading
 * "module." (DDP/DataParallel prefix) is stripped as the reference does (:350-358).
 * `host` is float32 host memory of `shape[0..ndim)`; it is copied. */
int  ccsm_set_weight(ccsm_model* m, const char* key, const float* host, const int64_t* shape, int32_t ndim);

/* Replaces `.cuda(device)` + `.eval()`, reference call_modifications.py:367-369: validates that every
 * tensor was provided, packs (transposes / splits hi-lo / tiles) the weights and uploads them.
 * May be called again after further ccsm_set_weight calls. */
int  ccsm_finalize(ccsm_model* m);

/* What a NULL h0 pointer means in the forward calls.
 *   CCSM_H0_ZEROS (default): zero initial state.
 *   CCSM_H0_DEVICE_RANDOM:   N(0,1) drawn on the device (Philox4x32-10, Box-Muller) inside the feature-packing
 *     kernel -- the reference draws torch.randn inside forward (models.py:77-87,125-130); this is the same
 *     distribution without materialising or transferring 12 KB/site of noise.  Value u of (site, strand, layer,
 *     direction) is output u of Philox subsequence ((site*2+strand)*2*layers + 2*layer+direction) at offset
 *     256*call, so every forward call sees fresh noise and all arithmetic modes see the same noise.
 *   CCSM_H0_TORCH_STREAM:    the reference's own stream, reproduced on the device bit for bit: what
 *     torch.manual_seed(seed) followed by the reference's per-model-call torch.randn(2*layers, n_call, hidden) draws
 *     (strand 1, then strand 2; models.py:77-87, seeded at call_modifications.py:479-481) yields on an AVX2-capable x86
 *     host -- MT19937 outputs, 24-bit uniforms and ATen's 16-wide Box-Muller with its single-precision log / sincos
 *     polynomials (csrc/mtstream.cu).  The generator state lives on the device and advances from call to call like the
 *     reference's process-wide generator.  How a call's n sites split into the reference's model calls is announced with
 *     ccsm_set_h0_batching; GRU att2s models only. */
#define CCSM_H0_ZEROS 0
#define CCSM_H0_DEVICE_RANDOM 1
#define CCSM_H0_TORCH_STREAM 2
int  ccsm_set_h0_mode(ccsm_model* m, int32_t mode, uint64_t seed);

/* CCSM_H0_TORCH_STREAM: the NEXT forward call covers n_holebatches consecutive hole-batches of holebatch_sites[i] sites,
 * each of which the reference cuts into model calls of at most batch_size sites (call_modifications.py:177-181, --batch_size).
 * One-shot (consumed by the next call; without it a call is one hole-batch); batch_size stays in force. */
int  ccsm_set_h0_batching(ccsm_model* m, const int64_t* holebatch_sites, int64_t n_holebatches, int32_t batch_size);

/* CCSM_H0_TORCH_STREAM: hand over / read back the MT19937 state (624 words + position of the next output inside the
 * block, 624 = regenerate first) -- the host mirror (ccsmeth_b200/models.py) uses it to draw from, and advance, torch's
 * own CPU generator, so that forward(h0=None) consumes the process-wide stream exactly like the reference's forward.
 * The getter waits for the generator's stream only. */
int  ccsm_h0_stream_set_state(ccsm_model* m, const uint32_t* words624, int32_t pos);
int  ccsm_h0_stream_get_state(ccsm_model* m, uint32_t* words624, int32_t* pos);

/* Change the arithmetic mode of a finalized model (re-packs weights if needed). */
int  ccsm_set_precision(ccsm_model* m, int32_t precision);

/* Replaces ModelAttRNN.forward, reference models.py:89-150 (call site call_modifications.py:201-208).
 * fwd/rev: device pointers.  h0_fwd/h0_rev: (2*layers, n, hidden) float32 device, the initial hidden
 * states the reference draws with torch.randn (models.py:77-87); NULL means zeros.
 * logits/probs: (n, num_classes) float32 device (either may be NULL). */
int  ccsm_forward_att2s(ccsm_model* m, int64_t n, const ccsm_strand* fwd, const ccsm_strand* rev,
                        const float* h0_fwd, const float* h0_rev, float* logits, float* probs, void* stream);

/* ModelAttRNN.forward of the LSTM variant (models.py:89-150 with rnn_cell == "lstm"): as ccsm_forward_att2s plus the
 * initial cell states c0_fwd / c0_rev, (2*layers, n, hidden) float32 device or NULL (zeros); the reference draws h0
 * then c0 with torch.randn per strand (models.py:77-87).  Calling ccsm_forward_att2s on an LSTM model uses c0 = 0. */
int  ccsm_forward_att2s_lstm(ccsm_model* m, int64_t n, const ccsm_strand* fwd, const ccsm_strand* rev,
                             const float* h0_fwd, const float* c0_fwd, const float* h0_rev, const float* c0_rev,
                             float* logits, float* probs, void* stream);

/* Same computation with HOST buffers (pageable or pinned): the library stages host->device copies,
 * the forward and the device->host copy of the results in double-buffered chunks on its own streams,
 * and returns when `logits`/`probs` are complete.  This is the call the batch loop
 * (reference call_modifications.py:170-227) makes once per hole-batch instead of once per 512 sites. */
int  ccsm_forward_att2s_host(ccsm_model* m, int64_t n, const ccsm_strand* fwd, const ccsm_strand* rev,
                             const float* h0_fwd, const float* h0_rev, float* logits, float* probs);

/* Replaces AggrAttRNN.forward, reference models.py:673-694 (call site call_mods_freq_bam.py:301).
 * offsets (n, L), histos (n, L, bins), h0 (2*layers, n, hidden) or NULL, out (n, num_classes); device. */
int  ccsm_forward_aggr(ccsm_model* m, int64_t n, const float* offsets, const float* histos,
                       const float* h0, float* out, void* stream);

/* The same forward from per-SITE rows: the kernel gathers each site's window of L neighbouring sites itself -- zero
 * histogram rows and positions pos[0] - 1000 / pos[n-1] + 1000 beyond the ends, offsets |pos_j - pos_centre|, or with
 * only_close the 0/1 "next CpG is 2 bp away" flags (reference call_mods_freq_bam.py:265-293) -- so the (n, L, bins + 1)
 * windows the reference materialises on the host never exist.  site_pos (n) int64, site_histo (n, bins) float32, h0
 * (2*layers, n, hidden) or NULL, out (n, num_classes); device.  Fused-kernel configurations only (GRU, hidden 32, one
 * layer, 20 bins); CCSM_EUNSUPPORTED otherwise. */
int  ccsm_forward_aggr_sites(ccsm_model* m, int64_t n, const int64_t* site_pos, const float* site_histo,
                             int32_t only_close, const float* h0, float* out, void* stream);

/* AggrAttRNN.forward with model_type="attbilstm" (models.py:640-643, 661-671): as ccsm_forward_aggr plus the initial
 * cell state c0, (2*layers, n, hidden) float32 device or NULL (zeros); the reference draws h0 then c0 with torch.randn.
 * ccsm_forward_aggr on an LSTM model uses c0 = 0. */
int  ccsm_forward_aggr_lstm(ccsm_model* m, int64_t n, const float* offsets, const float* histos,
                            const float* h0, const float* c0, float* out, void* stream);

/* ---- reads in, calls out: feature extraction on the device -----------------------
 * Replaces, for one batch of reads, the per-read host work of
 *     ccsmeth/extract_features.py:261-406   extract_features_from_double_strand_read
 *         (CodecV1 decode process_utils.py:426-449, per-read normalisation :181-199, motif scan
 *          process_utils.py:122-137, forward / reverse-complement windows :343-363)
 *     ccsmeth/call_modifications.py:73-123  _batch_feature_list2s (the 16-tensor layout)
 *     ccsmeth/call_modifications.py:222-224 prob_1_norm = round(p1 / (p0 + p1), 6)
 *     ccsmeth/_bam2modbam.py:187-208        _convert_locs_to_mmtag / _convert_probs_to_mltag
 * The caller hands over the reads as raw bytes (e.g. the inflated BAM records themselves) plus one
 * descriptor per read that says where the read's sequence and kinetics arrays start inside that blob. */
#define CCSM_READ_REVERSE   1   /* stored sequence is the reverse complement of the forward read (BAM flag 0x10):
                                   forward base i = complement(stored[len-1-i]); kinetics arrays are used as stored
                                   (extract_features.py:313-319 does not flip them) */
#define CCSM_READ_SEQ_4BIT  2   /* sequence is BAM 4-bit packed ("=ACMGRSVTWYHKDBN", high nibble first); else ASCII */

typedef struct ccsm_read {
  int64_t seq_off;            /* byte offsets into the blob */
  int64_t fi_off, ri_off;     /* IPD codes, forward / reverse strand, `len` bytes each (tags fi, ri) */
  int64_t fp_off, rp_off;     /* PW codes (tags fp, rp) */
  int32_t len;                /* read length in bases */
  int32_t fn, rn;             /* number of passes per strand (tags fn, rn) */
  int32_t flags;              /* CCSM_READ_* */
  int32_t win_lo, win_hi;     /* keep only sites with win_lo <= loc < win_hi (align mode + --skip_unmapped yes:
                                 the aligned part of the query, extract_features.py:374,388-390; else 0, len) */
  float   sn[4];              /* tag sn (used iff CCSM_FEAT_SN) */
} ccsm_read;

#define CCSM_NORM_ZSCORE  0    /* --norm, extract_features.py:181-199 */
#define CCSM_NORM_MINMEAN 1
#define CCSM_NORM_MINMAX  2
#define CCSM_NORM_NONE    3
#define CCSM_NORM_MAD     4    /* shift = np.median, scale = statsmodels.robust.scale.mad (statsmodels 0.14.0, environment.yml:13:
                                * median(|a - median(a)| / Gaussian.ppf(3/4)); the package is absent from this image, its
                                * published formula is restated) */

typedef struct ccsm_extract_opts {
  int32_t mod_loc;            /* --mod_loc */
  int32_t norm;               /* CCSM_NORM_* */
  int32_t decode;             /* 1: CodecV1 codes -> frames (default), 0: --no_decode */
  int32_t n_motifs;           /* expanded (ACGT-only) motifs, all of length motif_len, e.g. 1 x "CG" */
  int32_t motif_len;          /* <= 8 */
  char    motifs[64];         /* n_motifs * motif_len characters, concatenated */
} ccsm_extract_opts;

/* Step 1 (host buffers): upload the batch, compute the per-read normalisation statistics, scan the motif and
 * build the site list (ordered by read, then by position -- the reference's order).  Synchronous; *n_sites is
 * the number of candidate sites.  The batch stays resident in the handle until the next call. */
int  ccsm_reads_extract_host(ccsm_model* m, const ccsm_extract_opts* opts, const uint8_t* blob, int64_t blob_bytes,
                             const ccsm_read* reads, int32_t n_reads, int64_t* n_sites);

/* The site list of the resident batch: index of the read in the batch and 0-based position `loc` in the forward
 * read (host int32[n_sites] each; either may be NULL). */
int  ccsm_reads_sites(ccsm_model* m, int32_t* site_read, int32_t* site_loc);

/* Materialises sites [s0, s0+cn) of the resident batch in the reference's 16-tensor layout (device pointers,
 * written; dead slots may be NULL) -- what _batch_feature_list2s + the FloatTensor stacking produce. */
int  ccsm_reads_features(ccsm_model* m, int64_t s0, int64_t cn, const ccsm_strand* fwd_out,
                         const ccsm_strand* rev_out, void* stream);

/* Step 2 (host buffers): features -> forward -> per-site outputs for every site of the resident batch.
 * h0_fwd/h0_rev: host (2*layers, n_sites, hidden) or NULL (h0 mode of the handle).  Outputs are host arrays of
 * n_sites entries (any may be NULL): logits/probs (n, classes); prob1 = round(p1/(p0+p1), 6) in float32;
 * mm_delta = the MM-tag number of the site (count of uncalled C's of the forward read since the previous
 * called site of the read); ml = floor(prob1 * 256), 255 if prob1 >= 1. */
int  ccsm_reads_forward_host(ccsm_model* m, const float* h0_fwd, const float* h0_rev, float* logits, float* probs,
                             float* prob1, int32_t* mm_delta, uint8_t* ml);

/* ---- call_freqb on the device: one region's pileup -> per-site modification frequencies -------
 * Replaces _call_modfreq_of_one_region / _call_modfreq_of_one_region_aggregate_mode and their helpers
 * _cal_mod_prob, _cal_modfreq_in_count_mode, _get_normalized_histo, _cal_modfreq_in_aggregate_mode
 * (reference call_mods_freq_bam.py:102-107, 200-237, 265-305, 308-442).  The caller supplies the pileup of a region
 * in CSR form: site i (reference position refpos[i], ascending) is covered by entries [ptr[i], ptr[i+1]) of `ml`
 * (the ML byte of each read's call at that position) and `hap` (the read's haplotype tag value: 0 none, 1, 2; NULL =
 * all 0).  Building that pileup from an aligned modbam (region fetch, CIGAR walk, MM/ML parsing,
 * call_mods_freq_bam.py:457-594) is host work outside this library. */
typedef struct ccsm_pileup_opts {
  int32_t call_mode;    /* --call_mode: 0 count, 1 aggregate */
  int32_t cov_cf;       /* --cov_cf: sites with fewer calls than this are counted, not modelled (aggregate mode) */
  double  prob_cf;      /* --prob_cf */
  int32_t no_amb_cov;   /* --no_amb_cov */
  int32_t no_hap;       /* --no_hap: only the "all reads" group */
  int32_t discrete;     /* --discrete: ignored here -- discretize_score (:240-262) is scalar post-processing the caller
                           applies to the returned model frequencies */
  int32_t only_close;   /* --only_close: the 21st model input is "this neighbour is the CpG right after its predecessor"
                           (position distance 2) instead of the distance to the centre site (:285-290) */
} ccsm_pileup_opts;

/* The two tables the kernels use: prob[v] = _cal_mod_prob(v) and bin[v] = np.histogram bin of prob[v] (v = ML byte). */
int  ccsm_pileup_luts(const ccsm_pileup_opts* opts, int32_t bins, double* prob, int32_t* bin);

/* Step 1 (host buffers): upload the pileup, compute per-group coverage and the count-mode results, and find the
 * sites the aggregate model will see.  n_high[g] = their number for group g (0 all reads, 1 haplotype 1, 2 haplotype 2),
 * which is the n of the h0 tensors step 2 takes.  `m` is an aggregate (CCSM_KIND_AGGR) model. */
int  ccsm_pileup_begin_host(ccsm_model* m, const ccsm_pileup_opts* opts, int64_t n_sites, const int64_t* refpos,
                            const int64_t* ptr, const uint8_t* ml, const uint8_t* hap, int64_t* n_high);

/* Step 2 (host buffers): histograms -> fused aggregate model (windows formed in the kernel) -> results.
 * h0_*: host (2*layers, n_high[g], hidden) float32 or NULL (zeros).  The shipped configuration (GRU, hidden 32, 20 bins,
 * one layer) runs the fused kernel; any other AggrAttRNN shape materialises the windows and runs the fp32 layer kernels.  Outputs are (3, n_sites) arrays, group-major:
 * cov = coverage reported for the site (-1 = the group has no call there: the reference's None), cnt_mod, freq, and
 * (optional) kind = which value types the reference would hold: 0 None, 1 count path with an integer count, 2 count
 * path with np.round(len * freq, 2) (float64), 3 model path (cnt_mod and freq are float32 values). */
int  ccsm_pileup_finish_host(ccsm_model* m, const float* h0_all, const float* h0_hp1, const float* h0_hp2,
                             int32_t* cov, double* cnt_mod, double* freq, uint8_t* kind);

/* Step 2 for an LSTM aggregate model (CCSM_AGGR_LSTM): h0[g] / c0[g], g = 0 all reads, 1 haplotype 1, 2 haplotype 2,
 * are host (2*layers, n_high[g], hidden) float32 buffers or NULL (zeros); both arrays hold three pointers. */
int  ccsm_pileup_finish_lstm_host(ccsm_model* m, const float* const* h0, const float* const* c0,
                                  int32_t* cov, double* cnt_mod, double* freq, uint8_t* kind);

/* ---- host I/O helpers: BGZF block codec on a thread team (SAM/BAM spec 4.1) --------------------------------
 * The reference reads and writes BAM through pysam/htslib with `threads=` (extract_features.py:60-73,
 * call_modifications.py:410-462).  Pure host code; buffers are host memory.
 * ccsm_bgzf_inflated_size: sum of the inflated sizes of the COMPLETE blocks found in src; *consumed = their bytes.
 * ccsm_bgzf_inflate:       inflates those blocks into dst (CRC32 checked); returns bytes written.  Blocks go through
 *                          the library's table decoder (csrc/inflate_fast.h); zlib decodes any block it rejects or whose
 *                          CRC32 / size does not match afterwards (CCSM_INFLATE=zlib: zlib for every block).
 * ccsm_bgzf_inflate_stats: blocks decoded by the table decoder / by zlib so far in this process.
 * ccsm_bgzf_deflate:       cuts src into 65280-byte blocks, deflates them in parallel, writes the concatenated
 *                          BGZF blocks (no EOF marker) into dst (capacity >= ccsm_bgzf_deflate_bound(src_bytes));
 *                          returns bytes written.  `level` = zlib level 0..9, optionally ORed with CCSM_BGZF_RLE:
 *                          run-length matching + dynamic Huffman only (the token stream of zlib's Z_RLE, produced by
 *                          the library's own encoder csrc/deflate_rle.h; `level` is then ignored).  On HiFi records
 *                          (packed bases, qualities, kinetics: high-entropy bytes where LZ77 finds nothing) that is
 *                          6-8x faster than zlib's default strategy at the same size within 3 %.
 * Negative return = CCSM_E* code. */
#define CCSM_BGZF_RLE 0x100
int64_t ccsm_bgzf_inflated_size(const uint8_t* src, int64_t src_bytes, int64_t* consumed);
int64_t ccsm_bgzf_inflate(const uint8_t* src, int64_t src_bytes, uint8_t* dst, int64_t dst_cap, int32_t threads,
                          int64_t* consumed);
void    ccsm_bgzf_inflate_stats(int64_t* fast_blocks, int64_t* zlib_blocks);
int64_t ccsm_bgzf_deflate_bound(int64_t src_bytes);
int64_t ccsm_bgzf_deflate(const uint8_t* src, int64_t src_bytes, uint8_t* dst, int64_t dst_cap, int32_t level,
                          int32_t threads);

/* Replaces the record walk of `samtools sort` / `samtools index`, which the reference runs on its output through pysam
 * (call_modifications.py:592-607).  Walks the complete alignment records of an inflated BAM stream positioned at a
 * record boundary; per record: the coordinate sort key samtools uses, (uint32) refID << 32 | (pos + 1) << 1 | reverse
 * strand (refID -1 sorts last), the byte offset and length of the record including its 4-byte block_size, refID, pos,
 * end = pos + reference length of the CIGAR (pos + 1 for unmapped / CIGAR-less records) and the flag.  Returns the number
 * of records (at most max_recs); *consumed = bytes of the complete records walked.  Host code (ccsmeth_b200/bamsort.py
 * builds the sorted BAM and its .bai from it). */
int64_t ccsm_bam_scan_records(const uint8_t* buf, int64_t n_bytes, int64_t max_recs, uint64_t* key, int64_t* off,
                              int32_t* len, int32_t* ref_id, int32_t* pos, int32_t* end, int32_t* flag,
                              int64_t* consumed);

/* ---- host I/O helpers: BAM record indexing and re-tagging (SAM/BAM spec 4.2) -------------------------------
 * ccsm_bam_index walks the complete alignment records in an inflated BAM byte stream `buf` (positioned at a
 * record boundary) and fills, per record, a ccsm_bam_rec, and per read that takes part in calling, a ccsm_read
 * whose offsets address `buf` itself -- so `buf` is the blob ccsm_reads_extract_host takes, with no copy.  It stands
 * in for the pysam accessors of reference extract_features.py:88-126 and the read-level filters of :269-288,321-326.
 * *consumed = bytes of complete records seen (the caller carries the tail over to the next piece).
 * ccsm_bam_tag_records writes the records back (each with its 4-byte block_size) without MM/ML -- and without
 * fi/fp/ri/rp unless keep_pulse -- and, for read r with sites [site_begin[r], site_begin[r+1]), appends
 * "MM:Z:C+m?,<mm...>;" and "ML:B:C,<ml...>" (reference _bam2modbam.py:211-226, call_modifications.py:230-266);
 * pass mm = ml = NULL to write no tags at all.  Returns bytes written (or a negative CCSM_E* code). */
typedef struct ccsm_bam_rec {
  int64_t off;        /* offset of the record's block_size field in buf */
  int32_t len;        /* block_size */
  int32_t aux_off;    /* offset of the first aux field, from the start of the record body (off + 4) */
  int32_t flag, mapq, l_seq, n_cigar;
  int32_t read_idx;   /* index into the ccsm_read array, or -1: the read takes no part (filtered / no kinetics) */
  int32_t pad_;
} ccsm_bam_rec;

typedef struct ccsm_bam_filter {
  int32_t mode_align;        /* --mode align: drop unmapped / secondary / duplicate reads, apply mapq */
  int32_t mapq;              /* --mapq */
  int32_t no_supplementary;  /* --no_supplementary */
  int32_t skip_unmapped;     /* --skip_unmapped yes: only sites inside the aligned part of the query */
  int32_t want_sn;           /* --is_sn yes: copy the sn tag */
  int32_t pad_;
  double  identity;          /* --identity (--mode align): drop reads whose CIGAR identity, matches (M, =) over all
                                aligned operations but clips (process_utils.py:174-186), is below it */
} ccsm_bam_filter;

int     ccsm_bam_index(const uint8_t* buf, int64_t n_bytes, const ccsm_bam_filter* f, ccsm_bam_rec* recs,
                       int32_t max_recs, ccsm_read* descs, int32_t* n_recs, int32_t* n_descs, int64_t* consumed);
int64_t ccsm_bam_tag_records(const uint8_t* buf, const ccsm_bam_rec* recs, int32_t n_recs, int32_t keep_pulse,
                             const int64_t* site_begin, const int32_t* mm, const uint8_t* ml, uint8_t* out,
                             int64_t out_cap, int32_t* n_with_mm);

/* call_freqb, host half: walks aligned records (as indexed by ccsm_bam_index) and emits one tuple per modification
 * call of the read's MM/ML tags ("C+m", first entry) that sits on an aligned reference base -- the work of
 * _get_moddict_in_tags (reference call_mods_freq_bam.py:118-168) and of the read loop of
 * _readmods_to_bed_of_one_region (:466-520, matches-only aligned pairs, --base_clip on the pair list).  Arrays have
 * `cap` entries; the return value is the number of calls found (if > cap nothing beyond cap was written: retry), or a
 * negative CCSM_E* code.  strand: 0 forward, 1 reverse (+2 for --refsites_all zero calls); hap: the --hap_tag value if
 * 1 or 2, else 0. */
typedef struct ccsm_modcall_opts {
  int32_t mapq;              /* --mapq */
  int32_t no_supplementary;  /* --no_supplementary */
  int32_t base_clip;         /* --base_clip */
  char    hap_tag[4];        /* --hap_tag, two characters (default "HP") */
  double  identity;          /* --identity */
  /* --refsites_all (:497-520): every reference motif site a read spans without calling it counts as an unmodified
   * call (ML 0).  sites_fwd / sites_rev: one byte per reference base, all references concatenated, reference i at
   * [ref_off[i], ref_off[i+1]); non-zero = a motif site of that strand.  Such calls come back with bit 1 set in
   * `strand` (2 = forward, 3 = reverse) so that the caller can apply region-local rules. */
  int32_t refsites_all;
  int32_t n_refs;
  const int64_t* ref_off;
  const uint8_t* sites_fwd;
  const uint8_t* sites_rev;
} ccsm_modcall_opts;
int64_t ccsm_bam_modcalls(const uint8_t* buf, const ccsm_bam_rec* recs, int32_t n_recs, const ccsm_modcall_opts* opts,
                          int32_t* ref_id, int32_t* ref_pos, uint8_t* ml, uint8_t* hap, uint8_t* strand, int64_t cap,
                          int32_t* n_reads_used);

/* Introspection used by tests: copies the last layer-stack output of the most recent forward chunk.
 * Returns the number of floats written (<= cap) or a negative error. */
int64_t ccsm_debug_last_rnn_out(ccsm_model* m, float* host, int64_t cap);

/* Per-kernel-class device timing for bench.py's roofline: while enabled, every tensor-core kernel launch is
 * bracketed by CUDA events recorded on the launching stream.  ccsm_profile_read waits for them and returns,
 * per class {0: feature/h0 packing, 1: GRU layer 0, 2: GRU layers >= 1, 3: attention + head, 4: read scan (per-read
 * statistics + motif scan + site list; units = bases), 5: window gather (units = sites)}, the summed milliseconds,
 * the summed units processed and the launch count (arrays of nclass >= 4 entries; classes >= nclass are dropped),
 * then resets. */
int  ccsm_profile_enable(ccsm_model* m, int32_t on);
int  ccsm_profile_read(ccsm_model* m, double* ms, double* units, int64_t* launches, int32_t nclass);

/* Test hook (tensor-core path): unpacks layer `layer`'s output image of the most recent forward chunk into
 * (tiles*128 rows, L, 512) float32 host memory; row R = 2*site + strand.  Returns floats written. */
int64_t ccsm_debug_tc_layer_out(ccsm_model* m, int32_t layer, float* host, int64_t cap);

/* Test hook: one-CTA tcgen05 GEMM  D(128,N) = A(128,K) . B(N,K)^T  with operands rounded to bf16 (or fp16),
 * fp32 accumulate; pins the UMMA descriptor conventions the tensor-core path relies on.  Host pointers. */
int  ccsm_debug_umma_gemm(int32_t device, int32_t N, int32_t K, int32_t is_f16, int32_t swap_lbo_sbo,
                          const float* A, const float* B, float* D);

/* Test hook: CTA-pair tcgen05 GEMM (cta_group::2)  D(256,N) = A(256,K) . B(N,K)^T, plus a tcgen05.st check:
 * Z(256,32) = [columns 16..31 after zeroing | columns 0..15 untouched].  Host pointers. */
int  ccsm_debug_umma_pair_gemm(int32_t device, int32_t N, int32_t K, int32_t is_f16, const float* A,
                               const float* B, float* D, float* Z);
/* D (128, N) = fp16(A) . fp16(B)^T (kind::f16) + e4m3(A) . e4m3(B)^T (kind::f8f6f4) accumulated in the same TMEM
 * columns; A (128, K), B (N, K) host fp32, K a multiple of 32.  Pins the 8-bit operand layout and the mixed-kind
 * accumulation CCSM_PREC_FP16C8 relies on (tests/test_umma_gpu.py). */
int  ccsm_debug_umma_mixed_gemm(int32_t device, int32_t N, int32_t K, const float* A, const float* B, float* D);
/* Tensor-pipe rate probe (one CTA, M = 128, N columns, operands fixed in shared memory): cycles for `iters` rounds of
 * four MMAs.  mode 0 = kind::f16, 1 = kind::f8f6f4 e4m3, 2 = f16 f16 e4m3 e4m3, 3 = alternating, 4 = rounds alternate. */
int  ccsm_debug_umma_rate(int32_t device, int32_t N, int32_t mode, int32_t iters, int64_t* cycles);
/* torch.manual_seed(seed); torch.randn(skip); torch.randn(n) reproduced on the device -> out (host, n floats; skip and
 * n multiples of 16).  tests/test_h0_stream_gpu.py compares it with torch bit for bit. */
int  ccsm_debug_torch_randn(int32_t device, uint64_t seed, int64_t skip, int64_t n, float* out);
/* Host-only self-check of the MT19937 jump-ahead polynomials the sub-stream generator uses (csrc/mtjump.h): the state
 * 2^k * 2^21 words ahead obtained through x^(2^k 2^21) mod phi against plain generation from `seed`.  Returns the number of
 * mismatching state words (0 = pass) or a negative CCSM_E* code.  No GPU needed. */
int  ccsm_debug_mt_jump_check(uint32_t seed, int32_t k);

#ifdef __cplusplus
}
#endif
#endif /* CCSM_H_ */
