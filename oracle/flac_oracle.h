/*
 * flac_oracle.h -- TEST INFRASTRUCTURE ONLY. Never imported/linked by the product path
 * (flac_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it.
 *
 * A plain-C, single-threaded restatement of the reference libFLAC (xiph/flac 1.5.0,
 * /root/reference) per-frame encode pipeline and frame decoder. Every function in
 * flac_oracle.c cites the reference file:line it follows.
 *
 * Floating point semantics = SOURCE ORDER (what the reference computes when its sources
 * are compiled without -fassociative-math: oracle/_ref/libFLAC_ref_strict.so). The
 * shipped-flags build (oracle/_ref/libFLAC_ref.so) lets GCC reassociate the
 * autocorrelation/Levinson sums; tests report parity against both.
 *
 * Parity status: PINNED -- tests/test_oracle_vs_reference.py compares every frame this
 * oracle emits with the frames of the compiled reference on all input families.
 */
#ifndef FLAC_ORACLE_H
#define FLAC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FO_MAX_CHANNELS 8
#define FO_MAX_LPC_ORDER 32
#define FO_MAX_APODIZATIONS 32
#define FO_MAX_PARTITIONS 256 /* partition order <= 8 */

/* FLAC__ApodizationFunction (src/libFLAC/include/protected/stream_encoder.h:45-65), oracle numbering */
enum {
	FO_APOD_TUKEY = 0, FO_APOD_SUBDIVIDE_TUKEY = 1, FO_APOD_BARTLETT = 2, FO_APOD_BARTLETT_HANN = 3, FO_APOD_BLACKMAN = 4,
	FO_APOD_BLACKMAN_HARRIS = 5, FO_APOD_CONNES = 6, FO_APOD_FLATTOP = 7, FO_APOD_GAUSS = 8, FO_APOD_HAMMING = 9, FO_APOD_HANN = 10,
	FO_APOD_KAISER_BESSEL = 11, FO_APOD_NUTTALL = 12, FO_APOD_RECTANGLE = 13, FO_APOD_TRIANGLE = 14, FO_APOD_PARTIAL_TUKEY = 15,
	FO_APOD_PUNCHOUT_TUKEY = 16, FO_APOD_WELCH = 17
};

typedef struct {
	int32_t type;   /* FO_APOD_* */
	float p;        /* tukey p (for subdivide: already divided by parts); gauss stddev */
	int32_t parts;  /* subdivide_tukey parts */
	float start, end; /* partial_/punchout_tukey extents (fractions of the block) */
} fo_apodization;

typedef struct {
	uint32_t channels, bits_per_sample, sample_rate, blocksize;
	int32_t do_mid_side, loose_mid_side;
	uint32_t max_lpc_order;
	uint32_t qlp_coeff_precision; /* 0 = choose like the reference */
	int32_t do_qlp_coeff_prec_search;
	int32_t do_exhaustive_model_search;
	uint32_t min_residual_partition_order, max_residual_partition_order;
	uint32_t num_apodizations;
	fo_apodization apodizations[FO_MAX_APODIZATIONS];
	int32_t disable_constant_subframes, disable_fixed_subframes, disable_verbatim_subframes;
	int32_t limit_min_bitrate;
	/* 0: the reference's C routines (default; what the CUDA engine follows). 1: restate the x86 AVX2 dispatch of the fixed-order
	 * guess, which leaves the last (blocksize - 4) % 4 samples out of its error sums (fixed_intrin_avx2.c:138) -- only a short last
	 * block can tell the difference. */
	int32_t x86_avx2_fixed_guess;
} fo_config;

/* What the search decided for one subframe (debug/stage-level comparison with the CUDA path). */
typedef struct {
	int32_t type;            /* 0 constant, 1 verbatim, 2 fixed, 3 lpc */
	int32_t order;
	int32_t wasted_bits;
	int32_t subframe_bps;
	int32_t qlp_precision, qlp_shift;
	int32_t qlp_coeff[FO_MAX_LPC_ORDER];
	int32_t rice_method;     /* 0 RICE, 1 RICE2 */
	int32_t partition_order;
	int32_t rice_params[FO_MAX_PARTITIONS];
	uint32_t estimate_bits;
} fo_subframe_plan;

typedef struct {
	int32_t channel_assignment;  /* 0 independent, 1 left/side, 2 right/side, 3 mid/side */
	fo_subframe_plan sub[FO_MAX_CHANNELS];
	/* all four candidates of a stereo frame: L, R, M, S (valid when mid-side evaluated) */
	fo_subframe_plan cand[4];
	uint32_t cand_valid[4];
} fo_frame_plan;

typedef struct fo_encoder fo_encoder;

/* Fill cfg with the reference's preset `level` (0..8) (stream_encoder.c:117-140, 1873-1904). */
void fo_config_preset(fo_config *cfg, uint32_t channels, uint32_t bps, uint32_t sample_rate, uint32_t level, uint32_t blocksize);

/* Resolves blocksize / qlp precision defaults like init_stream_internal_ (stream_encoder.c:742-830).
 * Returns NULL when the configuration is outside what this oracle restates (bps > 24, bad sizes ...). */
fo_encoder *fo_encoder_new(const fo_config *cfg);
void fo_encoder_delete(fo_encoder *e);
const fo_config *fo_encoder_config(const fo_encoder *e);

/* Encode one frame of `blocksize` (<= cfg.blocksize) samples of interleaved int32 PCM.
 * Returns the frame length in bytes, 0 on error. plan may be NULL. */
size_t fo_encode_frame(fo_encoder *e, const int32_t *interleaved, uint32_t blocksize, uint32_t frame_number,
                       uint8_t *out, size_t out_cap, fo_frame_plan *plan);

/* Encode a whole PCM buffer as consecutive frames (last one short), frame numbers from 0. */
int fo_encode_stream(fo_encoder *e, const int32_t *interleaved, uint64_t samples_per_channel,
                     uint8_t *out, size_t out_cap, size_t *out_len,
                     uint32_t *frame_sizes, size_t max_frames, size_t *nframes);

/* Stage-level entry points (same arithmetic the frame encoder uses). */
void fo_window_tukey(float *window, int32_t L, float p);
int fo_window(const fo_apodization *a, float *window, int32_t L);     /* any FO_APOD_* ; 0 on unknown type */
void fo_config_set_apodization(fo_config *cfg, const char *spec);     /* FLAC__stream_encoder_set_apodization */
void fo_autocorrelation(const float *data, uint32_t data_len, uint32_t lag, double *autoc);
void fo_lp_coefficients(const double *autoc, uint32_t *max_order, float lp_coeff[][FO_MAX_LPC_ORDER], double *error);
uint32_t fo_best_order(const double *lpc_error, uint32_t max_order, uint32_t total_samples, uint32_t overhead_bits_per_order);
int fo_quantize_coefficients(const float *lp_coeff, uint32_t order, uint32_t precision, int32_t *qlp_coeff, int *shift);
uint8_t fo_crc8(const uint8_t *data, size_t len);
uint16_t fo_crc16(const uint8_t *data, size_t len);

/* ---- decode ---- */
typedef struct {
	uint32_t channels, bits_per_sample, sample_rate; /* from STREAMINFO, used when the header says "get from STREAMINFO" */
} fo_streaminfo;

/* Decode one frame starting at data[0]. Writes interleaved int32 PCM. Returns bytes consumed,
 * 0 on error (bad sync, CRC mismatch, unsupported). */
size_t fo_decode_frame(const uint8_t *data, size_t len, const fo_streaminfo *si,
                       int32_t *out_interleaved, size_t out_cap_samples_per_channel,
                       uint32_t *blocksize, uint32_t *channels, uint32_t *bps, uint64_t *frame_or_sample_number);

#ifdef __cplusplus
}
#endif
#endif
