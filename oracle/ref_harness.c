/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin driver around the *unmodified* reference libFLAC (compiled by
 * oracle/Makefile from the sources where they lie under /root/reference into
 * oracle/_ref/). It exposes a flat, ctypes-friendly C interface that
 *   - encodes a PCM buffer through FLAC__stream_encoder_* with in-memory
 *     callbacks and reports the byte length of every audio frame
 *     (the write callback delivers exactly one frame per call when samples>0,
 *     reference src/libFLAC/stream_encoder.c:3121), and
 *   - decodes a FLAC stream through FLAC__stream_decoder_* into interleaved PCM.
 *
 * It is used by tests/ (to pin oracle/flac_oracle.c and the CUDA path against
 * the real reference) and by bench.py's reference arm / cpu_baseline.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "FLAC/stream_encoder.h"
#include "FLAC/stream_decoder.h"
#include "share/private.h"

typedef struct {
	uint8_t *out;
	size_t cap, len;
	uint32_t *frame_sizes;
	size_t max_frames, nframes;
	size_t header_len;
	int overflow;
} enc_sink;

static FLAC__StreamEncoderWriteStatus enc_write_cb(const FLAC__StreamEncoder *e, const FLAC__byte buffer[], size_t bytes, uint32_t samples, uint32_t current_frame, void *client)
{
	enc_sink *s = (enc_sink *)client;
	(void)e; (void)current_frame;
	if(s->len + bytes > s->cap) {
		s->overflow = 1;
		return FLAC__STREAM_ENCODER_WRITE_STATUS_FATAL_ERROR;
	}
	if(s->out)
		memcpy(s->out + s->len, buffer, bytes);
	s->len += bytes;
	if(samples > 0) {
		if(s->nframes < s->max_frames && s->frame_sizes)
			s->frame_sizes[s->nframes] = (uint32_t)bytes;
		s->nframes++;
	}
	else
		s->header_len += bytes;
	return FLAC__STREAM_ENCODER_WRITE_STATUS_OK;
}

/* Options beyond the compression level; negative/zero = leave the preset alone. */
typedef struct {
	int32_t exhaustive;      /* -1 keep, 0/1 set */
	int32_t mid_side;        /* -1 keep, 0/1 set */
	int32_t loose_mid_side;  /* -1 keep, 0/1 set */
	int32_t max_lpc_order;   /* -1 keep */
	int32_t qlp_precision;   /* -1 keep */
	int32_t min_part_order;  /* -1 keep */
	int32_t max_part_order;  /* -1 keep */
	int32_t disable_isa;     /* value for FLAC__stream_encoder_disable_instruction_set, 0 none */
	int32_t streamable_subset; /* -1 keep, 0/1 */
	int32_t limit_min_bitrate; /* -1 keep, 0/1 */
	int32_t prec_search;     /* -1 keep, 0/1: do_qlp_coeff_prec_search */
	const char *apodization; /* NULL keep */
} ref_enc_opts;

/*
 * Encode interleaved int32 PCM. Returns 0 on success, else a negative code
 * (-1 alloc, -2 init (init status in *aux), -3 process, -4 finish, -5 output overflow).
 * out may be NULL (timing only) as long as out_cap is large enough to count into.
 */
int ref_encode(const int32_t *interleaved, uint64_t samples_per_channel,
               uint32_t channels, uint32_t bps, uint32_t sample_rate,
               uint32_t level, uint32_t blocksize, uint32_t num_threads, int do_md5,
               const ref_enc_opts *opts,
               uint8_t *out, size_t out_cap, size_t *out_len, size_t *header_len,
               uint32_t *frame_sizes, size_t max_frames, size_t *nframes, int *aux)
{
	enc_sink sink;
	FLAC__StreamEncoder *e = FLAC__stream_encoder_new();
	FLAC__StreamEncoderInitStatus st;
	uint64_t done = 0;
	const uint64_t chunk = 1u << 16;
	int rc = 0;
	if(!e) return -1;
	memset(&sink, 0, sizeof sink);
	sink.out = out; sink.cap = out_cap; sink.frame_sizes = frame_sizes; sink.max_frames = max_frames;

	FLAC__stream_encoder_set_channels(e, channels);
	FLAC__stream_encoder_set_bits_per_sample(e, bps);
	FLAC__stream_encoder_set_sample_rate(e, sample_rate);
	FLAC__stream_encoder_set_compression_level(e, level);
	if(blocksize) FLAC__stream_encoder_set_blocksize(e, blocksize);
	FLAC__stream_encoder_set_verify(e, false);
	FLAC__stream_encoder_set_do_md5(e, do_md5 ? true : false);
	FLAC__stream_encoder_set_num_threads(e, num_threads ? num_threads : 1);
	if(opts) {
		if(opts->exhaustive >= 0) FLAC__stream_encoder_set_do_exhaustive_model_search(e, opts->exhaustive);
		if(opts->mid_side >= 0) FLAC__stream_encoder_set_do_mid_side_stereo(e, opts->mid_side);
		if(opts->loose_mid_side >= 0) FLAC__stream_encoder_set_loose_mid_side_stereo(e, opts->loose_mid_side);
		if(opts->max_lpc_order >= 0) FLAC__stream_encoder_set_max_lpc_order(e, (uint32_t)opts->max_lpc_order);
		if(opts->qlp_precision >= 0) FLAC__stream_encoder_set_qlp_coeff_precision(e, (uint32_t)opts->qlp_precision);
		if(opts->min_part_order >= 0) FLAC__stream_encoder_set_min_residual_partition_order(e, (uint32_t)opts->min_part_order);
		if(opts->max_part_order >= 0) FLAC__stream_encoder_set_max_residual_partition_order(e, (uint32_t)opts->max_part_order);
		if(opts->disable_isa > 0) FLAC__stream_encoder_disable_instruction_set(e, opts->disable_isa);
		if(opts->streamable_subset >= 0) FLAC__stream_encoder_set_streamable_subset(e, opts->streamable_subset);
		if(opts->limit_min_bitrate >= 0) FLAC__stream_encoder_set_limit_min_bitrate(e, opts->limit_min_bitrate);
		if(opts->prec_search >= 0) FLAC__stream_encoder_set_do_qlp_coeff_prec_search(e, opts->prec_search);
		if(opts->apodization) FLAC__stream_encoder_set_apodization(e, opts->apodization);
	}

	st = FLAC__stream_encoder_init_stream(e, enc_write_cb, NULL, NULL, NULL, &sink);
	if(st != FLAC__STREAM_ENCODER_INIT_STATUS_OK) {
		if(aux) *aux = (int)st;
		FLAC__stream_encoder_delete(e);
		return -2;
	}
	while(done < samples_per_channel) {
		uint64_t n = samples_per_channel - done;
		if(n > chunk) n = chunk;
		if(!FLAC__stream_encoder_process_interleaved(e, interleaved + done * channels, (uint32_t)n)) {
			if(aux) *aux = (int)FLAC__stream_encoder_get_state(e);
			rc = -3;
			break;
		}
		done += n;
	}
	if(!FLAC__stream_encoder_finish(e) && rc == 0) {
		if(aux) *aux = (int)FLAC__stream_encoder_get_state(e);
		rc = -4;
	}
	FLAC__stream_encoder_delete(e);
	if(sink.overflow) rc = -5;
	if(out_len) *out_len = sink.len;
	if(header_len) *header_len = sink.header_len;
	if(nframes) *nframes = sink.nframes;
	return rc;
}

/* ------------------------------------------------------------------ one encoder per host thread
 *
 * The "all host cores" arm for a batch of independent blocks/files: `workers` pthreads, each owning its
 * own FLAC__StreamEncoder (num_threads = 1) and encoding a contiguous range of `blocksize`-sample blocks
 * as its own stream ("also one process per file across all cores"). One
 * libFLAC encoder with set_num_threads(64) serialises frame hand-off through one thread and scales
 * far worse than this. Output bytes are discarded; *seconds = wall time from the first thread's start
 * to the last thread's join (CLOCK_MONOTONIC). Returns 0 or the first failing worker's code. */
typedef struct {
	const int32_t *pcm;
	uint64_t samples;
	uint32_t channels, bps, sample_rate, level, blocksize;
	const ref_enc_opts *opts;
	size_t cap;
	int rc;
	size_t nframes, out_len;
} par_job;

static void *par_worker(void *arg)
{
	par_job *j = (par_job *)arg;
	size_t out_len = 0, header_len = 0, nframes = 0;
	int aux = 0;
	j->rc = ref_encode(j->pcm, j->samples, j->channels, j->bps, j->sample_rate, j->level, j->blocksize, 1, 0, j->opts,
	                   NULL, j->cap, &out_len, &header_len, NULL, 0, &nframes, &aux);
	j->nframes = nframes;
	j->out_len = out_len - header_len;
	return NULL;
}

int ref_encode_parallel(const int32_t *interleaved, uint64_t nblocks, uint32_t channels, uint32_t bps, uint32_t sample_rate,
                        uint32_t level, uint32_t blocksize, uint32_t workers, const ref_enc_opts *opts,
                        double *seconds, uint64_t *frames_total, uint64_t *bytes_total)
{
	pthread_t *th;
	par_job *jobs;
	struct timespec t0, t1;
	uint32_t w, started = 0;
	int rc = 0;
	uint64_t first = 0;
	if(workers == 0 || blocksize == 0) return -1;
	if(workers > nblocks) workers = (uint32_t)(nblocks ? nblocks : 1);
	th = (pthread_t *)calloc(workers, sizeof *th);
	jobs = (par_job *)calloc(workers, sizeof *jobs);
	if(!th || !jobs) { free(th); free(jobs); return -1; }
	for(w = 0; w < workers; w++) {
		const uint64_t nb = nblocks / workers + (w < nblocks % workers ? 1 : 0);
		jobs[w].pcm = interleaved + first * blocksize * channels;
		jobs[w].samples = nb * blocksize;
		jobs[w].channels = channels; jobs[w].bps = bps; jobs[w].sample_rate = sample_rate;
		jobs[w].level = level; jobs[w].blocksize = blocksize; jobs[w].opts = opts;
		jobs[w].cap = (size_t)1 << 62;
		first += nb;
	}
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for(w = 0; w < workers; w++) {
		if(pthread_create(&th[w], NULL, par_worker, &jobs[w]) != 0) { rc = -1; break; }
		started++;
	}
	for(w = 0; w < started; w++) pthread_join(th[w], NULL);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if(frames_total) *frames_total = 0;
	if(bytes_total) *bytes_total = 0;
	for(w = 0; w < started; w++) {
		if(jobs[w].rc != 0 && rc == 0) rc = jobs[w].rc;
		if(frames_total) *frames_total += jobs[w].nframes;
		if(bytes_total) *bytes_total += jobs[w].out_len;
	}
	if(seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
	free(th); free(jobs);
	return rc;
}

/* ------------------------------------------------------------------ decode */

typedef struct {
	const uint8_t *in;
	size_t len, pos;
	int32_t *out;          /* interleaved */
	uint64_t cap_samples;  /* per channel */
	uint64_t written;      /* per channel */
	uint32_t channels, bps, sample_rate;
	int errors, overflow;
} dec_ctx;

static FLAC__StreamDecoderReadStatus dec_read_cb(const FLAC__StreamDecoder *d, FLAC__byte buffer[], size_t *bytes, void *client)
{
	dec_ctx *c = (dec_ctx *)client;
	size_t n = c->len - c->pos;
	(void)d;
	if(n == 0) { *bytes = 0; return FLAC__STREAM_DECODER_READ_STATUS_END_OF_STREAM; }
	if(n > *bytes) n = *bytes;
	memcpy(buffer, c->in + c->pos, n);
	c->pos += n;
	*bytes = n;
	return FLAC__STREAM_DECODER_READ_STATUS_CONTINUE;
}

static FLAC__StreamDecoderWriteStatus dec_write_cb(const FLAC__StreamDecoder *d, const FLAC__Frame *frame, const FLAC__int32 *const buffer[], void *client)
{
	dec_ctx *c = (dec_ctx *)client;
	const uint32_t bs = frame->header.blocksize, ch = frame->header.channels;
	uint32_t i, k;
	(void)d;
	c->channels = ch; c->bps = frame->header.bits_per_sample; c->sample_rate = frame->header.sample_rate;
	if(c->written + bs > c->cap_samples) { c->overflow = 1; return FLAC__STREAM_DECODER_WRITE_STATUS_ABORT; }
	if(c->out)
		for(i = 0; i < bs; i++)
			for(k = 0; k < ch; k++)
				c->out[(c->written + i) * ch + k] = buffer[k][i];
	c->written += bs;
	return FLAC__STREAM_DECODER_WRITE_STATUS_CONTINUE;
}

static void dec_error_cb(const FLAC__StreamDecoder *d, FLAC__StreamDecoderErrorStatus status, void *client)
{
	(void)d; (void)status;
	((dec_ctx *)client)->errors++;
}

/*
 * Decode a complete FLAC stream (with "fLaC" + metadata). Returns 0 on success.
 * info[0..3] = channels, bps, sample_rate, decoder error count.
 */
int ref_decode(const uint8_t *stream, size_t len, int32_t *out_interleaved, uint64_t cap_samples_per_channel,
               uint64_t *samples_per_channel, uint32_t info[4], int md5_check)
{
	dec_ctx c;
	FLAC__StreamDecoder *d = FLAC__stream_decoder_new();
	int rc = 0;
	if(!d) return -1;
	memset(&c, 0, sizeof c);
	c.in = stream; c.len = len; c.out = out_interleaved; c.cap_samples = cap_samples_per_channel;
	FLAC__stream_decoder_set_md5_checking(d, md5_check ? true : false);
	if(FLAC__stream_decoder_init_stream(d, dec_read_cb, NULL, NULL, NULL, NULL, dec_write_cb, NULL, dec_error_cb, &c) != FLAC__STREAM_DECODER_INIT_STATUS_OK) {
		FLAC__stream_decoder_delete(d);
		return -2;
	}
	if(!FLAC__stream_decoder_process_until_end_of_stream(d))
		rc = -3;
	if(!FLAC__stream_decoder_finish(d) && rc == 0)
		rc = -4; /* MD5 mismatch */
	FLAC__stream_decoder_delete(d);
	if(c.overflow) rc = -5;
	if(samples_per_channel) *samples_per_channel = c.written;
	if(info) { info[0] = c.channels; info[1] = c.bps; info[2] = c.sample_rate; info[3] = (uint32_t)c.errors; }
	return rc;
}

const char *ref_version(void) { return FLAC__VERSION_STRING; }
