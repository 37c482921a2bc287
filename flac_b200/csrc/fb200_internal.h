// fb200_internal.h -- shared declarations for the libflac_b200.so translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/flac_b200.h"

namespace fb200 {

void set_error(const char *fmt, ...);
const char *get_error();

#define FB_CUDA(call)                                                                              \
	do {                                                                                           \
		cudaError_t e_ = (call);                                                                   \
		if(e_ != cudaSuccess) {                                                                    \
			fb200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
			return FB200_ERR_CUDA;                                                                 \
		}                                                                                          \
	} while(0)

// ---- FLAC bitstream constants (reference src/libFLAC/format.c:117-154) ----
constexpr uint32_t kSubframeHeaderBits = 8;  // zero pad 1 + type 6 + wasted flag 1
constexpr uint32_t kQlpPrecisionLen = 4, kQlpShiftLen = 5;
constexpr uint32_t kEntropyTypeLen = 2, kRiceOrderLen = 4;
constexpr uint32_t kRiceParamLen = 4, kRice2ParamLen = 5;
constexpr uint32_t kRiceEscape = 15, kRice2Escape = 31;
constexpr uint32_t kMaxFixedOrder = 4;
constexpr uint32_t kMinQlpPrecision = 5, kMaxQlpPrecision = 15;
constexpr int kQlpPrecisionSteps = 11;  // precisions a precision search tries per order: 5 .. 15
constexpr uint32_t kMaxExtraResidualBps = 4;  // private/stream_encoder.h:44
constexpr int kMaxPartitionOrder = 8;         // engine scope (-0..-8 use <= 6)
constexpr int kMaxPartitions = 1 << kMaxPartitionOrder;

// subframe types as stored in plans
enum : int { SF_CONSTANT = 0, SF_VERBATIM = 1, SF_FIXED = 2, SF_LPC = 3 };

// One autocorrelation "section" of a block: a full window or a partial (subdivide_tukey) window.
struct DevSection {
	int win_off;     // offset of the window table (floats) for this apodization
	int partial;     // 0: out[i]=x[i]*w[i], i<bs.  1: partial window (lpc.c:82-94)
	int data_len;    // samples the autocorrelation runs over
	int part_size;   // partial: half length
	int data_shift;  // partial: first sample
};

// One LPC candidate in evaluation order (stream_encoder.c:4318-4392).
struct DevCand {
	int kind;  // 0: autoc = section[sec]; 1: punch-out: autoc[i<max] = section[root][i]-section[sec][i], autoc[max] = section[sec][max]
	int sec;
	int root;
};

struct SigMeta {
	int wasted;
	int bps;  // subframe bps (after wasted bits, +1 for side)
};

// Output of the LPC analysis kernel for one candidate slot.
struct CandDesc {
	int valid;
	int order;
	int precision;
	int shift;
	int wide;   // needs 64-bit accumulation (FLAC__lpc_max_prediction_before_shift_bps > 32)
	int limit;  // residual range must be checked (FLAC__lpc_max_residual_bps > 32)
	int pad0, pad1;
	int qlp[FB200_MAX_LPC_ORDER];
};

// What the search kernel decides for one signal of one block.
struct SubframePlan {
	int type, order, wasted, bps;
	int precision, shift, method, porder;
	uint32_t est_bits;
	int wide;
	int pad0, pad1;
	int qlp[FB200_MAX_LPC_ORDER];
	uint8_t params[kMaxPartitions];
};

// Per-launch constants (passed by value to every kernel).
struct EncK {
	int channels, bps, sample_rate;
	int bs, bs_stride, nsig;
	int do_ms, loose_ms;
	int max_order;       // min(cfg.max_lpc_order, bs-1)
	int lags;            // max_order + 1
	int lag_stride;      // doubles per (section,item) autocorrelation record
	int qlp_precision, exhaustive;
	int prec_search;     // do_qlp_coeff_prec_search: every order is tried at precisions 5 .. 15 (stream_encoder.c:4230-4243)
	int min_po, max_po;  // resolved for this blocksize (stream_encoder.c:3759-3761)
	int rice_limit;      // 15 (stream bps <= 16) or 31 (stream_encoder.c:4076)
	int dis_const, dis_fixed, dis_verb;
	int nsec, nwin, nslots;
	int slot_stride;     // bytes per frame slot in the staging buffer (multiple of 16)
	int slot_words;
	int emit3_words;     // word-buffer capacity of k_emit3: the largest frame the search can produce (tighter than slot_words)
	uint32_t first_frame;  // frame number of the call's first block
	uint32_t blk0;         // index (inside the call) of this launch's first block
	int f64b;            // k_search5: the second warp of a signal evaluates its candidates on the FP64 pipe
	int file_blocks;     // > 0: frame numbers restart every file_blocks blocks (many-file batches: one stream per file, stream_encoder.c:3772)
	int limit_min_bitrate;  // stream_encoder.c:3874: a frame must not consist of constant subframes only
	int sig_group;       // k_search5: signals per CTA when a block's signals are split over several CTAs (0: one CTA per block)
	const int *redo;     // search kernels, second pass of limit_min_bitrate: only blocks with redo[blk] != 0, only signals >= channels - 1
};

// ---- row layout shared by k_search5 and k_emit3: a signal lives in rows of R_T samples with a 36-word stride and one
// zero row in front (the history of row 0)
constexpr int kSearch4ZeroRow = 36;

// ---- k_emit3 (emit_kernel.cuh)
// device table of k_emit3's CRC-16 pass (uint16 entries): 4 x 256 slicing tables, combine multipliers x^(32 Lw 2^s) for odd
// Lw < 64 and s < 9, and their nibble-product tables
constexpr int kCrcLwRows = 32;
constexpr int kCrcMulBase = 1024 + kCrcLwRows * 9 + 32;
constexpr int kCrcTabEntries = kCrcMulBase + kCrcLwRows * 9 * 64;
constexpr unsigned kLbEpochMask = 0x3fffffu;  // decoupled look-back status word: value << 24 | epoch (22 bits) << 2 | flag
// Per-launch arguments of k_emit3 beyond EncK.
struct Emit3Args {
	const int32_t *pcm;                  // interleaved int32, block b at pcm + b * bs * channels
	const SigMeta *meta;                 // [nb * nsig]
	const int *blkflags;                 // [nb]
	const SubframePlan *plans;           // [nb * nsig]
	const uint16_t *crc_tab;             // 4 x 256
	uint8_t *out;
	unsigned long long out_cap;
	unsigned long long *offsets;         // offsets[0 .. nb] of this launch
	const unsigned long long *running_in;  // bytes emitted by earlier launches of the same call
	unsigned long long *running_out;
	unsigned long long *lookback;        // [>= nb] status words
	unsigned *ticket;                    // monotonically increasing across launches
	unsigned ticket_base;                // value of *ticket when this launch starts
	unsigned epoch;
	int nb;
	uint32_t *chan_assign_out;           // may be null
	int *err;
	// more than two channels: k_emit3<PAIR> packs channel pairs into staging regions, k_join splices them into frames
	uint32_t *stage;                     // [nb][npairs][pair_words]
	uint32_t *stage_bits;                // [nb][npairs] bits of each region (0xffffffff: the pair overflowed its region)
	int pair_words, join_words;
};

// Fixed-size shared state of one k_emit3 CTA (in front of the signal / word buffers).
struct Emit3Shared {
	unsigned long long mbar;
	unsigned long long off;      // this frame's byte offset in the output stream
	uint32_t scan[16];           // per-warp totals of the run scan
	uint32_t mlev[16];           // CRC combine multipliers x^(32 Lw 2^s)
	uint32_t part[16];           // per-warp CRC partials
	uint32_t hdr[4];             // frame header bytes (big-endian words, left aligned)
	int blk, zero_words, pad0, pad1;
	int warm[2][FB200_MAX_LPC_ORDER];  // the first 32 samples of each channel (warm-up samples / constant value)
};

__host__ __device__ inline size_t emit3_sig_bytes(int bs, int R_T, int nch)
{
	const size_t planar = (size_t)nch * (size_t)(kSearch4ZeroRow + (bs / R_T) * 36) * 4;
	const size_t raw = (size_t)bs * nch * 4;
	return ((planar > raw ? planar : raw) + 15) / 16 * 16;
}
__host__ __device__ inline size_t emit3_smem_bytes(int bs, int R_T, int nch, int slot_words)
{
	return (sizeof(Emit3Shared) + 15) / 16 * 16 + emit3_sig_bytes(bs, R_T, nch) + (size_t)(slot_words + 8) * 4;
}


// ---- launchers (one translation unit per kernel family; encoder.cu holds no device code)
// general_kernels.cu
void launch_unpack(const void *packed, int bytes_per_sample, int32_t *pcm, unsigned long long n, int bps, int *err, cudaStream_t st);
void launch_meta(const EncK &k, const int32_t *pcm, SigMeta *meta, int *blkflags, int nb, cudaStream_t st);
void launch_prep(const EncK &k, const int32_t *pcm, int32_t *sig, SigMeta *meta, int *blkflags, int nb, cudaStream_t st);
void launch_autoc_general(const EncK &k, const int32_t *sig, const SigMeta *meta, const float *windows, const DevSection *secs, double *autoc, int nitems, cudaStream_t st);
void launch_minbr_flags(const EncK &k, const SubframePlan *plans, const int *blkflags, int nb, int *flags, cudaStream_t st);
void launch_lpc(const EncK &k, const double *autoc, const DevCand *cands, SigMeta *meta, const uint32_t *sigor, CandDesc *cdesc, int nitems, int autoc_unshifted, cudaStream_t st);
void launch_search_general(const EncK &k, size_t smem, const int32_t *sig, const SigMeta *meta, const CandDesc *cdesc, SubframePlan *plans, int nitems, cudaStream_t st);
void launch_emit_general(const EncK &k, size_t smem, const int32_t *sig, const int *blkflags, const SubframePlan *plans, uint8_t *slots, uint32_t *frame_bytes, uint32_t *chan_assign, int nb, cudaStream_t st);
void launch_scan(const uint32_t *bytes, int n, unsigned long long *offsets, unsigned long long *running, cudaStream_t st);
void launch_gather(const EncK &k, const uint8_t *slots, const uint32_t *bytes, const unsigned long long *offsets, uint8_t *out, unsigned long long capacity, int *err, int nb, cudaStream_t st);
void launch_debug_log(const double *dx, double *dy, int n);
void general_kernels_init(int device);  // raises the dynamic shared-memory limits once per device (never lowers them)
// autoc_kernel.cu
void launch_autoc3(const EncK &k, const int32_t *sig, const SigMeta *meta, const float *windows, const DevSection *secs, double *autoc, int nitems, cudaStream_t st);
void autoc3_init(int device);
void launch_autoc4(const EncK &k, const int32_t *pcm, const SigMeta *meta, const float *secwin, int secwin_stride, const DevSection *secs, double *autoc, int nitems, uint32_t *sigor, int or_sec, cudaStream_t st);
void autoc4_init(int device);
constexpr int kSecwinSlack = 160;  // zero floats behind every per-section weight table (>= the largest autocorrelation tile)
// search_kernel.cu
void launch_search5(const EncK &k, int rt, int maxord_t, int wps, size_t smem, const int32_t *pcm, const SigMeta *meta, const CandDesc *cdesc, SubframePlan *plans, int nb, cudaStream_t st);
size_t search5_smem(int bs, int rt, int nsig, int wps, int max_po);
void search5_init(int device);
// emit_kernel.cu
void launch_emit3_pairs(const EncK &k, int rt, int maxord_t, size_t pair_smem, size_t join_smem, const Emit3Args &a, int nb, cudaStream_t st);
size_t join_smem(int join_words);
void launch_emit3(const EncK &k, int rt, int maxord_t, size_t smem, const Emit3Args &a, int nb, cudaStream_t st);
void launch_crc16_tables(uint16_t *tab, cudaStream_t st);
void emit3_init(int device);

}  // namespace fb200
