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
	int min_po, max_po;  // resolved for this blocksize (stream_encoder.c:3759-3761)
	int rice_limit;      // 15 (stream bps <= 16) or 31 (stream_encoder.c:4076)
	int dis_const, dis_fixed, dis_verb;
	int nsec, nwin, nslots;
	int slot_stride;     // bytes per frame slot in the staging buffer (multiple of 16)
	int slot_words;
	uint32_t first_frame;
};

}  // namespace fb200
