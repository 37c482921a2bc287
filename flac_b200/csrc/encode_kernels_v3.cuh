// encode_kernels_v3.cuh -- k_search3: one WARP per (block, signal).
//
// The v2 search kernel spends a large part of its time at block barriers (3 warps/SMSP, two
// barriers per candidate, warp 0 doing the serial tail). Here a warp owns a signal outright:
//   * the signal sits in the warp's slice of shared memory; the block is cut into tiles of
//     32 lanes x R_T samples, a lane keeps its R_T-sample run + MAXORD history in registers;
//   * a candidate = ntiles x (R_T x NTAPS IMADs per lane), NTAPS chosen per candidate from
//     {4, 8, 12, 32} so short predictors do not pay for MAXORD taps;
//   * partition sums land in a heap-indexed tree (node n = 2^po + p); merging partitions is
//     tree[n] = tree[2n] + tree[2n+1]; every (order, partition) Rice parameter / bit estimate
//     is evaluated in one sweep over the nodes; orders 0..4 share one 32-lane chunk;
//   * only __syncwarp() is ever needed.
// Arithmetic and decision order are identical to k_search / k_search2 (and the reference).
#pragma once

#include "encode_kernels_v2.cuh"

namespace fb200 {

struct SearchWarpShared {
	unsigned long long tree[2 * kMaxPartitions];  // heap: node n = (1<<po) + p ; [0] unused
	uint8_t params_all[2 * kMaxPartitions];       // same indexing
	uint8_t b_params[kMaxPartitions];
	uint32_t obits[kMaxPartitionOrder + 1];
};

// Row layout of a warp's signal slice: sample i lives at word (i / R_T) * 36 + (i % R_T). A row is one
// lane's run; the 36-word row stride makes every 128-bit load of a quarter warp hit 8 distinct
// bank groups (R_T = 36: identity layout; R_T = 32: 4 pad words per row).
template <int R_T>
__device__ __forceinline__ int row_word(int i)
{
	return R_T == 32 ? i + 4 * (i >> 5) : i;
}

// |residual| sum of G consecutive outputs. xg[MAXORD + m] = output sample m, xg[0..MAXORD) = history.
// MASKED: leave out the first ord0 outputs (the warm-up samples of the block's very first group;
// order <= MAXORD <= ... so only outputs m < MAXORD can be masked).
// The body is deliberately small (G x NTAPS MACs): the callers loop over groups with a ROLLED loop so
// the hot code stays inside the instruction cache (a fully unrolled 32 x 12 run per variant did not:
// ncu showed 30 % "no_instructions" stalls).
template <int G, int MAXORD, int NTAPS, bool WIDE, bool MASKED, bool NARROW>
__device__ __forceinline__ void group_abs_sum(const int (&xg)[MAXORD + G], const int (&q)[MAXORD], int shift, int ord0, int limit,
                                              uint32_t &s32, unsigned long long &s64, bool &bad)
{
	constexpr bool narrow_acc = NARROW;
#pragma unroll
	for(int m = 0; m < G; m++) {
		uint32_t a;
		bool keep = true;
		if(MASKED && m < MAXORD) keep = m >= ord0;
		if(!WIDE) {
			int sum = 0;
#pragma unroll
			for(int j = 0; j < NTAPS; j++) sum += q[j] * xg[MAXORD + m - 1 - j];
			const int pred = sum >> shift;
			if(narrow_acc && !(MASKED && m < MAXORD)) { s32 = __sad(xg[MAXORD + m], pred, s32); continue; }
			a = __sad(xg[MAXORD + m], pred, 0u);
		}
		else {
			long long sum = 0;
#pragma unroll
			for(int j = 0; j < NTAPS; j++) sum += (long long)q[j] * (long long)xg[MAXORD + m - 1 - j];
			const long long rr = (long long)xg[MAXORD + m] - (sum >> shift);
			if(limit && (rr <= (long long)INT32_MIN || rr > (long long)INT32_MAX)) bad = true;  // lpc.c:868-884
			a = abs_u32((int)rr);
		}
		if(keep) { if(narrow_acc) s32 += a; else s64 += a; }
	}
}

// WIDEK: the stream needs 64-bit FIR accumulation (bits_per_sample > 16). For <= 16-bit streams the
// reference's precision clamp (stream_encoder.c:4591-4595) guarantees 32-bit sums, so the narrow kernel
// only keeps a compact wide fallback for the (never observed) `limit` case.
template <int R_T, int MAXORD, int NW, bool WIDEK>
__global__ void __launch_bounds__(NW * 32) k_search3(EncK P, const int32_t *__restrict__ sig, const SigMeta *__restrict__ meta,
                                                    const CandDesc *__restrict__ cdesc, SubframePlan *__restrict__ plans, int nitems)
{
	static_assert(MAXORD >= 8 && MAXORD % 4 == 0 && R_T % 4 == 0, "vector loads need multiples of 4");
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, bs = P.bs;
	const int item = blockIdx.x * NW + warp;
	if(item >= nitems) return;
	const int nrows = bs / R_T;
	const size_t xs_bytes = ((size_t)nrows * 36 * 4 + 15) / 16 * 16;
	const size_t per_warp = xs_bytes + (sizeof(SearchWarpShared) + 15) / 16 * 16;
	int32_t *xs = reinterpret_cast<int32_t *>(smem_raw + per_warp * warp);
	SearchWarpShared &S = *reinterpret_cast<SearchWarpShared *>(smem_raw + per_warp * warp + xs_bytes);

	const SigMeta M = meta[item];
	SubframePlan *plan = plans + item;
	if(M.bps == 0) {
		if(lane == 0) { plan->type = -1; plan->est_bits = 0xffffffffu; }
		return;
	}
	const int sbps = M.bps, wasted = M.wasted;
	const int32_t *g = sig + (size_t)item * P.bs_stride;
	for(int i = lane; i < bs; i += 32) xs[row_word<R_T>(i)] = g[i];
	for(int p = lane; p < kMaxPartitions; p += 32) S.b_params[p] = 0;
	__syncwarp();

	constexpr int TILE = 32 * R_T;
	constexpr int GAPV = (R_T == 32 ? 4 : 0) / 4;  // pad int4s between a row's history and its start
	const int ntiles = bs / TILE;

	constexpr int G = (R_T == 32) ? 16 : 12;   // outputs per group
	constexpr int NG = R_T / G;                // groups per run
	// loads group `g` of row `row` (+ MAXORD history samples) with 128-bit shared loads
	auto load_group = [&](int row, int g, int (&xg)[MAXORD + G]) {
		const int sv0 = (row * R_T + g * G) / 4 - MAXORD / 4;  // first sample-vector (4 samples) needed
#pragma unroll
		for(int k = 0; k < (MAXORD + G) / 4; k++) {
			const int sv = sv0 + k;
			int4 v = make_int4(0, 0, 0, 0);
			if(sv >= 0) v = *reinterpret_cast<const int4 *>(xs + (R_T == 32 ? 4 * sv + 4 * (sv >> 3) : 4 * sv));
			xg[4 * k] = v.x; xg[4 * k + 1] = v.y; xg[4 * k + 2] = v.z; xg[4 * k + 3] = v.w;
		}
	};

	// best-so-far (uniform across the warp)
	uint32_t best_bits;
	int b_type = SF_VERBATIM, b_order = 0, b_prec = 0, b_shift = 0, b_method = 0, b_po = 0, b_wide = 0;
	int b_q[MAXORD];
#pragma unroll
	for(int j = 0; j < MAXORD; j++) b_q[j] = 0;
	if(P.dis_verb && bs >= (int)kMaxFixedOrder) best_bits = 0xffffffffu;
	else best_bits = kSubframeHeaderBits + (uint32_t)wasted + (uint32_t)bs * (uint32_t)sbps;

	auto evaluate = [&](int type, int order, int precision, int shift, int wide, int limit, const int (&q)[MAXORD]) {
		int max_po = P.max_po;
		while(max_po > 0 && (bs >> max_po) <= order) max_po--;
		const int min_po = min(P.min_po, max_po);
		const int psize = bs >> max_po;
		const bool narrow = (uint32_t)(sbps + (int)kMaxExtraResidualBps) < 32u - ilog2_u32((uint32_t)psize);
		const int lpp = psize / R_T;  // lanes (runs) per partition at max_po: a power of two
		const int lpp_log = (int)ilog2_u32((uint32_t)lpp);
		const int nbase = 1 << max_po;
		for(int n = nbase + lane; n < 2 * nbase; n += 32) S.tree[n] = 0;
		__syncwarp();
		bool bad = false;
		// tap class of this candidate: the rolled group loop below runs one small body per class
		const int cls = wide ? (WIDEK ? ((MAXORD > 12 && order > 12) ? 5 : (MAXORD > 8 && order > 8) ? 4 : 3) : 6)
		                     : ((MAXORD > 12 && order > 12) ? 2 : (MAXORD > 8 && order > 8) ? 1 : order > 4 ? 0 : 7);
		constexpr int NT12 = MAXORD < 12 ? MAXORD : 12;
		uint32_t s32 = 0;
		unsigned long long s64 = 0;
#define FB200_RUN_GROUP(MK, ORD0)                                                                                            \
	do {                                                                                                                   \
		if(narrow) {                                                                                                       \
			switch(cls) {                                                                                                  \
				case 7: group_abs_sum<G, MAXORD, 4, false, MK, true>(xg, q, shift, ORD0, 0, s32, s64, bad); break;         \
				case 0: group_abs_sum<G, MAXORD, 8, false, MK, true>(xg, q, shift, ORD0, 0, s32, s64, bad); break;         \
				case 1: group_abs_sum<G, MAXORD, NT12, false, MK, true>(xg, q, shift, ORD0, 0, s32, s64, bad); break;      \
				case 2: group_abs_sum<G, MAXORD, MAXORD, false, MK, true>(xg, q, shift, ORD0, 0, s32, s64, bad); break;    \
				case 3: group_abs_sum<G, MAXORD, 8, true, MK, true>(xg, q, shift, ORD0, limit, s32, s64, bad); break;      \
				case 4: group_abs_sum<G, MAXORD, NT12, true, MK, true>(xg, q, shift, ORD0, limit, s32, s64, bad); break;   \
				default: group_abs_sum<G, MAXORD, MAXORD, true, MK, true>(xg, q, shift, ORD0, limit, s32, s64, bad); break; \
			}                                                                                                              \
		}                                                                                                                  \
		else {                                                                                                             \
			switch(cls) {                                                                                                  \
				case 7: group_abs_sum<G, MAXORD, 4, false, MK, false>(xg, q, shift, ORD0, 0, s32, s64, bad); break;        \
				case 0: group_abs_sum<G, MAXORD, 8, false, MK, false>(xg, q, shift, ORD0, 0, s32, s64, bad); break;        \
				case 1: group_abs_sum<G, MAXORD, NT12, false, MK, false>(xg, q, shift, ORD0, 0, s32, s64, bad); break;     \
				case 2: group_abs_sum<G, MAXORD, MAXORD, false, MK, false>(xg, q, shift, ORD0, 0, s32, s64, bad); break;   \
				case 3: group_abs_sum<G, MAXORD, 8, true, MK, false>(xg, q, shift, ORD0, limit, s32, s64, bad); break;     \
				case 4: group_abs_sum<G, MAXORD, NT12, true, MK, false>(xg, q, shift, ORD0, limit, s32, s64, bad); break;  \
				default: group_abs_sum<G, MAXORD, MAXORD, true, MK, false>(xg, q, shift, ORD0, limit, s32, s64, bad); break; \
			}                                                                                                              \
		}                                                                                                                  \
	} while(0)
		{   // group 0 of row 0 carries the warm-up mask (only lane 0 has ord0 != 0)
			int xg[MAXORD + G];
			load_group(lane, 0, xg);
			const int ord0 = lane == 0 ? order : 0;
			FB200_RUN_GROUP(true, ord0);
		}
#pragma unroll 1
		for(int gi = 1; gi <= ntiles * NG; gi++) {
			const int t = (gi - 1) / NG, gdone = (gi - 1) % NG;  // group (t, gdone) has just been accumulated
			if(gdone == NG - 1) {
				// a run is complete: fold it into its partition
				unsigned long long s = narrow ? (unsigned long long)s32 : s64;
				s32 = 0; s64 = 0;
			if(lpp_log <= 5) {
				for(int o = lpp >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
				if((lane & (lpp - 1)) == 0) S.tree[nbase + ((t * 32 + lane) >> lpp_log)] += s;
			}
			else {
				s = warp_sum_u64(s);
				if(lane == 0) S.tree[nbase + ((t * 32) >> lpp_log)] += s;
			}
			__syncwarp();
			}
			if(gi == ntiles * NG) break;
			{
				int xg[MAXORD + G];
				load_group((gi / NG) * 32 + lane, gi % NG, xg);
				if(MAXORD > G && gi < NG) {
					// the warm-up region (order <= MAXORD samples) reaches past the first group of the block's first run
					const int ord_g = (lane == 0 && order > gi * G) ? order - gi * G : 0;
					FB200_RUN_GROUP(true, ord_g);
				}
				else FB200_RUN_GROUP(false, 0);
			}
		}
		if(__any_sync(0xffffffffu, bad)) return;  // evaluate_lpc_subframe_ returns 0 (stream_encoder.c:4601-4609)
		if(narrow) {
			for(int n = nbase + lane; n < 2 * nbase; n += 32) S.tree[n] = (unsigned long long)(uint32_t)S.tree[n];
			__syncwarp();
		}
		// merge partitions for the lower orders (precompute_partition_info_sums_, :4837-4851)
		for(int po = max_po - 1; po >= min_po; po--) {
			for(int n = (1 << po) + lane; n < (2 << po); n += 32) S.tree[n] = S.tree[2 * n] + S.tree[2 * n + 1];
			__syncwarp();
		}
		// Rice parameter + bit estimate of every (order, partition) node (set_partitioned_rice_, :4954-5075)
		const int n_lo = 1 << min_po, n_hi = 2 << max_po;
		for(int c = n_lo >> 5; c * 32 < n_hi; c++) {
			const int n = c * 32 + lane;
			const bool valid = n >= n_lo && n < n_hi;
			unsigned long long bits = 0;
			int po = 0;
			if(valid) {
				po = (int)ilog2_u32((uint32_t)n);
				const int p = n - (1 << po);
				const uint32_t pbase = (uint32_t)(bs >> po);
				uint32_t psamp = pbase, div = 0x40000u / pbase;
				if(p == 0) { psamp -= (uint32_t)order; div = 0x40000u / psamp; }
				const unsigned long long mean = S.tree[n];
				uint32_t k;
				if(mean < 2 || (((mean - 1) * div) >> 18) == 0) k = 0;
				else k = ilog2_u64(((mean - 1) * div) >> 18) + 1;
				if(k >= (uint32_t)P.rice_limit) k = (uint32_t)P.rice_limit - 1;
				S.params_all[n] = (uint8_t)k;
				bits = count_rice_bits(k, psamp, mean);
			}
			if(c == 0) {
				// orders 0..4 live in lanes [2^po, 2^(po+1)): segmented butterfly inside aligned power-of-two blocks
				const int seg = lane ? (1 << ilog2_u32((uint32_t)lane)) : 1;
#pragma unroll
				for(int o = 1; o < 16; o <<= 1) {
					const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o);
					if(o < seg) bits += other;
				}
				if(valid && lane == seg) {
					const unsigned long long total = bits + (kEntropyTypeLen + kRiceOrderLen);
					S.obits[po] = (uint32_t)(total < 0xffffffffull ? total : 0xffffffffull);
				}
			}
			else {
				// a whole chunk belongs to one order (>= 5); orders with more than 32 partitions span several chunks
				const unsigned long long part = warp_sum_u64(bits);
				const int cpo = (int)ilog2_u32((uint32_t)(c * 32));
				const bool first = (c * 32) == (1 << cpo);
				const bool last = ((c + 1) * 32) == (2 << cpo);
				unsigned long long acc = part;
				if(!first) acc += S.tree[0];  // tree[0] is unused by the heap: running total of the current order
				if(lane == 0) {
					if(last) {
						const unsigned long long total = acc + (kEntropyTypeLen + kRiceOrderLen);
						S.obits[cpo] = (uint32_t)(total < 0xffffffffull ? total : 0xffffffffull);
					}
					else S.tree[0] = acc;
				}
			}
			__syncwarp();
		}
		uint32_t best_r = 0;
		int best_po = 0;
		for(int po = max_po; po >= min_po; po--) {
			const uint32_t bits = S.obits[po];
			if(best_r == 0 || bits < best_r) { best_r = bits; best_po = po; }
		}
		uint32_t estimate = kSubframeHeaderBits + (uint32_t)wasted;
		if(type == SF_FIXED) estimate += (uint32_t)order * (uint32_t)sbps;
		else estimate += kQlpPrecisionLen + kQlpShiftLen + (uint32_t)order * (uint32_t)(precision + sbps);
		if(best_r < 0xffffffffu - estimate) estimate += best_r;
		else estimate = 0xffffffffu;
		const bool better = (type == SF_LPC ? estimate > 0 : true) && estimate < best_bits;
		if(better) {
			uint32_t any15 = 0;
			for(int p = lane; p < (1 << best_po); p += 32) {
				const uint8_t k = S.params_all[(1 << best_po) + p];
				S.b_params[p] = k;
				any15 |= (k >= kRiceEscape) ? 1u : 0u;
			}
			any15 = warp_or(any15);
			best_bits = estimate;
			b_type = type; b_order = order; b_prec = precision; b_shift = shift; b_method = any15 ? 1 : 0; b_po = best_po; b_wide = wide;
#pragma unroll
			for(int j = 0; j < MAXORD; j++) b_q[j] = (type == SF_LPC) ? q[j] : 0;
		}
		__syncwarp();
	};

	if(bs > (int)kMaxFixedOrder) {
		// fixed-predictor scan (fixed.c:222-290) + constant detection over all tiles
		unsigned long long te[5] = {0, 0, 0, 0, 0};
		uint32_t diff = 0;
		const int32_t x0 = xs[0];
		// The order-k error is the k-th finite difference; in wrapping 32-bit arithmetic it equals the
		// reference's value whenever that fits an int32: |e4| <= 16 * 2^(sbps-1), i.e. sbps <= 27 (the
		// engine's scope is sbps <= 25). |e| sums: a lane adds bs/32 values below 2^(sbps+3) per order, so
		// 32-bit lane totals are exact when sbps + 3 + log2(bs/32) <= 32; otherwise they are flushed to
		// 64 bits every 4 samples (4 * 2^28 < 2^32).
		const bool lane_total_fits = (uint32_t)(sbps + 3) + ilog2_u32((uint32_t)(2 * (bs / 32) - 1)) <= 32u;
		uint32_t t32[5] = {0, 0, 0, 0, 0};
		for(int t = 0; t < ntiles; t++) {
			const int row = t * 32 + lane;
			int xw[4 + R_T];
			{
				const int4 *pv = reinterpret_cast<const int4 *>(xs + row * 36);
				if(row == 0) { xw[0] = xw[1] = xw[2] = xw[3] = 0; }
				else { const int4 v = pv[-1 - GAPV]; xw[0] = v.x; xw[1] = v.y; xw[2] = v.z; xw[3] = v.w; }
#pragma unroll
				for(int k = 0; k < R_T / 4; k++) {
					const int4 v = pv[k];
					xw[4 + 4 * k] = v.x; xw[5 + 4 * k] = v.y; xw[6 + 4 * k] = v.z; xw[7 + 4 * k] = v.w;
				}
			}
			// difference pyramid: e1[k] belongs to sample k-3, e2[k] to k-2, e3[k] to k-1, e4 to m
			int e1[R_T + 3], e2[R_T + 2], e3[R_T + 1];
#pragma unroll
			for(int k = 0; k < R_T + 3; k++) e1[k] = xw[k + 1] - xw[k];
#pragma unroll
			for(int k = 0; k < R_T + 2; k++) e2[k] = e1[k + 1] - e1[k];
#pragma unroll
			for(int k = 0; k < R_T + 1; k++) e3[k] = e2[k + 1] - e2[k];
#pragma unroll
			for(int m = 0; m < R_T; m++) {
				diff |= (uint32_t)(xw[4 + m] ^ x0);
				const bool counted = m >= (int)kMaxFixedOrder || row != 0;  // fixed.c:222-290 starts at sample 4
				if(counted) {
					t32[0] = __sad(xw[4 + m], 0, t32[0]);
					t32[1] = __sad(e1[m + 3], 0, t32[1]);
					t32[2] = __sad(e2[m + 2], 0, t32[2]);
					t32[3] = __sad(e3[m + 1], 0, t32[3]);
					t32[4] = __sad(e3[m + 1] - e3[m], 0, t32[4]);
				}
				if((m & 3) == 3 && !lane_total_fits) {
#pragma unroll
					for(int k = 0; k < 5; k++) { te[k] += t32[k]; t32[k] = 0; }
				}
			}
		}
#pragma unroll
		for(int k = 0; k < 5; k++) te[k] += t32[k];
		uint32_t eq = diff == 0 ? 1u : 0u;
#pragma unroll
		for(int k = 0; k < 5; k++) te[k] = warp_sum_u64(te[k]);
		eq = warp_and(eq);
		int guess;
		{
			const unsigned long long m34 = te[3] < te[4] ? te[3] : te[4], m234 = te[2] < m34 ? te[2] : m34, m1234 = te[1] < m234 ? te[1] : m234;
			if(te[0] <= m1234) guess = 0;
			else if(te[1] <= m234) guess = 1;
			else if(te[2] <= m34) guess = 2;
			else if(te[3] <= te[4]) guess = 3;
			else guess = 4;
		}
		float rbps[5];
		{
			const double n = (double)(uint32_t)(bs - (int)kMaxFixedOrder);
#pragma unroll
			for(int k = 0; k < 5; k++)
				rbps[k] = (float)((te[k] > 0) ? fb_log(M_LN2 * (double)te[k] / n) / M_LN2 : 0.0);
		}
		const bool is_constant = !P.dis_const && rbps[1] == 0.0f && eq;
		if(is_constant) {
			const uint32_t cbits = kSubframeHeaderBits + (uint32_t)wasted + (uint32_t)sbps;
			if(cbits < best_bits) { best_bits = cbits; b_type = SF_CONSTANT; }
		}
		else {
			if(!P.dis_fixed || (P.max_order == 0 && best_bits == 0xffffffffu)) {
				int lo, hi;
				if(P.exhaustive) { lo = 0; hi = (int)kMaxFixedOrder; }
				else lo = hi = guess;
				if(hi >= bs) hi = bs - 1;
				for(int fo = lo; fo <= hi; fo++) {
					if(rbps[fo] >= (float)sbps) continue;
					int q[MAXORD];
#pragma unroll
					for(int j = 0; j < MAXORD; j++) q[j] = fixed_tap(fo, j);
					evaluate(SF_FIXED, fo, 0, 0, 0, 0, q);
				}
			}
			if(P.max_order > 0) {
				const CandDesc *cd = cdesc + (size_t)item * P.nslots;
				for(int c = 0; c < P.nslots; c++) {
					const CandDesc *D = cd + c;
					if(!D->valid) continue;
					int q[MAXORD];
#pragma unroll
					for(int j = 0; j < MAXORD; j++) q[j] = __ldg(&D->qlp[j]);
					evaluate(SF_LPC, D->order, D->precision, D->shift, D->wide, D->limit, q);
				}
			}
		}
	}
	if(best_bits == 0xffffffffu) {
		b_type = SF_VERBATIM;
		best_bits = kSubframeHeaderBits + (uint32_t)wasted + (uint32_t)bs * (uint32_t)sbps;
	}
	if(lane == 0) {
		plan->type = b_type; plan->order = b_order; plan->wasted = wasted; plan->bps = sbps;
		plan->precision = b_prec; plan->shift = b_shift; plan->method = b_method; plan->porder = b_po;
		plan->est_bits = best_bits; plan->wide = b_wide;
#pragma unroll
		for(int j = 0; j < FB200_MAX_LPC_ORDER; j++) plan->qlp[j] = (j < MAXORD) ? b_q[j < MAXORD ? j : 0] : 0;
	}
	for(int p = lane; p < kMaxPartitions; p += 32) plan->params[p] = S.b_params[p];
}

}  // namespace fb200
