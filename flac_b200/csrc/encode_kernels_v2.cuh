// encode_kernels_v2.cuh -- the fast-path kernels for "regular" blocksizes (multiple of the CTA
// width, at most R_T samples per thread): register-resident FIR windows instead of
// shared-memory tap loops, vectorised autocorrelation loads. Arithmetic is identical to the
// general kernels in encode_kernels.cuh (which remain the path for every other blocksize).
#pragma once

#include "encode_kernels.cuh"

namespace fb200 {

// ================================================================ k_autoc2
// Same chain-per-thread scheme as k_autoc, but every thread fetches its samples with aligned
// 128-bit loads (one L1 wavefront per lane per 4 samples instead of per sample) in bodies of
// U = lcm(LAGS,4) samples, so the rotating-history indices stay compile-time.
// SPLIT threads share one chain: thread `part` owns lags [part*NACC, part*NACC+NACC) and therefore needs a
// history of HIST = SPLIT*NACC samples. Every partial chain still adds its terms in ascending sample order,
// so the result is bit-identical to the unsplit kernel; the split only buys parallelism (the kernel is
// latency-bound: ncu shows the FP64 pipe ~1/3 busy).
template <int NACC, int SPLIT, int U>
__global__ void __launch_bounds__(128) k_autoc2(EncK P, const int32_t *__restrict__ sig, const SigMeta *__restrict__ meta,
                                               const float *__restrict__ windows, const DevSection *__restrict__ secs,
                                               double *__restrict__ autoc, int nitems)
{
	constexpr int HIST = NACC * SPLIT;
	static_assert(U % HIST == 0 && U % 4 == 0, "U must be a common multiple of the history depth and 4");
	const int gid = blockIdx.x * blockDim.x + threadIdx.x;
	if(gid >= nitems * P.nsec * SPLIT) return;
	const int part = gid / (nitems * P.nsec);
	const int rem = gid - part * (nitems * P.nsec);
	const int sec = rem / nitems, item = rem - sec * nitems;
	if(meta[item].bps == 0) return;
	const DevSection S = secs[sec];
	const int32_t *x = sig + (size_t)item * P.bs_stride;
	const float *w = windows + S.win_off;
	const int shift = S.partial ? S.data_shift : 0;
	const int a0 = shift & ~3;                      // aligned start; samples before `shift` get weight 0
	const int nvalid = S.partial ? 2 * S.part_size : S.data_len;  // d[i] == 0 beyond (lpc.c:90-91)
	const int span = shift - a0 + S.data_len;
	const int wsecond = P.bs - 2 * S.part_size;     // window index offset of the falling half
	const int l0 = part * NACC;

	double acc[NACC], h[HIST];
#pragma unroll
	for(int l = 0; l < NACC; l++) acc[l] = 0.0;
#pragma unroll
	for(int l = 0; l < HIST; l++) h[l] = 0.0;

	int4 v[U / 4];
#pragma unroll
	for(int q = 0; q < U / 4; q++) v[q] = __ldg(reinterpret_cast<const int4 *>(x + a0) + q);
	for(int base = 0; base < span; base += U) {
		// software prefetch of the next body (the sig allocation has slack past the last block)
		int4 vn[U / 4];
#pragma unroll
		for(int q = 0; q < U / 4; q++) vn[q] = __ldg(reinterpret_cast<const int4 *>(x + a0 + base + U) + q);
#pragma unroll
		for(int u = 0; u < U; u++) {
			const int i = a0 + base + u - shift;  // index inside the section
			const int xv = (u & 3) == 0 ? v[u >> 2].x : (u & 3) == 1 ? v[u >> 2].y : (u & 3) == 2 ? v[u >> 2].z : v[u >> 2].w;
			float d = 0.0f;
			if(i >= 0 && i < nvalid) {
				const int wi = (!S.partial || i < S.part_size) ? i : wsecond + i;
				d = __fmul_rn((float)xv, __ldg(w + wi));
			}
			const double dv = (double)d;
			const int su = (HIST - (u % HIST)) % HIST;  // slot of the newest sample; slot (su+l)%HIST holds d[i-l]
			h[su] = dv;
			if(SPLIT == 1) {
#pragma unroll
				for(int l = 0; l < NACC; l++) acc[l] = fma(dv, h[(su + l) % HIST], acc[l]);
			}
			else if(part == 0) {
#pragma unroll
				for(int l = 0; l < NACC; l++) acc[l] = fma(dv, h[(su + l) % HIST], acc[l]);
			}
			else {
#pragma unroll
				for(int l = 0; l < NACC; l++) acc[l] = fma(dv, h[(su + NACC + l) % HIST], acc[l]);
			}
		}
#pragma unroll
		for(int q = 0; q < U / 4; q++) v[q] = vn[q];
	}
	double *out = autoc + ((size_t)sec * nitems + item) * P.lag_stride;
#pragma unroll
	for(int l = 0; l < NACC; l++)
		if(l0 + l < P.lag_stride) out[l0 + l] = acc[l];
}

// ================================================================ register-window residual
// xr[MAXORD + m] = sample (base + m), xr[0..MAXORD) = the MAXORD samples before the run.
// q[] is zero beyond the predictor order. Returns r[m] for the whole run (callers mask
// positions < order and >= run length).
template <int R_T, int MAXORD>
__device__ __forceinline__ void run_residual_narrow(const int (&xr)[MAXORD + R_T], const int (&q)[MAXORD], int shift, int (&r)[R_T])
{
#pragma unroll
	for(int m = 0; m < R_T; m++) {
		int sum = 0;
#pragma unroll
		for(int j = 0; j < MAXORD; j++) sum += q[j] * xr[MAXORD + m - 1 - j];
		r[m] = xr[MAXORD + m] - (sum >> shift);
	}
}

template <int R_T, int MAXORD>
__device__ __forceinline__ bool run_residual_wide(const int (&xr)[MAXORD + R_T], const int (&q)[MAXORD], int shift, int (&r)[R_T], int limit)
{
	bool bad = false;
#pragma unroll
	for(int m = 0; m < R_T; m++) {
		long long sum = 0;
#pragma unroll
		for(int j = 0; j < MAXORD; j++) sum += (long long)q[j] * (long long)xr[MAXORD + m - 1 - j];
		const long long rr = (long long)xr[MAXORD + m] - (sum >> shift);
		if(limit && (rr <= (long long)INT32_MIN || rr > (long long)INT32_MAX)) bad = true;
		r[m] = (int)rr;
	}
	return bad;
}

// fixed predictors as FIR taps (fixed.c:470-530): r = x[i] - sum_j c[j] x[i-1-j]
__device__ __forceinline__ int fixed_tap(int order, int j)
{
	// order 1: {1}; 2: {2,-1}; 3: {3,-3,1}; 4: {4,-6,4,-1}
	const int tab[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
	return j < 4 ? tab[order][j] : 0;
}

template <int R_T, int MAXORD>
__device__ __forceinline__ void load_run(const int32_t *__restrict__ xs, int base, int R, int (&xr)[MAXORD + R_T])
{
#pragma unroll
	for(int hh = 0; hh < MAXORD; hh++) {
		const int idx = base - MAXORD + hh;
		xr[hh] = idx >= 0 ? xs[skew(idx)] : 0;
	}
#pragma unroll
	for(int m = 0; m < R_T; m++) xr[MAXORD + m] = (m < R) ? xs[skew(base + m)] : 0;
}

// ================================================================ k_search2

struct SearchShared2 {
	unsigned long long sums[kMaxPartitions];
	unsigned long long te[5];
	uint32_t obits[kMaxPartitionOrder + 1];
	uint8_t params_all[2 * kMaxPartitions];
	int fail;
	int all_equal;
	uint32_t best_bits;
	int b_type, b_order, b_prec, b_shift, b_method, b_po, b_wide;
	int b_qlp[FB200_MAX_LPC_ORDER];
	uint8_t b_params[kMaxPartitions];
};

// R_EXACT: the run length equals R_T (compile time); otherwise R = bs/NT <= R_T at run time.
template <int R_T, int MAXORD, bool R_EXACT>
__global__ void __launch_bounds__(128) k_search2(EncK P, const int32_t *__restrict__ sig, const SigMeta *__restrict__ meta,
                                                const CandDesc *__restrict__ cdesc, SubframePlan *__restrict__ plans)
{
	constexpr int NT = 128, LOG_NT = 7;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	__shared__ SearchShared2 S;
	int32_t *xs = reinterpret_cast<int32_t *>(smem_raw);
	const int item = blockIdx.x, tid = threadIdx.x, bs = P.bs;
	const int warp = tid >> 5, lane = tid & 31;
	const int R = R_EXACT ? R_T : bs / NT;

	const SigMeta M = meta[item];
	SubframePlan *plan = plans + item;
	if(M.bps == 0) {
		if(tid == 0) { plan->type = -1; plan->est_bits = 0xffffffffu; }
		return;
	}
	const int sbps = M.bps, wasted = M.wasted;
	const int32_t *g = sig + (size_t)item * P.bs_stride;
	for(int i = tid; i < bs; i += NT) xs[skew(i)] = g[i];
	if(tid == 0) {
		S.fail = 0; S.all_equal = 1;
		S.b_type = SF_VERBATIM; S.b_order = 0; S.b_prec = 0; S.b_shift = 0; S.b_method = 0; S.b_po = 0; S.b_wide = 0;
		if(P.dis_verb && bs >= (int)kMaxFixedOrder) S.best_bits = 0xffffffffu;
		else S.best_bits = kSubframeHeaderBits + (uint32_t)wasted + (uint32_t)bs * (uint32_t)sbps;
	}
	if(tid < 5) S.te[tid] = 0;
	if(tid < FB200_MAX_LPC_ORDER) S.b_qlp[tid] = 0;
	__syncthreads();

	const int base = tid * R;
	int xr[MAXORD + R_T];
	load_run<R_T, MAXORD>(xs, base, R, xr);

	// ---- candidate evaluation (find_best_partition_order_ et al., stream_encoder.c:4701-5075)
	auto evaluate = [&](int type, int order, int precision, int shift, int wide, int limit, const int (&q)[MAXORD]) {
		int r[R_T];
		bool bad = false;
		if(!wide) run_residual_narrow<R_T, MAXORD>(xr, q, shift, r);
		else bad = run_residual_wide<R_T, MAXORD>(xr, q, shift, r, limit);

		int max_po = P.max_po;
		while(max_po > 0 && (bs >> max_po) <= order) max_po--;
		const int min_po = min(P.min_po, max_po);
		const int psize = bs >> max_po;
		const bool narrow = (uint32_t)(sbps + (int)kMaxExtraResidualBps) < 32u - ilog2_u32((uint32_t)psize);
		const int ltpp = LOG_NT - max_po;  // log2(threads per partition)
		unsigned long long s = 0;
		if(narrow) {
			uint32_t s32 = 0;
#pragma unroll
			for(int m = 0; m < R_T; m++)
				if((R_EXACT || m < R) && base + m >= order) s32 += abs_u32(r[m]);
			s = s32;
		}
		else {
#pragma unroll
			for(int m = 0; m < R_T; m++)
				if((R_EXACT || m < R) && base + m >= order) s += abs_u32(r[m]);
		}
		if(ltpp <= 5) {
			for(int o = (1 << ltpp) >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
			if((tid & ((1 << ltpp) - 1)) == 0) S.sums[tid >> ltpp] = narrow ? (unsigned long long)(uint32_t)s : s;
		}
		else {
			// partitions wider than a warp (max_po < 2): combine warp sums through shared memory
			if(tid < 2) S.sums[tid] = 0;
			__syncthreads();
			s = warp_sum_u64(s);
			if(lane == 0) atomicAdd(&S.sums[tid >> ltpp], s);
			__syncthreads();
			if(narrow && tid < (1 << max_po)) S.sums[tid] = (unsigned long long)(uint32_t)S.sums[tid];
		}
		if(bad) S.fail = 1;
		__syncthreads();  // (1) sums + fail visible
		if(S.fail) {
			__syncthreads();
			if(tid == 0) S.fail = 0;
			__syncthreads();
			return;
		}
		// each warp takes partition orders max_po - warp, max_po - warp - 4, ...
		for(int po = max_po - warp; po >= min_po; po -= NT / 32) {
			const int nparts = 1 << po, d = max_po - po;
			const uint32_t pbase = (uint32_t)(bs >> po);
			const uint32_t div_base = 0x40000u / pbase;
			unsigned long long local = 0;
			for(int p = lane; p < nparts; p += 32) {
				unsigned long long mean = 0;
				for(int e = 0; e < (1 << d); e++) mean += S.sums[(p << d) + e];
				uint32_t psamp = pbase, div = div_base;
				if(p == 0) { psamp -= (uint32_t)order; div = 0x40000u / psamp; }
				uint32_t k;
				if(mean < 2 || (((mean - 1) * div) >> 18) == 0) k = 0;
				else k = ilog2_u64(((mean - 1) * div) >> 18) + 1;
				if(k >= (uint32_t)P.rice_limit) k = (uint32_t)P.rice_limit - 1;
				S.params_all[(1 << po) - 1 + p] = (uint8_t)k;
				local += count_rice_bits(k, psamp, mean);
			}
			const unsigned long long total = warp_sum_u64(local) + (kEntropyTypeLen + kRiceOrderLen);
			if(lane == 0) S.obits[po] = (uint32_t)(total < 0xffffffffull ? total : 0xffffffffull);
		}
		__syncthreads();  // (2) per-order results visible
		if(warp == 0) {
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
			const bool better = (type == SF_LPC ? estimate > 0 : true) && estimate < S.best_bits;
			__syncwarp();
			if(better) {
				const int nparts = 1 << best_po;
				uint32_t any15 = 0;
				for(int p = lane; p < nparts; p += 32) {
					const uint8_t k = S.params_all[(1 << best_po) - 1 + p];
					S.b_params[p] = k;
					any15 |= (k >= kRiceEscape) ? 1u : 0u;
				}
				any15 = warp_or(any15);
#pragma unroll
				for(int j = 0; j < MAXORD; j++)
					if(lane == 0) S.b_qlp[j] = (type == SF_LPC) ? q[j] : 0;
				if(lane == 0) {
					S.best_bits = estimate;
					S.b_type = type; S.b_order = order; S.b_prec = precision; S.b_shift = shift;
					S.b_method = any15 ? 1 : 0;
					S.b_po = best_po; S.b_wide = wide;
				}
			}
		}
		// no barrier here: the next candidate only rewrites S.sums before its barrier (1), and
		// warp 0 reaches that barrier after it finished reading obits/params_all.
	};

	if(bs > (int)kMaxFixedOrder) {
		// ---- fixed-predictor scan (fixed.c:222-290) straight from the register window
		unsigned long long te[5] = {0, 0, 0, 0, 0};
#pragma unroll
		for(int m = 0; m < R_T; m++) {
			if((R_EXACT || m < R) && base + m >= (int)kMaxFixedOrder) {
				const long long d0 = xr[MAXORD + m], d1 = xr[MAXORD + m - 1], d2 = xr[MAXORD + m - 2], d3 = xr[MAXORD + m - 3], d4 = xr[MAXORD + m - 4];
				const long long e1 = d0 - d1, e2 = d0 - 2 * d1 + d2, e3 = d0 - 3 * d1 + 3 * d2 - d3, e4 = d0 - 4 * d1 + 6 * d2 - 4 * d3 + d4;
				te[0] += (unsigned long long)(d0 < 0 ? -d0 : d0);
				te[1] += (unsigned long long)(e1 < 0 ? -e1 : e1);
				te[2] += (unsigned long long)(e2 < 0 ? -e2 : e2);
				te[3] += (unsigned long long)(e3 < 0 ? -e3 : e3);
				te[4] += (unsigned long long)(e4 < 0 ? -e4 : e4);
			}
		}
#pragma unroll
		for(int k = 0; k < 5; k++) {
			te[k] = warp_sum_u64(te[k]);
			if(lane == 0) atomicAdd(&S.te[k], te[k]);
		}
		__syncthreads();
		const unsigned long long t0 = S.te[0], t1 = S.te[1], t2 = S.te[2], t3 = S.te[3], t4 = S.te[4];
		int guess;
		{
			const unsigned long long m34 = t3 < t4 ? t3 : t4, m234 = t2 < m34 ? t2 : m34, m1234 = t1 < m234 ? t1 : m234;
			if(t0 <= m1234) guess = 0;
			else if(t1 <= m234) guess = 1;
			else if(t2 <= m34) guess = 2;
			else if(t3 <= t4) guess = 3;
			else guess = 4;
		}
		float rbps[5];
		{
			const double n = (double)(uint32_t)(bs - (int)kMaxFixedOrder);
			const unsigned long long tt[5] = {t0, t1, t2, t3, t4};
#pragma unroll
			for(int k = 0; k < 5; k++)
				rbps[k] = (float)((tt[k] > 0) ? fb_log(M_LN2 * (double)tt[k] / n) / M_LN2 : 0.0);
		}
		bool is_constant = false;
		if(!P.dis_const && rbps[1] == 0.0f) {
			uint32_t eq = 1;
			const int32_t x0 = xs[0];
#pragma unroll
			for(int m = 0; m < R_T; m++)
				if(R_EXACT || m < R) eq &= (xr[MAXORD + m] == x0) ? 1u : 0u;
			eq = warp_and(eq);
			if(lane == 0 && !eq) S.all_equal = 0;
			__syncthreads();
			is_constant = S.all_equal != 0;
		}
		if(is_constant) {
			if(tid == 0) {
				const uint32_t cbits = kSubframeHeaderBits + (uint32_t)wasted + (uint32_t)sbps;
				if(cbits < S.best_bits) { S.best_bits = cbits; S.b_type = SF_CONSTANT; }
			}
		}
		else {
			if(!P.dis_fixed || (P.max_order == 0 && S.best_bits == 0xffffffffu)) {
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
	__syncthreads();
	if(S.best_bits == 0xffffffffu) {
		if(tid == 0) {
			S.b_type = SF_VERBATIM;
			S.best_bits = kSubframeHeaderBits + (uint32_t)wasted + (uint32_t)bs * (uint32_t)sbps;
		}
		__syncthreads();
	}
	if(tid == 0) {
		plan->type = S.b_type; plan->order = S.b_order; plan->wasted = wasted; plan->bps = sbps;
		plan->precision = S.b_prec; plan->shift = S.b_shift; plan->method = S.b_method; plan->porder = S.b_po;
		plan->est_bits = S.best_bits; plan->wide = S.b_wide;
	}
	if(tid < FB200_MAX_LPC_ORDER) plan->qlp[tid] = S.b_qlp[tid];
	for(int p = tid; p < kMaxPartitions; p += NT) plan->params[p] = S.b_params[p];
}

// ================================================================ k_emit2
template <int NT>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *s_warp /*[NT/32+1]*/, uint32_t *total)
{
	constexpr int NW = NT / 32;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for(int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
		if(lane >= o) inc += t;
	}
	if(lane == 31) s_warp[warp] = inc;
	__syncthreads();
	if(warp == 0) {
		const uint32_t w = (lane < NW) ? s_warp[lane] : 0;
		uint32_t winc = w;
#pragma unroll
		for(int o = 1; o < NW; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
			if(lane >= o) winc += t;
		}
		if(lane < NW) s_warp[lane] = winc - w;
		if(lane == NW - 1) s_warp[NW] = winc;
	}
	__syncthreads();
	const uint32_t res = s_warp[warp] + inc - v;
	*total = s_warp[NW];
	return res;  // callers place a barrier before the next scan reuses s_warp (k_emit2: the end-of-channel barrier)
}

template <int NT, int R_T, int MAXORD, bool R_EXACT>
// 256-thread CTAs: three per SM (shared memory allows it), which caps the kernel at 80 registers per thread
__global__ void __launch_bounds__(NT, NT == 256 ? 3 : 1) k_emit2(EncK P, const int32_t *__restrict__ sig, const int *__restrict__ blkflags,
                                             const SubframePlan *__restrict__ plans, uint8_t *__restrict__ slots,
                                             uint32_t *__restrict__ frame_bytes, uint32_t *__restrict__ chan_assign_out)
{
	constexpr int NW = NT / 32;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int blk = blockIdx.x, tid = threadIdx.x, bs = P.bs;
	const int xcap = skew(P.bs_stride) + 1;
	int32_t *xs = reinterpret_cast<int32_t *>(smem_raw);
	uint32_t *words = reinterpret_cast<uint32_t *>(xs + xcap);
	__shared__ uint16_t s_crctab[256];
	__shared__ uint32_t s_warp[NW + 1];
	__shared__ uint32_t s_crc[NT];
	__shared__ uint32_t s_mlev[8];
	__shared__ int s_ca;
	const int R = R_EXACT ? R_T : bs / NT;

	for(int i = tid; i < P.slot_words; i += NT) words[i] = 0;
	for(int e = tid; e < 256; e += NT) {
		uint32_t c = (uint32_t)e << 8;
#pragma unroll
		for(int j = 0; j < 8; j++) c = (c & 0x8000u) ? ((c << 1) ^ 0x8005u) : (c << 1);
		s_crctab[e] = (uint16_t)c;
	}
	const SubframePlan *bp = plans + (size_t)blk * P.nsig;
	if(tid == 0) {
		int ca = 0;
		const int flags = blkflags[blk];
		if(P.channels == 2 && P.do_ms) {
			if(P.loose_ms) ca = (flags & 2) ? 3 : 0;
			else {
				const uint32_t b0 = bp[0].est_bits + bp[1].est_bits, b1 = bp[0].est_bits + bp[3].est_bits;
				const uint32_t b2 = bp[1].est_bits + bp[3].est_bits, b3 = bp[2].est_bits + bp[3].est_bits;
				uint32_t mn = b0;
				if(b1 < mn) { mn = b1; ca = 1; }
				if(b2 < mn) { mn = b2; ca = 2; }
				if(b3 < mn) { mn = b3; ca = 3; }
			}
		}
		s_ca = ca;
	}
	__syncthreads();
	const int ca = s_ca;
	const uint32_t frame_number = P.first_frame + (uint32_t)blk;

	// ---- frame header (identical to k_emit)
	uint32_t bs_code, bs_hint = 0, sr_code, sr_hint = 0;
	switch(bs) {
		case 192: bs_code = 1; break; case 576: bs_code = 2; break; case 1152: bs_code = 3; break;
		case 2304: bs_code = 4; break; case 4608: bs_code = 5; break; case 256: bs_code = 8; break;
		case 512: bs_code = 9; break; case 1024: bs_code = 10; break; case 2048: bs_code = 11; break;
		case 4096: bs_code = 12; break; case 8192: bs_code = 13; break; case 16384: bs_code = 14; break;
		case 32768: bs_code = 15; break;
		default: bs_code = bs_hint = (bs <= 0x100) ? 6 : 7; break;
	}
	switch(P.sample_rate) {
		case 88200: sr_code = 1; break; case 176400: sr_code = 2; break; case 192000: sr_code = 3; break;
		case 8000: sr_code = 4; break; case 16000: sr_code = 5; break; case 22050: sr_code = 6; break;
		case 24000: sr_code = 7; break; case 32000: sr_code = 8; break; case 44100: sr_code = 9; break;
		case 48000: sr_code = 10; break; case 96000: sr_code = 11; break;
		default:
			if(P.sample_rate <= 255000 && P.sample_rate % 1000 == 0) sr_code = sr_hint = 12;
			else if(P.sample_rate <= 655350 && P.sample_rate % 10 == 0) sr_code = sr_hint = 14;
			else if(P.sample_rate <= 0xffff) sr_code = sr_hint = 13;
			else sr_code = 0;
			break;
	}
	uint32_t utf8_len;
	if(frame_number < 0x80) utf8_len = 1;
	else if(frame_number < 0x800) utf8_len = 2;
	else if(frame_number < 0x10000) utf8_len = 3;
	else if(frame_number < 0x200000) utf8_len = 4;
	else if(frame_number < 0x4000000) utf8_len = 5;
	else utf8_len = 6;
	const uint32_t header_bits = 32 + 8 * utf8_len + (bs_hint ? (bs_hint == 6 ? 8 : 16) : 0) + (sr_hint ? (sr_hint == 12 ? 8 : 16) : 0) + 8;
	if(tid == 0) {
		BitPut bw;
		bw.init(words, 0);
		uint32_t ca_code;
		switch(ca) { case 0: ca_code = (uint32_t)P.channels - 1; break; case 1: ca_code = 8; break; case 2: ca_code = 9; break; default: ca_code = 10; break; }
		uint32_t bps_code;
		switch(P.bps) { case 8: bps_code = 1; break; case 12: bps_code = 2; break; case 16: bps_code = 4; break; case 20: bps_code = 5; break; case 24: bps_code = 6; break; case 32: bps_code = 7; break; default: bps_code = 0; break; }
		bw.put(0x3ffe, 14); bw.put(0, 1); bw.put(0, 1);
		bw.put(bs_code, 4); bw.put(sr_code, 4); bw.put(ca_code, 4); bw.put(bps_code, 3); bw.put(0, 1);
		const uint32_t v = frame_number;
		switch(utf8_len) {
			case 1: bw.put(v, 8); break;
			case 2: bw.put(0xC0 | (v >> 6), 8); break;
			case 3: bw.put(0xE0 | (v >> 12), 8); break;
			case 4: bw.put(0xF0 | (v >> 18), 8); break;
			case 5: bw.put(0xF8 | (v >> 24), 8); break;
			default: bw.put(0xFC | (v >> 30), 8); break;
		}
		for(int k = (int)utf8_len - 2; k >= 0; k--) bw.put(0x80 | ((v >> (6 * k)) & 0x3F), 8);
		if(bs_hint) bw.put((uint32_t)bs - 1, bs_hint == 6 ? 8 : 16);
		if(sr_hint == 12) bw.put((uint32_t)P.sample_rate / 1000, 8);
		else if(sr_hint == 13) bw.put((uint32_t)P.sample_rate, 16);
		else if(sr_hint == 14) bw.put((uint32_t)P.sample_rate / 10, 16);
		bw.finish();
		const uint32_t nb = (header_bits - 8) >> 3;
		uint32_t crc = 0;
		for(uint32_t b = 0; b < nb; b++) {
			crc ^= (words[b >> 2] >> (24 - 8 * (b & 3))) & 0xffu;
			for(int j = 0; j < 8; j++) crc = (crc & 0x80u) ? ((crc << 1) ^ 0x07u) & 0xffu : (crc << 1) & 0xffu;
		}
		bw.init(words, header_bits - 8);
		bw.put(crc, 8);
		bw.finish();
	}
	__syncthreads();

	uint32_t bitpos = header_bits;
	for(int c = 0; c < P.channels; c++) {
		int sidx = c;
		if(P.channels == 2) {
			if(c == 0) sidx = (ca == 0 || ca == 1) ? 0 : (ca == 2 ? 3 : 2);
			else sidx = (ca == 0 || ca == 2) ? 1 : 3;
		}
		const SubframePlan *pl = bp + sidx;
		const int type = pl->type, order = pl->order, wasted = pl->wasted, sbps = pl->bps;
		const int32_t *g = sig + ((size_t)blk * P.nsig + sidx) * P.bs_stride;
		const uint32_t hdr_bits = kSubframeHeaderBits + (uint32_t)wasted;

		// Subframe header (frame-format: type byte, wasted-bits unary, warm-up samples, precision/shift, coefficients,
		// entropy method + partition order): every field's bit position is known, so each goes out from its own thread.
		{
			const uint32_t warm0 = bitpos + hdr_bits;                              // first warm-up sample
			const uint32_t after_warm = warm0 + (uint32_t)order * (uint32_t)sbps;  // precision (LPC) or entropy method (fixed)
			const bool predicted = type == SF_FIXED || type == SF_LPC;
			BitPut bw;
			if(tid == 0) {
				uint32_t tb;
				switch(type) {
					case SF_CONSTANT: tb = 0x00; break;
					case SF_VERBATIM: tb = 0x02; break;
					case SF_FIXED: tb = 0x10 | ((uint32_t)order << 1); break;
					default: tb = 0x40 | ((uint32_t)(order - 1) << 1); break;
				}
				bw.init(words, bitpos);
				bw.put(tb | (wasted ? 1u : 0u), 8);
				if(wasted) { bw.skip((uint32_t)wasted - 1); bw.put(1, 1); }
				if(type == SF_CONSTANT) bw.put(mask_bits(g[0], (uint32_t)sbps), (uint32_t)sbps);
				bw.finish();
			}
			else if(predicted && tid >= 32 && tid < 32 + order) {
				const int i = tid - 32;
				bw.init(words, warm0 + (uint32_t)i * (uint32_t)sbps);
				bw.put(mask_bits(g[i], (uint32_t)sbps), (uint32_t)sbps);
				bw.finish();
			}
			else if(type == SF_LPC && tid >= 64 && tid < 64 + order) {
				const int i = tid - 64;
				const uint32_t prec = (uint32_t)pl->precision;
				bw.init(words, after_warm + kQlpPrecisionLen + kQlpShiftLen + (uint32_t)i * prec);
				bw.put(mask_bits(pl->qlp[i], prec), prec);
				bw.finish();
			}
			else if(predicted && tid == 96) {
				bw.init(words, after_warm);
				if(type == SF_LPC) {
					bw.put((uint32_t)pl->precision - 1, kQlpPrecisionLen);
					bw.put(mask_bits(pl->shift, kQlpShiftLen), kQlpShiftLen);
					bw.finish();
					bw.init(words, after_warm + kQlpPrecisionLen + kQlpShiftLen + (uint32_t)order * (uint32_t)pl->precision);
				}
				bw.put((uint32_t)pl->method, kEntropyTypeLen);
				bw.put((uint32_t)pl->porder, kRiceOrderLen);
				bw.finish();
			}
		}
		bitpos += hdr_bits;

		if(type == SF_CONSTANT) {
			bitpos += (uint32_t)sbps;
		}
		else if(type == SF_VERBATIM) {
			const int i0 = tid * R, i1 = min(bs, i0 + R);
			if(i0 < i1) {
				BitPut bw;
				bw.init(words, bitpos + (uint32_t)i0 * (uint32_t)sbps);
				for(int i = i0; i < i1; i++) bw.put(mask_bits(g[i], (uint32_t)sbps), (uint32_t)sbps);
				bw.finish();
			}
			bitpos += (uint32_t)bs * (uint32_t)sbps;
		}
		else {
			bitpos += (uint32_t)order * (uint32_t)sbps;
			if(type == SF_LPC) bitpos += kQlpPrecisionLen + kQlpShiftLen + (uint32_t)order * (uint32_t)pl->precision;
			bitpos += kEntropyTypeLen + kRiceOrderLen;

			for(int i = tid; i < bs; i += NT) xs[skew(i)] = g[i];
			__syncthreads();
			const int base = tid * R;
			int r[R_T];
			{
				int xr[MAXORD + R_T];
				load_run<R_T, MAXORD>(xs, base, R, xr);
				int q[MAXORD];
				if(type == SF_FIXED) {
#pragma unroll
					for(int j = 0; j < MAXORD; j++) q[j] = fixed_tap(order, j);
					run_residual_narrow<R_T, MAXORD>(xr, q, 0, r);
				}
				else {
#pragma unroll
					for(int j = 0; j < MAXORD; j++) q[j] = __ldg(&pl->qlp[j]);
					if(!pl->wide) run_residual_narrow<R_T, MAXORD>(xr, q, pl->shift, r);
					else (void)run_residual_wide<R_T, MAXORD>(xr, q, pl->shift, r, 0);
				}
			}
			// partitioned Rice coding: this thread's run is positions [base, base+R), residual positions >= order
			const int po = pl->porder;
			const int psize = bs >> po;
			const uint32_t plen = pl->method ? kRice2ParamLen : kRiceParamLen;
			uint32_t mybits = 0;
			// A run of R_T samples lies inside one partition whenever the partition size is a multiple of R_T (always for
			// the standard blocksizes and orders <= 8 here): one Rice parameter per run, and the parameter field precedes
			// exactly one position -- `order` in partition 0, the partition's first sample elsewhere.
			const bool one_partition = R_EXACT && (psize % R_T) == 0;
			uint32_t total;
			if(one_partition) {
				const int p = base / psize;
				const uint32_t k = __ldg(&pl->params[p]);
				const int first_res = p == 0 ? order : p * psize;
				uint32_t u[R_T];
#pragma unroll
				for(int m = 0; m < R_T; m++) {
					u[m] = ((uint32_t)r[m] << 1) ^ (uint32_t)(r[m] >> 31);
					if(base + m >= order) mybits += (u[m] >> k) + 1 + k + (base + m == first_res ? plen : 0u);
				}
				const uint32_t start = block_exclusive_scan<NT>(mybits, s_warp, &total);
				if(mybits) {
					BitPut bw;
					bw.init(words, bitpos + start);
#pragma unroll
					for(int m = 0; m < R_T; m++) {
						if(base + m >= order) {
							if(base + m == first_res) bw.put(k, plen);
							bw.skip(u[m] >> k);
							bw.put((1u << k) | (u[m] & ((1u << k) - 1u)), k + 1);
						}
					}
					bw.finish();
				}
			}
			else {
				{
					int p = base / psize;
					int next = (p + 1) * psize;
					uint32_t k = __ldg(&pl->params[p]);
#pragma unroll
					for(int m = 0; m < R_T; m++) {
						const int i = base + m;
						if((R_EXACT || m < R) && i >= order) {
							if(i == next) { p++; next += psize; k = __ldg(&pl->params[p]); }
							if(i == p * psize || i == order) mybits += plen;
							const uint32_t u = ((uint32_t)r[m] << 1) ^ (uint32_t)(r[m] >> 31);
							mybits += (u >> k) + 1 + k;
						}
					}
				}
				const uint32_t start = block_exclusive_scan<NT>(mybits, s_warp, &total);
				if(mybits) {
					BitPut bw;
					bw.init(words, bitpos + start);
					int p = base / psize;
					int next = (p + 1) * psize;
					uint32_t k = __ldg(&pl->params[p]);
#pragma unroll
					for(int m = 0; m < R_T; m++) {
						const int i = base + m;
						if((R_EXACT || m < R) && i >= order) {
							if(i == next) { p++; next += psize; k = __ldg(&pl->params[p]); }
							if(i == p * psize || i == order) bw.put(k, plen);
							const uint32_t u = ((uint32_t)r[m] << 1) ^ (uint32_t)(r[m] >> 31);
							bw.skip(u >> k);
							bw.put((1u << k) | (u & ((1u << k) - 1u)), k + 1);
						}
					}
					bw.finish();
				}
			}
			bitpos += total;
		}
		__syncthreads();
	}

	// ---- pad to byte, CRC-16, copy out (identical to k_emit)
	const uint32_t nbytes = (bitpos + 7) >> 3;
	{
		const uint32_t L = (nbytes + NT - 1) / NT;
		const int64_t cstart = (int64_t)nbytes - (int64_t)(NT - tid) * L;
		const int64_t cend = cstart + L;
		uint32_t crc = 0;
		for(int64_t b = (cstart < 0 ? 0 : cstart); b < cend; b++) {
			const uint32_t byte = (words[b >> 2] >> (24 - 8 * ((uint32_t)b & 3))) & 0xffu;
			crc = ((crc << 8) & 0xffffu) ^ s_crctab[((crc >> 8) ^ byte) & 0xffu];
		}
		// CRC of a concatenation: crc(A || B) = crc(A) * x^(8 len(B)) + crc(B) in GF(2)[x] / (x^16+x^15+x^2+1) (init 0).
		// Level s of the combine tree needs x^(8 L 2^s); thread s builds it from the constants x^(2^j) (at most popcount(8L) products).
		if(tid < 8) {
			uint32_t e = (8u * L) << tid, m = 1;
			for(int j = 0; e; j++, e >>= 1)
				if(e & 1u) m = gf16_mul(m, kCrcXPow2[j % 15]);
			s_mlev[tid] = m;
		}
		__syncthreads();
		{
			const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
			for(int sl = 0; sl < 5; sl++) {
				const uint32_t other = __shfl_down_sync(0xffffffffu, crc, 1 << sl);
				crc = gf16_mul(crc, s_mlev[sl]) ^ other;  // meaningful in lanes that are multiples of 2 << sl; lane 0 is what counts
			}
			if(lane == 0) s_crc[warp] = crc;
			__syncthreads();
			if(warp == 0) {
				crc = lane < NW ? s_crc[lane] : 0;
#pragma unroll
				for(int sl = 0; (1 << sl) < NW; sl++) {
					const uint32_t other = __shfl_down_sync(0xffffffffu, crc, 1 << sl);
					crc = gf16_mul(crc, s_mlev[5 + sl]) ^ other;
				}
				if(lane == 0) s_crc[0] = crc;
			}
			__syncthreads();
		}
	}
	if(tid == 0) {
		const uint32_t crc = s_crc[0];
		BitPut bw;
		bw.init(words, nbytes * 8);
		bw.put(crc, 16);
		bw.finish();
		frame_bytes[blk] = nbytes + 2;
		if(chan_assign_out) chan_assign_out[blk] = (uint32_t)ca;
	}
	__syncthreads();
	{
		const uint32_t nwords = (nbytes + 2 + 3) >> 2;
		uint32_t *dst = reinterpret_cast<uint32_t *>(slots + (size_t)blk * P.slot_stride);
		for(uint32_t i = tid; i < nwords; i += NT) dst[i] = __byte_perm(words[i], 0, 0x0123);
	}
}

}  // namespace fb200
