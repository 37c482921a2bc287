// decode_kernels.cuh -- sm_100a kernels of the FLAC batch frame decoder.
//
// Frames carry no length field and Rice codes have data-dependent lengths, so parsing a frame is
// a serial recurrence; parallelism is across frames:
//   k_dec_parse : one THREAD per frame -- header, subframes, Rice decode fused with the
//                 fixed/LPC restore (history in registers), planar int32 scratch output
//                 (stream_decoder.c:2373-3357, lpc.c:978-1491, fixed.c:571-629)
//   k_dec_crc   : one WARP per frame -- CRC-16 of the frame bytes, chunked + GF(2) combine
//                 (stream_decoder.c:2443-2452, crc.c:78-396)
//   k_dec_merge : undo channel decorrelation + interleave, fully coalesced
//                 (stream_decoder.c:3476-3527)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "fb200_internal.h"

namespace fb200 {

enum : uint32_t {
	DEC_OK = 0, DEC_BAD_SYNC = 1, DEC_BAD_HEADER = 2, DEC_CRC8 = 3, DEC_UNSUPPORTED = 4, DEC_PARSE = 5,
	DEC_LENGTH = 6, DEC_CRC16 = 7, DEC_MISMATCH = 8
};

struct DecK {
	int channels, bps, sample_rate, blocksize;  // STREAMINFO facts
	int bs_stride;                              // planar scratch stride per channel
};

struct DecFrameMeta {
	uint32_t status;
	uint32_t blocksize;
	uint32_t channel_assignment;  // 0 independent, 1 left/side, 2 right/side, 3 mid/side
	uint32_t channels;
};

// MSB-first bit reader over global memory with a two-word register cache.
struct BitReader {
	const uint32_t *words;  // aligned base
	uint32_t pos;           // bit position relative to words[0]
	uint32_t end;           // one past the last valid bit
	uint32_t w0, w1;        // big-endian words at index (pos>>5), +1
	uint32_t widx;
	__device__ __forceinline__ static uint32_t be(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
	__device__ __forceinline__ void init(const uint8_t *p, uint32_t nbytes)
	{
		const uintptr_t a = reinterpret_cast<uintptr_t>(p);
		words = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
		pos = (uint32_t)(a & 3) * 8;
		end = pos + nbytes * 8;
		widx = 0;
		w0 = be(__ldg(words));
		w1 = be(__ldg(words + 1));
	}
	__device__ __forceinline__ void sync_words()
	{
		const uint32_t wi = pos >> 5;
		if(wi != widx) {
			if(wi == widx + 1) { w0 = w1; w1 = be(__ldg(words + wi + 1)); }
			else { w0 = be(__ldg(words + wi)); w1 = be(__ldg(words + wi + 1)); }
			widx = wi;
		}
	}
	__device__ __forceinline__ uint32_t peek32()  // next 32 bits (zero-extended past `end` is the caller's problem)
	{
		sync_words();
		return __funnelshift_l(w1, w0, pos & 31);
	}
	__device__ __forceinline__ uint32_t get(uint32_t n)  // 1..32 bits
	{
		const uint32_t v = peek32() >> (32 - n);
		pos += n;
		return v;
	}
	__device__ __forceinline__ int32_t get_signed(uint32_t n)
	{
		const int32_t v = (int32_t)peek32() >> (32 - n);
		pos += n;
		return v;
	}
	__device__ __forceinline__ uint32_t unary()  // number of 0 bits before the next 1 (bitreader.c:725)
	{
		uint32_t q = 0;
		while(true) {
			const uint32_t v = peek32();
			if(v) {
				const uint32_t z = (uint32_t)__clz((int)v);
				pos += z + 1;
				return q + z;
			}
			q += 32;
			pos += 32;
			if(pos > end) return q;
		}
	}
	__device__ __forceinline__ bool overrun() const { return pos > end; }
};

// Restores MAXORD-tap predicted samples on the fly: hist[j] = sample (i-1-j).
template <int MAXORD, bool WIDE>
__device__ __forceinline__ int32_t predict(const int (&q)[MAXORD], const int (&hist)[MAXORD], int shift)
{
	if(WIDE) {
		long long sum = 0;
#pragma unroll
		for(int j = 0; j < MAXORD; j++) sum += (long long)q[j] * (long long)hist[j];
		return (int32_t)(sum >> shift);
	}
	else {
		int sum = 0;
#pragma unroll
		for(int j = 0; j < MAXORD; j++) sum += q[j] * hist[j];
		return sum >> shift;
	}
}

// Decode the residual of one subframe and restore the signal in the same pass.
// q[] = predictor taps (zero beyond the order), warm-up already in hist[] and written to out.
template <int MAXORD, bool WIDE>
__device__ bool decode_residual_restore(BitReader &br, uint32_t blocksize, uint32_t order, const int (&q)[MAXORD], int shift,
                                        int (&hist)[MAXORD], int32_t *__restrict__ out, uint32_t wasted)
{
	// stream_decoder.c:3299-3357 read_residual_partitioned_rice_
	const uint32_t method = br.get(2);
	if(method > 1) return false;
	const uint32_t plen = method ? kRice2ParamLen : kRiceParamLen, pesc = method ? kRice2Escape : kRiceEscape;
	const uint32_t po = br.get(4);
	const uint32_t psamples = blocksize >> po;
	if(po > 0 ? (psamples < order || (psamples << po) != blocksize) : blocksize < order) return false;
	uint32_t i = order;
	for(uint32_t p = 0; p < (1u << po); p++) {
		const uint32_t k = br.get(plen);
		const uint32_t pend = (po == 0) ? blocksize : (p + 1) * psamples;
		uint32_t raw = 0;
		const bool esc = k >= pesc;
		if(esc) raw = br.get(5);
		for(; i < pend; i++) {
			int32_t r;
			if(!esc) {
				// deduplication/bitreader_read_rice_signed_block.c: unary MSBs, k LSBs, zig-zag
				const uint32_t msbs = br.unary();
				const uint32_t u = (msbs << k) | (k ? br.get(k) : 0u);
				r = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
			}
			else r = raw ? br.get_signed(raw) : 0;
			const int32_t v = r + predict<MAXORD, WIDE>(q, hist, shift);
#pragma unroll
			for(int j = MAXORD - 1; j > 0; j--) hist[j] = hist[j - 1];
			hist[0] = v;
			out[i] = (int32_t)((uint32_t)v << wasted);
		}
		if(br.overrun()) return false;
	}
	return true;
}

__device__ __forceinline__ uint32_t dec_ilog2(uint32_t v) { return 31u - (uint32_t)__clz((int)v); }
__device__ __forceinline__ uint32_t dec_silog2(long long v)
{
	if(v == 0) return 0;
	if(v == -1) return 2;
	v = (v < 0) ? (-(v + 1)) : v;
	return (63u - (uint32_t)__clzll(v)) + 2;
}

template <int MAXORD>
__device__ bool decode_predicted(BitReader &br, uint32_t blocksize, uint32_t order, const int *qsrc, int shift, bool wide,
                                 uint32_t bps, int32_t *__restrict__ out, uint32_t wasted)
{
	int q[MAXORD], hist[MAXORD];
#pragma unroll
	for(int j = 0; j < MAXORD; j++) { q[j] = (j < (int)order) ? qsrc[j] : 0; hist[j] = 0; }
	// warm-up samples were already read into out[0..order) (unshifted) by the caller
#pragma unroll
	for(int j = 0; j < MAXORD; j++)
		if(j < (int)order) hist[j] = out[order - 1 - j];
	for(uint32_t i = 0; i < order; i++) out[i] = (int32_t)((uint32_t)out[i] << wasted);
	(void)bps;
	if(wide) return decode_residual_restore<MAXORD, true>(br, blocksize, order, q, shift, hist, out, wasted);
	return decode_residual_restore<MAXORD, false>(br, blocksize, order, q, shift, hist, out, wasted);
}

// stream_decoder.c:2949-3297 read_subframe_*
__device__ bool decode_subframe(BitReader &br, uint32_t blocksize, uint32_t bps, int32_t *__restrict__ out)
{
	uint32_t x = br.get(8);
	if(x & 0x80) return false;
	uint32_t wasted = 0;
	if(x & 1) {
		wasted = br.unary() + 1;
		if(wasted >= bps) return false;
		bps -= wasted;
	}
	x &= 0xfe;
	if(x == 0) {  // CONSTANT
		const int32_t v = (int32_t)((uint32_t)br.get_signed(bps) << wasted);
		for(uint32_t i = 0; i < blocksize; i++) out[i] = v;
		return !br.overrun();
	}
	if(x == 2) {  // VERBATIM
		for(uint32_t i = 0; i < blocksize; i++) out[i] = (int32_t)((uint32_t)br.get_signed(bps) << wasted);
		return !br.overrun();
	}
	if(x >= 16 && x <= 24) {  // FIXED (fixed.c:571-629 as taps)
		const uint32_t order = (x >> 1) & 7;
		if(order > 4 || blocksize <= order) return false;
		for(uint32_t i = 0; i < order; i++) out[i] = br.get_signed(bps);
		const int tab[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
		// 32-bit restore when bps + order <= 32 (stream_decoder.c:3139-3143); int64 otherwise. With bps <= 25
		// both agree with wrapping int32 arithmetic on the sums the reference forms.
		return decode_predicted<4>(br, blocksize, order, tab[order], 0, false, bps, out, wasted);
	}
	if(x >= 64) {  // LPC
		const uint32_t order = ((x >> 1) & 31) + 1;
		if(blocksize <= order) return false;
		for(uint32_t i = 0; i < order; i++) out[i] = br.get_signed(bps);
		const uint32_t prec = br.get(4);
		if(prec == 15) return false;
		const uint32_t precision = prec + 1;
		const int shift = br.get_signed(5);
		if(shift < 0) return false;
		int q[FB200_MAX_LPC_ORDER];
		uint32_t abs_sum = 0;
		for(uint32_t j = 0; j < order; j++) { q[j] = br.get_signed(precision); abs_sum += (uint32_t)abs(q[j]); }
		// variant rule of stream_decoder.c:3243-3247 (lpc.c:942-968)
		const unsigned long long max_abs = 1ull << (bps - 1);
		const unsigned long long max_pred = max_abs * abs_sum;
		const unsigned long long max_after = (unsigned long long)(-1 * ((-1 * (long long)max_pred) >> shift));
		const bool wide = !(dec_silog2((long long)(max_abs + max_after)) <= 32 && dec_silog2((long long)max_pred) <= 32);
		if(order <= 8) return decode_predicted<8>(br, blocksize, order, q, shift, wide, bps, out, wasted);
		if(order <= 12) return decode_predicted<12>(br, blocksize, order, q, shift, wide, bps, out, wasted);
		return decode_predicted<32>(br, blocksize, order, q, shift, wide, bps, out, wasted);
	}
	return false;  // reserved subframe type
}

__global__ void __launch_bounds__(64) k_dec_parse(DecK P, const uint8_t *__restrict__ frames, const unsigned long long *__restrict__ offsets,
                                                 int nframes, int32_t *__restrict__ scratch, DecFrameMeta *__restrict__ meta)
{
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	if(f >= nframes) return;
	const unsigned long long off = offsets[f];
	const uint32_t len = (uint32_t)(offsets[f + 1] - off);
	DecFrameMeta M;
	M.status = DEC_OK; M.blocksize = 0; M.channel_assignment = 0; M.channels = 0;
	if(len < 6) { M.status = DEC_LENGTH; meta[f] = M; return; }
	BitReader br;
	br.init(frames + off, len);
	const uint32_t start = br.pos;

	// ---- frame header (stream_decoder.c:2624-2947)
	if(br.get(14) != 0x3ffe) { M.status = DEC_BAD_SYNC; meta[f] = M; return; }
	if(br.get(1) != 0) { M.status = DEC_BAD_HEADER; meta[f] = M; return; }
	const uint32_t variable = br.get(1);
	const uint32_t bs_code = br.get(4), sr_code = br.get(4), ca_code = br.get(4), bps_code = br.get(3);
	if(br.get(1) != 0) { M.status = DEC_BAD_HEADER; meta[f] = M; return; }
	{   // UTF-8 coded frame/sample number
		const uint32_t first = br.get(8);
		int n;
		if(!(first & 0x80)) n = 0;
		else if((first & 0xE0) == 0xC0) n = 1;
		else if((first & 0xF0) == 0xE0) n = 2;
		else if((first & 0xF8) == 0xF0) n = 3;
		else if((first & 0xFC) == 0xF8) n = 4;
		else if((first & 0xFE) == 0xFC) n = 5;
		else if(first == 0xFE && variable) n = 6;
		else { M.status = DEC_BAD_HEADER; meta[f] = M; return; }
		for(int k = 0; k < n; k++)
			if((br.get(8) & 0xC0) != 0x80) { M.status = DEC_BAD_HEADER; meta[f] = M; return; }
	}
	uint32_t blocksize;
	switch(bs_code) {
		case 0: M.status = DEC_BAD_HEADER; meta[f] = M; return;
		case 1: blocksize = 192; break;
		case 2: case 3: case 4: case 5: blocksize = 576u << (bs_code - 2); break;
		case 6: blocksize = br.get(8) + 1; break;
		case 7: blocksize = br.get(16) + 1; break;
		default: blocksize = 256u << (bs_code - 8); break;
	}
	if(sr_code == 12) (void)br.get(8);
	else if(sr_code == 13 || sr_code == 14) (void)br.get(16);
	else if(sr_code == 15) { M.status = DEC_BAD_HEADER; meta[f] = M; return; }
	{   // CRC-8 over the header bytes (crc.c:39-76)
		const uint32_t nb = (br.pos - start) >> 3;
		const uint8_t *p = frames + off;
		uint32_t crc = 0;
		for(uint32_t b = 0; b < nb; b++) {
			crc ^= p[b];
			for(int j = 0; j < 8; j++) crc = (crc & 0x80u) ? ((crc << 1) ^ 0x07u) & 0xffu : (crc << 1) & 0xffu;
		}
		if(br.get(8) != crc) { M.status = DEC_CRC8; meta[f] = M; return; }
	}
	uint32_t channels, ca;
	if(ca_code < 8) { channels = ca_code + 1; ca = 0; }
	else if(ca_code <= 10) { channels = 2; ca = ca_code - 7; }
	else { M.status = DEC_BAD_HEADER; meta[f] = M; return; }
	uint32_t bps;
	switch(bps_code) {
		case 0: bps = (uint32_t)P.bps; break;
		case 1: bps = 8; break; case 2: bps = 12; break; case 4: bps = 16; break;
		case 5: bps = 20; break; case 6: bps = 24; break; case 7: bps = 32; break;
		default: M.status = DEC_BAD_HEADER; meta[f] = M; return;
	}
	M.blocksize = blocksize; M.channel_assignment = ca; M.channels = channels;
	if(blocksize > (uint32_t)P.blocksize || channels != (uint32_t)P.channels || bps != (uint32_t)P.bps || bps > 24) {
		M.status = (bps > 24) ? DEC_UNSUPPORTED : DEC_MISMATCH;
		meta[f] = M;
		return;
	}
	int32_t *base = scratch + (size_t)f * P.channels * P.bs_stride;
	for(uint32_t c = 0; c < channels; c++) {
		uint32_t sub_bps = bps;
		if((ca == 1 && c == 1) || (ca == 2 && c == 0) || (ca == 3 && c == 1)) sub_bps++;
		if(!decode_subframe(br, blocksize, sub_bps, base + (size_t)c * P.bs_stride)) { M.status = DEC_PARSE; meta[f] = M; return; }
	}
	// zero padding to a byte boundary, then the CRC-16 footer must end exactly at the frame end
	const uint32_t consumed_bits = ((br.pos - start) + 7) & ~7u;
	if(consumed_bits + 16 != len * 8) M.status = DEC_LENGTH;
	meta[f] = M;
}

// GF(2)[x] multiply mod x^16+x^15+x^2+1
__device__ __forceinline__ uint32_t dec_gf16_mul(uint32_t a, uint32_t b)
{
	uint32_t r = 0;
#pragma unroll
	for(int i = 15; i >= 0; i--) {
		r = (r & 0x8000u) ? ((r << 1) ^ 0x8005u) & 0xffffu : (r << 1);
		if((b >> i) & 1u) r ^= a;
	}
	return r;
}

// One warp per frame; lane l owns a contiguous chunk aligned to the end of the frame.
__global__ void __launch_bounds__(128) k_dec_crc(const uint8_t *__restrict__ frames, const unsigned long long *__restrict__ offsets,
                                                int nframes, DecFrameMeta *__restrict__ meta)
{
	__shared__ uint16_t s_tab[256];
	for(int e = threadIdx.x; e < 256; e += blockDim.x) {
		uint32_t c = (uint32_t)e << 8;
#pragma unroll
		for(int j = 0; j < 8; j++) c = (c & 0x8000u) ? ((c << 1) ^ 0x8005u) : (c << 1);
		s_tab[e] = (uint16_t)c;
	}
	__syncthreads();
	const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if(f >= nframes) return;
	const unsigned long long off = offsets[f];
	const uint32_t len = (uint32_t)(offsets[f + 1] - off);
	if(len < 3) return;
	const uint32_t nbytes = len - 2;
	const uint8_t *p = frames + off;
	const uint32_t L = (nbytes + 31) / 32;
	const long long cstart = (long long)nbytes - (long long)(32 - lane) * L, cend = cstart + L;
	uint32_t crc = 0;
	for(long long b = cstart < 0 ? 0 : cstart; b < cend; b++)
		crc = ((crc << 8) & 0xffffu) ^ s_tab[((crc >> 8) ^ p[b]) & 0xffu];
	// m = x^(8L) mod P, squared per tree level
	uint32_t m = 1, bb = 2, e = 8 * L;
	while(e) {
		if(e & 1) m = dec_gf16_mul(m, bb);
		bb = dec_gf16_mul(bb, bb);
		e >>= 1;
	}
#pragma unroll
	for(int s = 0; s < 5; s++) {
		const uint32_t other = __shfl_down_sync(0xffffffffu, crc, 1 << s);
		if((lane & ((2 << s) - 1)) == 0) crc = dec_gf16_mul(crc, m) ^ other;
		m = dec_gf16_mul(m, m);
	}
	if(lane == 0) {
		const uint32_t want = ((uint32_t)p[nbytes] << 8) | p[nbytes + 1];
		if(crc != want && meta[f].status == DEC_OK) meta[f].status = DEC_CRC16;
	}
}

// Undo the inter-channel decorrelation and interleave (stream_decoder.c:3476-3527).
__global__ void __launch_bounds__(256) k_dec_merge(DecK P, const int32_t *__restrict__ scratch, const DecFrameMeta *__restrict__ meta,
                                                  int32_t *__restrict__ pcm, unsigned long long capacity_samples, uint32_t *__restrict__ status_out)
{
	const int f = blockIdx.x;
	const DecFrameMeta M = meta[f];
	if(threadIdx.x == 0 && status_out) status_out[f] = M.status | (M.blocksize << 8);  // low byte: status, upper: decoded blocksize
	if(M.status != DEC_OK) return;
	const int bs = (int)M.blocksize, ch = P.channels;
	const unsigned long long first = (unsigned long long)f * P.blocksize;
	if(first + bs > capacity_samples) return;
	const int32_t *src = scratch + (size_t)f * ch * P.bs_stride;
	int32_t *dst = pcm + first * ch;
	if(ch == 2) {
		int2 *d2 = reinterpret_cast<int2 *>(dst);
		for(int i = threadIdx.x; i < bs; i += blockDim.x) {
			const int32_t a = src[i], b = src[P.bs_stride + i];
			int32_t l, r;
			switch(M.channel_assignment) {
				case 1: l = a; r = a - b; break;                 // left/side
				case 2: l = a + b; r = b; break;                 // right/side  (a = side)
				case 3: {                                         // mid/side
					const int32_t mid = (int32_t)(((uint32_t)a << 1) | ((uint32_t)b & 1u));
					l = (mid + b) >> 1; r = (mid - b) >> 1;
					break;
				}
				default: l = a; r = b; break;
			}
			d2[i] = make_int2(l, r);
		}
	}
	else {
		for(int i = threadIdx.x; i < bs * ch; i += blockDim.x) {
			const int s = i / ch, c = i - s * ch;
			dst[i] = src[(size_t)c * P.bs_stride + s];
		}
	}
}

}  // namespace fb200
