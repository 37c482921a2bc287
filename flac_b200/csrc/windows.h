// flac_b200 — host-side apodization window tables.
//
// The windows are generated once per (encoder, blocksize) on the host and uploaded; the kernels
// only ever see a float table, so every apodization function the reference knows costs nothing
// on the device.  They are generated on the host (not in a kernel) because the reference's
// values come out of the host libm's cosf/exp and must match to the bit.
//
// Restates src/libFLAC/window.c:50-302 (one generator per FLAC__window_*).  The arithmetic
// types of every sub-expression are kept as the reference's C promotes them (float coefficient
// × double M_PI → double argument → cosf(float); float sums left to right); this translation
// unit is compiled with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstdint>

#include "../../include/flac_b200.h"

namespace fbwin {

// window[n] = a0 ∓ a1·cos(2πn/N) ± a2·cos(4πn/N) ∓ … summed left to right in float
// (blackman, blackman_harris, flattop, hamming, hann, kaiser_bessel, nuttall; window.c:78-171).
static inline void cosine_sum(float *w, int32_t L, float a0, const float *a, int nterms)
{
	const int32_t N = L - 1;
	for(int32_t n = 0; n < L; n++) {
		float acc = a0;
		for(int k = 0; k < nterms; k++) {
			const float harmonic = (float)(2 * (k + 1));  // 2.0f, 4.0f, 6.0f, 8.0f
			const float c = cosf((float)(harmonic * M_PI * n / N));
			const float term = a[k] * c;
			acc = (k & 1) ? acc + term : acc - term;
		}
		w[n] = acc;
	}
}

static inline void rectangle(float *w, int32_t L) { for(int32_t n = 0; n < L; n++) w[n] = 1.0f; }  // window.c:173-179

static inline void hann(float *w, int32_t L) { const float a[] = {0.5f}; cosine_sum(w, L, 0.5f, a, 1); }  // :146-153

// :50-67 (both parities of L ramp up through n <= (L-1)/2, which is N/2 for odd L and L/2-1 for even L)
static inline void bartlett(float *w, int32_t L)
{
	const int32_t N = L - 1, peak = (L & 1) ? N / 2 : L / 2 - 1;
	for(int32_t n = 0; n < L; n++) {
		const float ramp = 2.0f * n / (float)N;
		w[n] = n <= peak ? ramp : 2.0f - ramp;
	}
}

// :69-76
static inline void bartlett_hann(float *w, int32_t L)
{
	const int32_t N = L - 1;
	for(int32_t n = 0; n < L; n++) {
		const float x = (float)n / (float)N;
		w[n] = (float)(0.62f - 0.48f * fabsf(x - 0.5f) - 0.38f * cosf((float)(2.0f * M_PI * x)));
	}
}

// :97-108 connes, :292-302 welch: both are polynomials of k = (n - N/2)/(N/2) in double
static inline void connes_or_welch(float *w, int32_t L, bool squared)
{
	const int32_t N = L - 1;
	const double N2 = (double)N / 2.;
	for(int32_t n = 0; n <= N; n++) {
		const double k = ((double)n - N2) / N2;
		const double v = 1.0f - k * k;
		w[n] = squared ? (float)(v * v) : (float)v;
	}
}

// :119-135
static inline void gauss(float *w, int32_t L, float stddev)
{
	if(!(stddev > 0.0f && stddev <= 0.5f)) { gauss(w, L, 0.25f); return; }
	const int32_t N = L - 1;
	const double N2 = (double)N / 2.;
	for(int32_t n = 0; n <= N; n++) {
		const double k = ((double)n - N2) / (stddev * N2);
		w[n] = (float)exp(-0.5f * k * k);
	}
}

// :181-197 (the two parities differ only in where the ramp turns: (L+1)/2 vs L/2, i.e. ceil(L/2))
static inline void triangle(float *w, int32_t L)
{
	const int32_t turn = (L + 1) / 2;
	for(int32_t n = 1; n <= L; n++)
		w[n - 1] = n <= turn ? 2.0f * n / ((float)L + 1.0f) : (float)(2 * (L - n + 1)) / ((float)L + 1.0f);
}

// half-cosine taper value used by the tukey family: 0.5 - 0.5·cos(π·i/Np)
static inline float taper(int32_t i, int32_t Np) { return (float)(0.5f - 0.5f * cosf((float)(M_PI * i / Np))); }

// :199-222
static inline void tukey(float *w, int32_t L, float p)
{
	if(p <= 0.0) rectangle(w, L);
	else if(p >= 1.0) hann(w, L);
	else if(!(p > 0.0f && p < 1.0f)) tukey(w, L, 0.5f);
	else {
		const int32_t Np = (int32_t)(p / 2.0f * L) - 1;
		rectangle(w, L);
		if(Np > 0)
			for(int32_t n = 0; n <= Np; n++) {
				w[n] = taper(n, Np);
				w[L - Np - 1 + n] = taper(n + Np, Np);
			}
	}
}

static inline float clamp_multi_p(float p)
{
	if(p <= 0.0f) return 0.05f;
	if(p >= 1.0f) return 0.95f;
	if(!(p > 0.0f && p < 1.0f)) return 0.5f;
	return p;
}

// :224-254 — zero outside [start·L, end·L), tukey(p) inside
static inline void partial_tukey(float *w, int32_t L, float p, float start, float end)
{
	p = clamp_multi_p(p);
	const int32_t start_n = (int32_t)(start * L), end_n = (int32_t)(end * L), N = end_n - start_n;
	const int32_t Np = (int32_t)(p / 2.0f * N);
	int32_t n = 0, i;
	for(; n < start_n && n < L; n++) w[n] = 0.0f;
	for(i = 1; n < (start_n + Np) && n < L; n++, i++) w[n] = taper(i, Np);
	for(; n < (end_n - Np) && n < L; n++) w[n] = 1.0f;
	for(i = Np; n < end_n && n < L; n++, i--) w[n] = taper(i, Np);
	for(; n < L; n++) w[n] = 0.0f;
}

// :256-290 — tukey(p) on [0, start·L) and on [end·L, L), zero between
static inline void punchout_tukey(float *w, int32_t L, float p, float start, float end)
{
	p = clamp_multi_p(p);
	const int32_t start_n = (int32_t)(start * L), end_n = (int32_t)(end * L);
	const int32_t Ns = (int32_t)(p / 2.0f * start_n), Ne = (int32_t)(p / 2.0f * (L - end_n));
	int32_t n = 0, i;
	for(i = 1; n < Ns && n < L; n++, i++) w[n] = taper(i, Ns);
	for(; n < start_n - Ns && n < L; n++) w[n] = 1.0f;
	for(i = Ns; n < start_n && n < L; n++, i--) w[n] = taper(i, Ns);
	for(; n < end_n && n < L; n++) w[n] = 0.0f;
	for(i = 1; n < end_n + Ne && n < L; n++, i++) w[n] = taper(i, Ne);
	for(; n < L - Ne && n < L; n++) w[n] = 1.0f;
	for(i = Ne; n < L; n++, i--) w[n] = taper(i, Ne);
}

// resize_buffers_'s switch (stream_encoder.c:2913-2975). Returns false for an unknown type.
static inline bool make(const fb200_apodization &ap, float *w, int32_t L)
{
	switch(ap.type) {
		case FB200_APOD_TUKEY:
		case FB200_APOD_SUBDIVIDE_TUKEY: tukey(w, L, ap.p); return true;
		case FB200_APOD_BARTLETT: bartlett(w, L); return true;
		case FB200_APOD_BARTLETT_HANN: bartlett_hann(w, L); return true;
		case FB200_APOD_BLACKMAN: { const float a[] = {0.5f, 0.08f}; cosine_sum(w, L, 0.42f, a, 2); return true; }
		case FB200_APOD_BLACKMAN_HARRIS_4TERM_92DB_SIDELOBE: { const float a[] = {0.48829f, 0.14128f, 0.01168f}; cosine_sum(w, L, 0.35875f, a, 3); return true; }
		case FB200_APOD_CONNES: connes_or_welch(w, L, true); return true;
		case FB200_APOD_FLATTOP: { const float a[] = {0.41663158f, 0.277263158f, 0.083578947f, 0.006947368f}; cosine_sum(w, L, 0.21557895f, a, 4); return true; }
		case FB200_APOD_GAUSS: gauss(w, L, ap.p); return true;
		case FB200_APOD_HAMMING: { const float a[] = {0.46f}; cosine_sum(w, L, 0.54f, a, 1); return true; }
		case FB200_APOD_HANN: hann(w, L); return true;
		case FB200_APOD_KAISER_BESSEL: { const float a[] = {0.498f, 0.098f, 0.001f}; cosine_sum(w, L, 0.402f, a, 3); return true; }
		case FB200_APOD_NUTTALL: { const float a[] = {0.4891775f, 0.1365995f, 0.0106411f}; cosine_sum(w, L, 0.3635819f, a, 3); return true; }
		case FB200_APOD_RECTANGLE: rectangle(w, L); return true;
		case FB200_APOD_TRIANGLE: triangle(w, L); return true;
		case FB200_APOD_PARTIAL_TUKEY: partial_tukey(w, L, ap.p, ap.start, ap.end); return true;
		case FB200_APOD_PUNCHOUT_TUKEY: punchout_tukey(w, L, ap.p, ap.start, ap.end); return true;
		case FB200_APOD_WELCH: connes_or_welch(w, L, false); return true;
	}
	return false;
}

}  // namespace fbwin
