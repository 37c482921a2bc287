// autoc_kernel.cu -- k_autoc3 and its launcher.
#include "autoc_kernel.cuh"

namespace fb200 {

template <int LAGS, int U, int K, int STAGES>
static void autoc3(const EncK &k, const int32_t *sig, const SigMeta *meta, const float *windows, const DevSection *secs, double *autoc, int nitems, cudaStream_t st)
{
	const int groups = (nitems + 31) / 32;
	k_autoc3<LAGS, U, K, STAGES><<<groups * k.nsec, 32, autoc3_smem_bytes<LAGS, U, K, STAGES>(), st>>>(k, sig, meta, windows, secs, autoc, nitems);
}

// warp per 32 chains, asynchronously staged tiles; tile shapes tuned per lag count (DESIGN.md "Kernels")
void launch_autoc3(const EncK &k, const int32_t *sig, const SigMeta *meta, const float *windows, const DevSection *secs, double *autoc, int nitems, cudaStream_t st)
{
	if(k.lags <= 7) autoc3<7, 28, 2, 4>(k, sig, meta, windows, secs, autoc, nitems, st);
	else if(k.lags <= 9) autoc3<9, 36, 2, 4>(k, sig, meta, windows, secs, autoc, nitems, st);
	else if(k.lags <= 13) autoc3<13, 52, 1, 4>(k, sig, meta, windows, secs, autoc, nitems, st);
	else if(k.lags <= 17) autoc3<17, 68, 1, 3>(k, sig, meta, windows, secs, autoc, nitems, st);
	else autoc3<33, 132, 1, 3>(k, sig, meta, windows, secs, autoc, nitems, st);
}

void autoc3_init(int)
{
	cudaFuncSetAttribute(k_autoc3<33, 132, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)autoc3_smem_bytes<33, 132, 1, 3>());
}

}  // namespace fb200
