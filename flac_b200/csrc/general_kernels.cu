// general_kernels.cu -- the general kernels (any blocksize / channel count) and their launchers.
#include "encode_kernels.cuh"

namespace fb200 {

void launch_prep(const EncK &k, const int32_t *pcm, int32_t *sig, SigMeta *meta, int *blkflags, int nb, cudaStream_t st)
{
	k_prep<true><<<nb, 256, 0, st>>>(k, pcm, sig, meta, blkflags);
}

void launch_meta(const EncK &k, const int32_t *pcm, SigMeta *meta, int *blkflags, int nb, cudaStream_t st)
{
	k_prep<false><<<nb, 256, 0, st>>>(k, pcm, nullptr, meta, blkflags);
}

void launch_unpack(const void *packed, int bytes_per_sample, int32_t *pcm, unsigned long long n, int bps, int *err, cudaStream_t st)
{
	const unsigned long long threads = (n + 3) / 4;
	const unsigned grid = (unsigned)((threads + 255) / 256);
	if(grid == 0) return;
	if(bytes_per_sample == 2) k_unpack<2><<<grid, 256, 0, st>>>(static_cast<const uint8_t *>(packed), pcm, n, bps, err);
	else k_unpack<3><<<grid, 256, 0, st>>>(static_cast<const uint8_t *>(packed), pcm, n, bps, err);
}

template <int LAGS>
static void autoc_general(const EncK &k, const int32_t *sig, const SigMeta *meta, const float *windows, const DevSection *secs, double *autoc, int nitems, cudaStream_t st)
{
	const int total = nitems * k.nsec;
	k_autoc<LAGS><<<(total + 127) / 128, 128, 0, st>>>(k, sig, meta, windows, secs, autoc, nitems);
}

void launch_autoc_general(const EncK &k, const int32_t *sig, const SigMeta *meta, const float *windows, const DevSection *secs, double *autoc, int nitems, cudaStream_t st)
{
	if(k.lags <= 7) autoc_general<7>(k, sig, meta, windows, secs, autoc, nitems, st);
	else if(k.lags <= 9) autoc_general<9>(k, sig, meta, windows, secs, autoc, nitems, st);
	else if(k.lags <= 13) autoc_general<13>(k, sig, meta, windows, secs, autoc, nitems, st);
	else if(k.lags <= 17) autoc_general<17>(k, sig, meta, windows, secs, autoc, nitems, st);
	else autoc_general<33>(k, sig, meta, windows, secs, autoc, nitems, st);
}

void launch_minbr_flags(const EncK &k, const SubframePlan *plans, const int *blkflags, int nb, int *flags, cudaStream_t st)
{
	k_minbr_flags<<<(nb + 127) / 128, 128, 0, st>>>(k, plans, blkflags, nb, flags);
}

void launch_lpc(const EncK &k, const double *autoc, const DevCand *cands, SigMeta *meta, const uint32_t *sigor, CandDesc *cdesc, int nitems, int autoc_unshifted, cudaStream_t st)
{
	const int total = nitems * k.nwin;
	if(k.max_order <= 8) k_lpc<8><<<(total + 127) / 128, 128, 0, st>>>(k, autoc, cands, meta, sigor, cdesc, nitems, autoc_unshifted);
	else if(k.max_order <= 12) k_lpc<12><<<(total + 127) / 128, 128, 0, st>>>(k, autoc, cands, meta, sigor, cdesc, nitems, autoc_unshifted);
	else k_lpc<32><<<(total + 127) / 128, 128, 0, st>>>(k, autoc, cands, meta, sigor, cdesc, nitems, autoc_unshifted);
}

void launch_search_general(const EncK &k, size_t smem, const int32_t *sig, const SigMeta *meta, const CandDesc *cdesc, SubframePlan *plans, int nitems, cudaStream_t st)
{
	k_search<<<nitems, 128, smem, st>>>(k, sig, meta, cdesc, plans);
}

void launch_emit_general(const EncK &k, size_t smem, const int32_t *sig, const int *blkflags, const SubframePlan *plans, uint8_t *slots, uint32_t *frame_bytes, uint32_t *chan_assign, int nb, cudaStream_t st)
{
	k_emit<<<nb, 256, smem, st>>>(k, sig, blkflags, plans, slots, frame_bytes, chan_assign);
}

void launch_scan(const uint32_t *bytes, int n, unsigned long long *offsets, unsigned long long *running, cudaStream_t st)
{
	k_scan<<<1, 1024, 0, st>>>(bytes, n, offsets, running);
}

void launch_gather(const EncK &k, const uint8_t *slots, const uint32_t *bytes, const unsigned long long *offsets, uint8_t *out, unsigned long long capacity, int *err, int nb, cudaStream_t st)
{
	k_gather<<<nb, 256, 0, st>>>(k, slots, bytes, offsets, out, capacity, err);
}

void launch_debug_log(const double *dx, double *dy, int n) { k_debug_log<<<(n + 255) / 256, 256>>>(dx, dy, n); }

// The dynamic shared-memory opt-in is a per-function, per-device attribute: set it ONCE to the device maximum, so that
// encoders with different blocksizes in one process never lower each other's limit.
void general_kernels_init(int device)
{
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, device);
	cudaFuncSetAttribute(k_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin - 8192);
	cudaFuncSetAttribute(k_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin - 4096);
}

}  // namespace fb200
