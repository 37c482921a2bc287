// search_kernel.cu -- k_search5 and its launcher.
#include "search_kernel.cuh"

namespace fb200 {

template <int MO, int WPS>
static void search5(const EncK &k, int rt, size_t smem, const int32_t *pcm, const SigMeta *meta, const CandDesc *cdesc, SubframePlan *plans, int nb, cudaStream_t st)
{
	const int gsz = k.sig_group ? k.sig_group : k.nsig;  // signals per CTA
	const int nt = 32 * WPS * gsz;
	nb *= (k.nsig + gsz - 1) / gsz;
	const bool widek = k.bps > 16 || (WPS == 2 && k.f64b);
	if(rt == 32) {
		if(widek) k_search5<32, MO, WPS, true><<<nb, nt, smem, st>>>(k, pcm, meta, cdesc, plans);
		else k_search5<32, MO, WPS, false><<<nb, nt, smem, st>>>(k, pcm, meta, cdesc, plans);
	}
	else {
		if(widek) k_search5<36, MO, WPS, true><<<nb, nt, smem, st>>>(k, pcm, meta, cdesc, plans);
		else k_search5<36, MO, WPS, false><<<nb, nt, smem, st>>>(k, pcm, meta, cdesc, plans);
	}
}

void launch_search5(const EncK &k, int rt, int maxord_t, int wps, size_t smem, const int32_t *pcm, const SigMeta *meta, const CandDesc *cdesc, SubframePlan *plans, int nb, cudaStream_t st)
{
	if(wps == 2) {
		if(maxord_t == 8) search5<8, 2>(k, rt, smem, pcm, meta, cdesc, plans, nb, st);
		else if(maxord_t == 12) search5<12, 2>(k, rt, smem, pcm, meta, cdesc, plans, nb, st);
		else search5<32, 2>(k, rt, smem, pcm, meta, cdesc, plans, nb, st);
	}
	else {
		if(maxord_t == 8) search5<8, 1>(k, rt, smem, pcm, meta, cdesc, plans, nb, st);
		else if(maxord_t == 12) search5<12, 1>(k, rt, smem, pcm, meta, cdesc, plans, nb, st);
		else search5<32, 1>(k, rt, smem, pcm, meta, cdesc, plans, nb, st);
	}
}

size_t search5_smem(int bs, int rt, int nsig, int wps, int max_po) { return search5_smem_bytes(bs, rt, nsig, wps, max_po); }

template <int MO, int WPS>
static void search5_attrs()
{
	cudaFuncSetAttribute(k_search5<32, MO, WPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_search5<36, MO, WPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_search5<32, MO, WPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_search5<36, MO, WPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

void search5_init(int)
{
	search5_attrs<8, 1>(); search5_attrs<12, 1>(); search5_attrs<32, 1>();
	search5_attrs<8, 2>(); search5_attrs<12, 2>(); search5_attrs<32, 2>();
}

}  // namespace fb200
