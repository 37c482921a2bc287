// search_kernel.cu -- k_search4 and its launcher.
#include "search_kernel.cuh"

namespace fb200 {

template <int MO>
static void search4(const EncK &k, int rt, size_t smem, const int32_t *sig, const SigMeta *meta, const CandDesc *cdesc, SubframePlan *plans, int nitems, cudaStream_t st)
{
	const int grid = (nitems + 1) / 2;
	const bool widek = k.bps > 16;
	if(rt == 32) {
		if(widek) k_search4<32, MO, 2, true><<<grid, 64, smem, st>>>(k, sig, meta, cdesc, plans, nitems);
		else k_search4<32, MO, 2, false><<<grid, 64, smem, st>>>(k, sig, meta, cdesc, plans, nitems);
	}
	else {
		if(widek) k_search4<36, MO, 2, true><<<grid, 64, smem, st>>>(k, sig, meta, cdesc, plans, nitems);
		else k_search4<36, MO, 2, false><<<grid, 64, smem, st>>>(k, sig, meta, cdesc, plans, nitems);
	}
}

void launch_search4(const EncK &k, int rt, int maxord_t, size_t smem, const int32_t *sig, const SigMeta *meta, const CandDesc *cdesc, SubframePlan *plans, int nitems, cudaStream_t st)
{
	if(maxord_t == 8) search4<8>(k, rt, smem, sig, meta, cdesc, plans, nitems, st);
	else if(maxord_t == 12) search4<12>(k, rt, smem, sig, meta, cdesc, plans, nitems, st);
	else search4<32>(k, rt, smem, sig, meta, cdesc, plans, nitems, st);
}

template <int MO>
static void search4_attrs()
{
	cudaFuncSetAttribute(k_search4<32, MO, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
	cudaFuncSetAttribute(k_search4<36, MO, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
	cudaFuncSetAttribute(k_search4<32, MO, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
	cudaFuncSetAttribute(k_search4<36, MO, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
}

void search4_init(int)
{
	search4_attrs<8>();
	search4_attrs<12>();
	search4_attrs<32>();
}

}  // namespace fb200
