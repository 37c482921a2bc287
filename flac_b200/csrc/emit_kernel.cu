// emit_kernel.cu -- k_emit3 and its launcher.
#include "emit_kernel.cuh"

namespace fb200 {

template <int MO>
static void emit3(const EncK &k, int rt, size_t smem, const Emit3Args &a, int nb, cudaStream_t st)
{
	const int nt = (k.bs / rt) * k.channels;
	if(k.bps > 16) {
		if(rt == 32) {
			if(k.channels == 2) k_emit3<32, MO, 2, true, false><<<nb, nt, smem, st>>>(k, a);
			else k_emit3<32, MO, 1, true, false><<<nb, nt, smem, st>>>(k, a);
		}
		else {
			if(k.channels == 2) k_emit3<36, MO, 2, true, false><<<nb, nt, smem, st>>>(k, a);
			else k_emit3<36, MO, 1, true, false><<<nb, nt, smem, st>>>(k, a);
		}
	}
	else if(rt == 32) {
		if(k.channels == 2) k_emit3<32, MO, 2, false, false><<<nb, nt, smem, st>>>(k, a);
		else k_emit3<32, MO, 1, false, false><<<nb, nt, smem, st>>>(k, a);
	}
	else {
		if(k.channels == 2) k_emit3<36, MO, 2, false, false><<<nb, nt, smem, st>>>(k, a);
		else k_emit3<36, MO, 1, false, false><<<nb, nt, smem, st>>>(k, a);
	}
}

void launch_emit3(const EncK &k, int rt, int maxord_t, size_t smem, const Emit3Args &a, int nb, cudaStream_t st)
{
	if(maxord_t == 8) emit3<8>(k, rt, smem, a, nb, st);
	else if(maxord_t == 12) emit3<12>(k, rt, smem, a, nb, st);
	else emit3<32>(k, rt, smem, a, nb, st);
}

// more than two channels: channel pairs into staging regions, then one CTA per frame splices them
template <int MO>
static void emit3_pairs(const EncK &k, int rt, size_t smem, const Emit3Args &a, int nb, cudaStream_t st)
{
	const int nt = (k.bs / rt) * 2, grid = nb * ((k.channels + 1) / 2);
	if(k.bps > 16) {
		if(rt == 32) k_emit3<32, MO, 2, true, true><<<grid, nt, smem, st>>>(k, a);
		else k_emit3<36, MO, 2, true, true><<<grid, nt, smem, st>>>(k, a);
	}
	else if(rt == 32) k_emit3<32, MO, 2, false, true><<<grid, nt, smem, st>>>(k, a);
	else k_emit3<36, MO, 2, false, true><<<grid, nt, smem, st>>>(k, a);
}

size_t join_smem(int join_words) { return join_smem_bytes(join_words); }

void launch_emit3_pairs(const EncK &k, int rt, int maxord_t, size_t pair_smem, size_t join_smem_b, const Emit3Args &a, int nb, cudaStream_t st)
{
	if(maxord_t == 8) emit3_pairs<8>(k, rt, pair_smem, a, nb, st);
	else if(maxord_t == 12) emit3_pairs<12>(k, rt, pair_smem, a, nb, st);
	else emit3_pairs<32>(k, rt, pair_smem, a, nb, st);
	k_join<<<nb, 256, join_smem_b, st>>>(k, a);
}

void launch_crc16_tables(uint16_t *tab, cudaStream_t st) { k_crc16_tables<<<1, 256, 0, st>>>(tab); }

template <int MO>
static void emit3_attrs()
{
	cudaFuncSetAttribute(k_emit3<32, MO, 2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<32, MO, 1, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<36, MO, 2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<36, MO, 1, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<32, MO, 2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<32, MO, 1, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<36, MO, 2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<36, MO, 1, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<32, MO, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<36, MO, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<32, MO, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<36, MO, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

void emit3_init(int)
{
	cudaFuncSetAttribute(k_join, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	emit3_attrs<8>();
	emit3_attrs<12>();
	emit3_attrs<32>();
}

}  // namespace fb200
