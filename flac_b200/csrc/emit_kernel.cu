// emit_kernel.cu -- k_emit3 and its launcher.
#include "emit_kernel.cuh"

namespace fb200 {

template <int MO>
static void emit3(const EncK &k, int rt, size_t smem, const Emit3Args &a, int nb, cudaStream_t st)
{
	const int nt = (k.bs / rt) * k.channels;
	if(k.bps > 16) {
		if(rt == 32) {
			if(k.channels == 2) k_emit3<32, MO, 2, true><<<nb, nt, smem, st>>>(k, a);
			else k_emit3<32, MO, 1, true><<<nb, nt, smem, st>>>(k, a);
		}
		else {
			if(k.channels == 2) k_emit3<36, MO, 2, true><<<nb, nt, smem, st>>>(k, a);
			else k_emit3<36, MO, 1, true><<<nb, nt, smem, st>>>(k, a);
		}
	}
	else if(rt == 32) {
		if(k.channels == 2) k_emit3<32, MO, 2, false><<<nb, nt, smem, st>>>(k, a);
		else k_emit3<32, MO, 1, false><<<nb, nt, smem, st>>>(k, a);
	}
	else {
		if(k.channels == 2) k_emit3<36, MO, 2, false><<<nb, nt, smem, st>>>(k, a);
		else k_emit3<36, MO, 1, false><<<nb, nt, smem, st>>>(k, a);
	}
}

void launch_emit3(const EncK &k, int rt, int maxord_t, size_t smem, const Emit3Args &a, int nb, cudaStream_t st)
{
	if(maxord_t == 8) emit3<8>(k, rt, smem, a, nb, st);
	else if(maxord_t == 12) emit3<12>(k, rt, smem, a, nb, st);
	else emit3<32>(k, rt, smem, a, nb, st);
}

void launch_crc16_tables(uint16_t *tab, cudaStream_t st) { k_crc16_tables<<<1, 256, 0, st>>>(tab); }

template <int MO>
static void emit3_attrs()
{
	cudaFuncSetAttribute(k_emit3<32, MO, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<32, MO, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<36, MO, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<36, MO, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<32, MO, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<32, MO, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<36, MO, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaFuncSetAttribute(k_emit3<36, MO, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

void emit3_init(int)
{
	emit3_attrs<8>();
	emit3_attrs<12>();
	emit3_attrs<32>();
}

}  // namespace fb200
