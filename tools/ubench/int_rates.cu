// Micro-benchmark: integer pipe rates on this GPU (inputs to k_search3's issue-bound floor).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP, int CHAINS>
__global__ void k_tput(int *out, const int *in, int iters)
{
	int a[CHAINS];
	long long w[CHAINS];
	const int b = in[threadIdx.x & 7], c = in[8 + (threadIdx.x & 7)];
#pragma unroll
	for(int i = 0; i < CHAINS; i++) { a[i] = in[i] + threadIdx.x; w[i] = in[i] * 77 + threadIdx.x; }
	for(int it = 0; it < iters; it++) {
#pragma unroll
		for(int i = 0; i < CHAINS; i++) {
			if(OP == 0) a[i] = a[i] * b + c;                                   // IMAD
			else if(OP == 1) a[i] = a[i] + b;                                  // IADD3 (may be optimised: keep dependent)
			else if(OP == 2) a[i] = (int)__sad(a[i], b, (unsigned)c);          // VABSDIFF
			else if(OP == 3) a[i] = (a[i] >> (b & 31)) ^ c;                    // SHF + LOP3
			else if(OP == 4) w[i] = (long long)(int)w[i] * (long long)b + w[i];// IMAD.WIDE
			else if(OP == 5) a[i] = __funnelshift_r(a[i], c, b);               // SHF
			else if(OP == 6) a[i] = (a[i] & b) | c;                            // LOP3
		}
	}
	int s = 0;
#pragma unroll
	for(int i = 0; i < CHAINS; i++) s += a[i] + (int)w[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP, int CHAINS>
static void run(const char *name, double ops_per, int blocks, int threads, int *d_out, int *d_in, int sms, double mhz)
{
	const int iters = 4096;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	k_tput<OP, CHAINS><<<blocks, threads>>>(d_out, d_in, iters);
	cudaEventRecord(e0);
	k_tput<OP, CHAINS><<<blocks, threads>>>(d_out, d_in, iters);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	const double ops = (double)blocks * threads * iters * CHAINS * ops_per;
	printf("%-28s chains=%d grid=%dx%d  %.3f ms  %.1f thread-ops/clk/SM\n", name, CHAINS, blocks, threads, ms, ops / (ms * 1e-3) / (mhz * 1e6) / sms);
}

int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	int khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	const double mhz = khz / 1000.0;
	const int sms = p.multiProcessorCount;
	printf("%s  SMs=%d  clock=%.0f MHz\n", p.name, sms, mhz);
	int *d_out, *d_in;
	cudaMalloc(&d_out, sizeof(int) * sms * 8 * 1024);
	int h[16]; for(int i = 0; i < 16; i++) h[i] = 3 + 2 * i;
	cudaMalloc(&d_in, sizeof h); cudaMemcpy(d_in, h, sizeof h, cudaMemcpyHostToDevice);
	run<0, 8>("IMAD", 1, sms * 8, 256, d_out, d_in, sms, mhz);
	run<1, 8>("IADD", 1, sms * 8, 256, d_out, d_in, sms, mhz);
	run<2, 8>("VABSDIFF (sad)", 1, sms * 8, 256, d_out, d_in, sms, mhz);
	run<3, 8>("SHF+LOP3 (2 ops)", 2, sms * 8, 256, d_out, d_in, sms, mhz);
	run<4, 8>("IMAD.WIDE", 1, sms * 8, 256, d_out, d_in, sms, mhz);
	run<5, 8>("SHF", 1, sms * 8, 256, d_out, d_in, sms, mhz);
	run<6, 8>("LOP3", 1, sms * 8, 256, d_out, d_in, sms, mhz);
	run<0, 8>("IMAD 2 warps/SMSP", 1, sms, 256, d_out, d_in, sms, mhz);
	run<0, 12>("IMAD 12 chains 2 warps/SMSP", 1, sms, 256, d_out, d_in, sms, mhz);
	run<0, 1>("IMAD dependent 1 warp/SM", 1, sms, 32, d_out, d_in, sms, mhz);
	cudaDeviceSynchronize();
	printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
