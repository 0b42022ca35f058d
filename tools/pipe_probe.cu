// Throughput probe for the integer instructions the FAST / matcher kernels lean on (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_probe tools/pipe_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define CHAINS 8

template <int OP> __device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c)
{
	if (OP == 0) return __vimax3_u16x2(a, b, c);          // VIMNMX3.U16x2
	if (OP == 1) return __vmaxu2(a, b);                   // VIMNMX.U16x2
	if (OP == 2) return max(a, b);                        // VIMNMX.U32 (2-input)
	if (OP == 3) return max(max(a, b), c);                // VIMNMX3.U32 ?
	if (OP == 4) return __byte_perm(a, b, 0x5140);        // PRMT
	if (OP == 5) return a * 0xFFFF0001u + b;              // IMAD
	if (OP == 6) return (a & b) ^ c;                      // LOP3
	if (OP == 7) return __popc(a) + b;                    // POPC (+IADD)
	if (OP == 8) return __dp4a(a, b, c);                  // IDP.4A
	if (OP == 9) return a + b + c;                        // IADD3
	if (OP == 10) return __vimax3_s32(a, b, c);           // VIMNMX3.S32
	if (OP == 11) return __vabsdiffu4(a, b);              // VABSDIFF4
	if (OP == 12) return __funnelshift_r(a, b, 8);        // SHF
	if (OP == 13) return __viaddmax_u16x2(a, b, c);       // VIADDMNMX.U16x2
	if (OP == 14) return __vimin_s32_relu((int)a, (int)b);
	return a;
}

template <int OP> __global__ void __launch_bounds__(256) k(uint32_t* out)
{
	uint32_t x[CHAINS], y = threadIdx.x * 2654435761u + 12345u, z = blockIdx.x * 40503u + 77u;
#pragma unroll
	for (int i = 0; i < CHAINS; i++) x[i] = y + i * 0x01010101u;
	for (int it = 0; it < ITERS; it++)
	{
#pragma unroll
		for (int i = 0; i < CHAINS; i++) x[i] = op<OP>(x[i], y, z);
		y += 0x00010001u;   // keeps the compiler from hoisting; one extra IADD per CHAINS ops
	}
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < CHAINS; i++) s ^= x[i];
	if (s == 0x12345678u) out[0] = s;
}

template <int OP> void run(const char* name, int sms, uint32_t* d)
{
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int blocks = sms * 8;
	k<OP><<<blocks, 256>>>(d); cudaDeviceSynchronize();
	cudaEventRecord(e0); k<OP><<<blocks, 256>>>(d); cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	const double ops = (double)blocks * 256 * ITERS * CHAINS;
	printf("%-18s %8.1f Gop/s  %6.1f lane-ops/clk/SM (at %d MHz nominal)\n", name, ops / ms / 1e6, ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
}

int main()
{
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	uint32_t* d; cudaMalloc(&d, 4);
	printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
	run<0>("VIMNMX3.U16x2", p.multiProcessorCount, d);
	run<1>("VIMNMX.U16x2", p.multiProcessorCount, d);
	run<2>("VIMNMX.U32", p.multiProcessorCount, d);
	run<3>("max3 u32", p.multiProcessorCount, d);
	run<10>("VIMNMX3.S32", p.multiProcessorCount, d);
	run<13>("VIADDMNMX.U16x2", p.multiProcessorCount, d);
	run<14>("vimin_s32_relu", p.multiProcessorCount, d);
	run<4>("PRMT", p.multiProcessorCount, d);
	run<5>("IMAD", p.multiProcessorCount, d);
	run<6>("LOP3", p.multiProcessorCount, d);
	run<9>("IADD3", p.multiProcessorCount, d);
	run<12>("SHF", p.multiProcessorCount, d);
	run<7>("POPC+IADD", p.multiProcessorCount, d);
	run<8>("IDP.4A", p.multiProcessorCount, d);
	run<11>("VABSDIFF4", p.multiProcessorCount, d);
	return 0;
}
