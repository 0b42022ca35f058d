// Issue-rate probe for sm_100a: which SASS ops share a pipe. For every op A alone and every pair (A, B) interleaved 1:1 it prints
// lane-ops per clock per SM (clock64-based, so independent of the boost state). If A+B interleaved reaches rate(A) + rate(B) they
// sit on different pipes; if the pair only reaches the harmonic combination they share one.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_probe2 tools/pipe_probe2.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define ITERS 1024
#define CH 6

enum Op { VMAX2, VMAX3, VMAXU32, PRMT, LOP3, IADD3, SHF, LEA, IMAD, IMADHI, IMADSHL, IDP4A, IDP2A, HMAX2, HADD2, HFMA2, HFMA2RELU, FMNMX, FADD, FFMA,
          ISETPSEL, POPC, VABSDIFF4, VIADDMAX, I2F, F2I, LDS32, LDS128, SHFL, VOTE, NOP_, NOPS };
static const char* kName[] = { "VIMNMX.U16x2", "VIMNMX3.U16x2", "VIMNMX.U32", "PRMT", "LOP3", "IADD3", "SHF", "LEA", "IMAD", "IMAD.HI", "IMAD.SHL", "IDP.4A", "IDP.2A",
                               "HMNMX2", "HADD2", "HFMA2", "HFMA2.RELU", "FMNMX", "FADD", "FFMA", "ISETP+SEL", "POPC", "VABSDIFF4", "VIADDMNMX", "I2F", "F2I",
                               "LDS.32", "LDS.128", "SHFL", "VOTE", "none" };

template <int OP> __device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c, const uint32_t* sm)
{
	if (OP == VMAX2) return __vmaxu2(a, b);
	if (OP == VMAX3) return __vimax3_u16x2(a, b, c);
	if (OP == VMAXU32) return max(a, b);
	if (OP == PRMT) return __byte_perm(a, b, 0x5140);
	if (OP == LOP3) return (a & b) ^ c;
	if (OP == IADD3) return a + b + c;
	if (OP == SHF) return __funnelshift_r(a, b, 8);
	if (OP == LEA) return (a << 3) + b;
	if (OP == IMAD) return a * b + c;
	if (OP == IMADHI) return __umulhi(a, b) + c;
	if (OP == IMADSHL) return a * 65536u + b;
	if (OP == IDP4A) return __dp4a(a, b, c);
	if (OP == IDP2A) return __dp2a_lo(a, b, c);
	if (OP == HMAX2) { __half2 x = *reinterpret_cast<__half2*>(&a), y = *reinterpret_cast<__half2*>(&b); __half2 r = __hmax2(x, y); return *reinterpret_cast<uint32_t*>(&r); }
	if (OP == HADD2) { __half2 x = *reinterpret_cast<__half2*>(&a), y = *reinterpret_cast<__half2*>(&b); __half2 r = __hadd2(x, y); return *reinterpret_cast<uint32_t*>(&r); }
	if (OP == HFMA2) { __half2 x = *reinterpret_cast<__half2*>(&a), y = *reinterpret_cast<__half2*>(&b), z = *reinterpret_cast<__half2*>(&c); __half2 r = __hfma2(x, y, z); return *reinterpret_cast<uint32_t*>(&r); }
	if (OP == HFMA2RELU) { __half2 x = *reinterpret_cast<__half2*>(&a), y = *reinterpret_cast<__half2*>(&b), z = *reinterpret_cast<__half2*>(&c); __half2 r = __hfma2_relu(x, y, z); return *reinterpret_cast<uint32_t*>(&r); }
	if (OP == FMNMX) return __float_as_uint(fmaxf(__uint_as_float(a), __uint_as_float(b)));
	if (OP == FADD) return __float_as_uint(__fadd_rn(__uint_as_float(a), __uint_as_float(b)));
	if (OP == FFMA) return __float_as_uint(__fmaf_rn(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c)));
	if (OP == ISETPSEL) return a > b ? c : a;
	if (OP == POPC) return __popc(a) + b;
	if (OP == VABSDIFF4) return __vabsdiffu4(a, b);
	if (OP == VIADDMAX) return __viaddmax_u16x2(a, b, c);
	if (OP == I2F) return __float_as_uint((float)(int)a) ^ b;
	if (OP == F2I) return (uint32_t)__float2int_rn(__uint_as_float(a)) ^ b;
	if (OP == LDS32) return sm[(a & 1023u)] + b;
	if (OP == LDS128) { const uint4 v = reinterpret_cast<const uint4*>(sm)[a & 255u]; return v.x + v.y + v.z + v.w; }
	if (OP == SHFL) return __shfl_xor_sync(0xffffffffu, a, 1) + b;
	if (OP == VOTE) return __ballot_sync(0xffffffffu, a > b) + c;
	return a;
}

template <int A, int B> __global__ void __launch_bounds__(256) k(uint32_t* out, unsigned long long* cyc)
{
	__shared__ uint32_t sm[1024];
	for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = i * 2654435761u;
	__syncthreads();
	uint32_t x[CH], w[CH], y = threadIdx.x * 2654435761u + 12345u, z = blockIdx.x * 40503u + 77u;
#pragma unroll
	for (int i = 0; i < CH; i++) { x[i] = y + i * 0x01010101u; w[i] = z + i * 0x00070003u; }
	const long long t0 = clock64();
	for (int it = 0; it < ITERS; it++)
	{
#pragma unroll
		for (int i = 0; i < CH; i++)
		{
			x[i] = op<A>(x[i], x[(i + 1) % CH], x[(i + 2) % CH], sm);      // operands from the other chains: nothing is idempotent or foldable
			if (B != NOP_) w[i] = op<B>(w[i], w[(i + 1) % CH], w[(i + 2) % CH], sm);
		}
		y += 0x00010001u;
	}
	const long long t1 = clock64();
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < CH; i++) s ^= x[i] ^ w[i];
	if (s == 0x12345678u) out[0] = s;
	if (threadIdx.x == 0) atomicMax(cyc, (unsigned long long)(t1 - t0));
}

static int g_sms;
static uint32_t* g_d;
static unsigned long long* g_c;

template <int A, int B> double run()
{
	const int blocks = g_sms * 8;                // 8 CTAs x 8 warps = 64 warps per SM: all resident at once
	k<A, B><<<blocks, 256>>>(g_d, g_c);
	cudaDeviceSynchronize();
	cudaMemset(g_c, 0, 8);
	k<A, B><<<blocks, 256>>>(g_d, g_c);
	cudaDeviceSynchronize();
	unsigned long long c = 0;
	cudaMemcpy(&c, g_c, 8, cudaMemcpyDeviceToHost);
	const double ops = 8.0 * 256 * ITERS * CH * (B == NOP_ ? 1 : 2);       // lane-ops per SM
	return ops / (double)c;
}

template <int A> void single() { printf("%-14s alone: %6.1f lane-ops/clk/SM\n", kName[A], run<A, NOP_>()); }
template <int A, int B> void pair()
{
	const double a = run<A, NOP_>(), b = run<B, NOP_>(), ab = run<A, B>();
	printf("%-14s + %-14s: %6.1f (alone %6.1f / %6.1f; shared-pipe prediction %6.1f, separate %6.1f)\n", kName[A], kName[B], ab, a, b,
	       2.0 / (1.0 / a + 1.0 / b), 2.0 * (a < b ? a : b));
}

int main()
{
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	g_sms = p.multiProcessorCount;
	cudaMalloc(&g_d, 4); cudaMalloc(&g_c, 8);
	printf("%s, %d SMs\n", p.name, g_sms);
	single<VMAX2>(); single<VMAX3>(); single<VMAXU32>(); single<PRMT>(); single<LOP3>(); single<IADD3>(); single<SHF>(); single<LEA>();
	single<IMAD>(); single<IMADHI>(); single<IMADSHL>(); single<IDP4A>(); single<IDP2A>(); single<HMAX2>(); single<HADD2>(); single<HFMA2>();
	single<HFMA2RELU>(); single<FMNMX>(); single<FADD>(); single<FFMA>(); single<ISETPSEL>(); single<POPC>(); single<VABSDIFF4>();
	single<VIADDMAX>(); single<I2F>(); single<F2I>(); single<LDS32>(); single<LDS128>(); single<SHFL>(); single<VOTE>();
	pair<VMAX2, PRMT>(); pair<VMAX2, VMAX3>(); pair<VMAX2, IMAD>(); pair<VMAX2, HFMA2>(); pair<VMAX2, HMAX2>(); pair<VMAX2, FFMA>(); pair<VMAX2, IADD3>();
	pair<VMAX3, PRMT>(); pair<VMAX3, IMAD>(); pair<VMAX3, HMAX2>(); pair<VMAX3, HFMA2>(); pair<VMAX3, LDS32>();
	pair<PRMT, IMAD>(); pair<PRMT, HFMA2>(); pair<PRMT, HMAX2>(); pair<PRMT, IADD3>(); pair<PRMT, LOP3>(); pair<PRMT, IDP4A>(); pair<PRMT, LDS32>();
	pair<HMAX2, HFMA2>(); pair<HMAX2, IMAD>(); pair<HMAX2, FMNMX>(); pair<HFMA2, IMAD>(); pair<HFMA2, IDP4A>(); pair<HFMA2, FFMA>();
	pair<IDP4A, IMAD>(); pair<IDP4A, FFMA>(); pair<IDP4A, IADD3>(); pair<IDP4A, SHF>(); pair<IDP4A, IDP2A>(); pair<IDP4A, IMADSHL>(); pair<IDP4A, LEA>();
	pair<IADD3, IMAD>(); pair<IADD3, LOP3>(); pair<IADD3, FFMA>(); pair<IMADHI, IMAD>(); pair<IMADHI, IADD3>(); pair<SHF, LEA>(); pair<IMAD, FFMA>();
	pair<IDP4A, LDS32>(); pair<IDP4A, LDS128>(); pair<SHFL, IDP4A>(); pair<SHFL, LDS32>();
	return 0;
}
