// Stand-alone probe of the 3-D u8 TMA tile load used by k_fast_cells (unaligned start coordinates, box 96 x H x 1).
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_probe tools/tma_probe.cu && ./tools/tma_probe
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define BW 96
#define BH 48
struct Maps { CUtensorMap m[2]; int box_h[2]; };

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
	asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar)
{
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
	             ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, unsigned parity)
{
	const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
	unsigned done = 0;
	for (int spin = 0; spin < (1 << 20) && !done; spin++)
		asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(a), "r"(parity) : "memory");
	return done != 0;
}

__global__ void k(const __grid_constant__ Maps maps, int which, int x0, int y0, int f, uint8_t* out, int* flag)
{
	__shared__ __align__(128) uint8_t tile[BW * BH];
	__shared__ __align__(8) uint64_t bar;
	if (threadIdx.x == 0)
	{
		mbar_init(&bar, 1);
		mbar_expect_tx(&bar, BW * maps.box_h[which]);
		tma_load_3d(tile, &maps.m[which], x0, y0, f, &bar);
	}
	__syncthreads();
	const bool ok = mbar_wait(&bar, 0);
	if (threadIdx.x == 0) *flag = ok ? 1 : -1;
	if (ok) for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}

int main()
{
	const int pitch = 640, h = 480, frames = 3;
	std::vector<uint8_t> img((size_t)pitch * h * frames);
	for (size_t i = 0; i < img.size(); i++) img[i] = (uint8_t)((i * 2654435761u) >> 13);
	uint8_t* d; cudaMalloc(&d, img.size() + 512); d += 256;
	cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
	typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
	                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	void* fn = nullptr; cudaDriverEntryPointQueryResult q;
	cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
	Maps maps; memset(&maps, 0, sizeof(maps));
	const cuuint64_t dims[3] = { (cuuint64_t)pitch, (cuuint64_t)h, (cuuint64_t)frames };
	const cuuint64_t strides[2] = { (cuuint64_t)pitch, (cuuint64_t)pitch * h };
	const cuuint32_t box[3] = { BW, BH, 1 }, es[3] = { 1, 1, 1 };
	for (int w = 0; w < 2; w++)
	{
		CUresult r = ((EncodeFn)fn)(&maps.m[w], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
		                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		printf("encode %d -> %d\n", w, (int)r);
		maps.box_h[w] = BH;
	}
	uint8_t* dout; int* dflag; cudaMalloc(&dout, BW * BH); cudaMalloc(&dflag, 4);
	const int tests[][4] = { {0, 16, 16, 0}, {1, 16, 16, 1}, {0, 19, 47, 2}, {1, 601, 450, 1}, {0, 47, 16, 0} };
	for (auto& t : tests)
	{
		cudaMemset(dflag, 0, 4);
		k<<<1, 128>>>(maps, t[0], t[1], t[2], t[3], dout, dflag);
		cudaError_t e = cudaDeviceSynchronize();
		int flag = 0; cudaMemcpy(&flag, dflag, 4, cudaMemcpyDeviceToHost);
		std::vector<uint8_t> out(BW * BH); cudaMemcpy(out.data(), dout, out.size(), cudaMemcpyDeviceToHost);
		int bad = 0;
		for (int y = 0; y < BH; y++) for (int x = 0; x < BW; x++)
		{
			const int gx = t[1] + x, gy = t[2] + y;
			const uint8_t want = (gx < pitch && gy < h) ? img[(size_t)t[3] * pitch * h + (size_t)gy * pitch + gx] : 0;
			bad += out[y * BW + x] != want;
		}
		printf("map %d x0 %d y0 %d f %d: err=%s flag=%d mismatches=%d\n", t[0], t[1], t[2], t[3], cudaGetErrorString(e), flag, bad);
		if (e != cudaSuccess) break;
	}
	return 0;
}
