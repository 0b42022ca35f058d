// orbx_knn2_sharded driven the way a C++ host would: one thread per GPU, one ncclComm_t per rank from ncclCommInitAll,
// each rank scans its contiguous slice of the train set; every rank's merged result must equal the single-GPU scan (orbx_knn2) of the
// whole set. With one GPU the program runs the nranks = 1 path (no collective). usage: knn_sharded_test [nranks] [nq] [nt]
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "orbx.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 11); }

int main(int argc, char** argv)
{
	int ndev = 0;
	CK(cudaGetDeviceCount(&ndev));
	int nranks = argc > 1 ? atoi(argv[1]) : (ndev >= 2 ? 2 : 1);
	if (nranks > ndev) { printf("SKIP %d ranks need %d GPUs, %d present\n", nranks, nranks, ndev); return 0; }
	const int64_t nq = argc > 2 ? atoll(argv[2]) : 3000, nt = argc > 3 ? atoll(argv[3]) : 200003;
	std::vector<uint8_t> q((size_t)nq * 32), t((size_t)nt * 32);
	for (auto& b : q) b = (uint8_t)rnd();
	for (auto& b : t) b = (uint8_t)rnd();
	// plant near-duplicates (accepted matches), exact duplicates at two indices in DIFFERENT shards (lowest index must win) and ties
	for (int64_t i = 0; i < nq; i += 3)
	{
		const int64_t j = (int64_t)(rnd() % (uint32_t)nt);
		memcpy(&t[(size_t)j * 32], &q[(size_t)i * 32], 32);
		for (int f = 0; f < (int)(rnd() % 40); f++) t[(size_t)j * 32 + rnd() % 32] ^= (uint8_t)(1u << (rnd() % 8));
		if (i % 9 == 0) memcpy(&t[(size_t)((j + nt / 2) % nt) * 32], &t[(size_t)j * 32], 32);
	}
	std::vector<int32_t> idx0(nq), match0(nq);
	std::vector<uint16_t> best0(nq), second0(nq);
	if (orbx_knn2(0, q.data(), nq, t.data(), nt, 50, 0.6f, idx0.data(), best0.data(), second0.data(), match0.data()) != ORBX_OK)
	{ printf("FAIL orbx_knn2: %s\n", orbx_last_error()); return 1; }

	std::vector<ncclComm_t> comms(nranks);
	std::vector<int> devs(nranks);
	for (int r = 0; r < nranks; r++) devs[r] = r;
	if (nranks > 1 && ncclCommInitAll(comms.data(), nranks, devs.data()) != ncclSuccess) { printf("FAIL ncclCommInitAll\n"); return 1; }
	std::vector<int> bad(nranks, 0);
	std::vector<int> accepted(nranks, 0);
	auto work = [&](int r) {
		CK(cudaSetDevice(r));
		cudaStream_t st;
		CK(cudaStreamCreate(&st));
		const int64_t lo = nt * r / nranks, hi = nt * (r + 1) / nranks;      // contiguous shards
		uint8_t *dq, *dt; int32_t *di, *dm; uint16_t *db, *ds;
		CK(cudaMalloc(&dq, (size_t)nq * 32)); CK(cudaMalloc(&dt, (size_t)(hi - lo) * 32 + 32));
		CK(cudaMalloc(&di, nq * 4)); CK(cudaMalloc(&dm, nq * 4)); CK(cudaMalloc(&db, nq * 2)); CK(cudaMalloc(&ds, nq * 2));
		CK(cudaMemcpyAsync(dq, q.data(), (size_t)nq * 32, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(dt, t.data() + (size_t)lo * 32, (size_t)(hi - lo) * 32, cudaMemcpyHostToDevice, st));
		const orbx_status s = orbx_knn2_sharded(nranks > 1 ? (void*)comms[r] : nullptr, r, nranks, dq, nq, dt, hi - lo, lo, 50, 0.6f, di, db, ds, dm,
		                                        nullptr, (void*)st);
		if (s != ORBX_OK) { printf("FAIL rank %d orbx_knn2_sharded: %s\n", r, orbx_last_error()); bad[r] = 1; return; }
		std::vector<int32_t> idx(nq), match(nq);
		std::vector<uint16_t> best(nq), second(nq);
		CK(cudaMemcpyAsync(idx.data(), di, nq * 4, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(match.data(), dm, nq * 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(best.data(), db, nq * 2, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(second.data(), ds, nq * 2, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		for (int64_t i = 0; i < nq; i++)
		{
			if (idx[i] != idx0[i] || best[i] != best0[i] || second[i] != second0[i] || match[i] != match0[i])
			{
				if (!bad[r]) printf("FAIL rank %d query %lld: (%d %d %d %d) vs single scan (%d %d %d %d)\n", r, (long long)i, idx[i], best[i], second[i], match[i],
				                    idx0[i], best0[i], second0[i], match0[i]);
				bad[r]++;
			}
			accepted[r] += match[i] >= 0;
		}
		cudaFree(dq); cudaFree(dt); cudaFree(di); cudaFree(dm); cudaFree(db); cudaFree(ds);
		cudaStreamDestroy(st);
	};
	std::vector<std::thread> th;
	for (int r = 0; r < nranks; r++) th.emplace_back(work, r);
	for (auto& x : th) x.join();
	for (int r = 0; r < nranks; r++) if (nranks > 1) ncclCommDestroy(comms[r]);
	int nbad = 0;
	for (int r = 0; r < nranks; r++) nbad += bad[r];
	if (nbad) { printf("FAIL %d mismatches\n", nbad); return 1; }
	printf("OK %d rank(s), %lld queries x %lld train rows, %d accepted matches, identical to the single-GPU scan on every rank\n", nranks, (long long)nq,
	       (long long)nt, accepted[0]);
	return 0;
}
