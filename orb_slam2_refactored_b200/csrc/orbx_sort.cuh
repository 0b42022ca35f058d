// libstdc++ std::sort replayed on the device (shared by the quadtree and the guided matchers' CheckOrientation).
#pragma once
#include <stdint.h>

// ---- libstdc++ (GCC 13) std::sort with comp(a,b) = a.size > b.size, on 64-bit items (size << 32 | position)
// /usr/include/c++/13/bits/stl_algo.h:85-104, 1792-1950; SURVEY App. E. Single thread, shared memory.
__device__ __forceinline__ bool qs_before(uint64_t a, uint64_t b) { return (uint32_t)(a >> 32) > (uint32_t)(b >> 32); }
__device__ __forceinline__ void qs_swap(uint64_t* a, int i, int j) { const uint64_t t = a[i]; a[i] = a[j]; a[j] = t; }

__device__ inline void qs_sift(uint64_t* a, int first, int hole, int len, uint64_t v)
{
	const int top = hole;
	int child = hole;
	while (child < (len - 1) / 2)
	{
		child = 2 * (child + 1);
		if (qs_before(a[first + child], a[first + child - 1])) child--;
		a[first + hole] = a[first + child];
		hole = child;
	}
	if ((len & 1) == 0 && child == (len - 2) / 2)
	{
		child = 2 * (child + 1);
		a[first + hole] = a[first + child - 1];
		hole = child - 1;
	}
	int parent = (hole - 1) / 2;
	while (hole > top && qs_before(a[first + parent], v))
	{
		a[first + hole] = a[first + parent];
		hole = parent;
		parent = (hole - 1) / 2;
	}
	a[first + hole] = v;
}

__device__ inline void qs_heapsort(uint64_t* a, int first, int last)
{
	const int len = last - first;
	if (len < 2) return;
	for (int parent = (len - 2) / 2;; parent--)
	{
		qs_sift(a, first, parent, len, a[first + parent]);
		if (parent == 0) break;
	}
	for (int end = last; end - first > 1;)
	{
		--end;
		const uint64_t v = a[end];
		a[end] = a[first];
		qs_sift(a, first, 0, end - first, v);
	}
}

__device__ __forceinline__ void qs_linear_insert(uint64_t* a, int last)
{
	const uint64_t v = a[last];
	int next = last - 1;
	while (qs_before(v, a[next])) { a[last] = a[next]; last = next; --next; }
	a[last] = v;
}

__device__ inline void qs_insertion(uint64_t* a, int first, int last)
{
	if (first == last) return;
	for (int i = first + 1; i != last; ++i)
	{
		if (qs_before(a[i], a[first]))
		{
			const uint64_t v = a[i];
			for (int p = i; p != first; --p) a[p] = a[p - 1];
			a[first] = v;
		}
		else qs_linear_insert(a, i);
	}
}

// One partition step of __introsort_loop on [first, last) (last - first > 16): median of three to the front, unguarded
// partition around it. Returns the cut. Executed by a single thread.
__device__ inline int qs_partition(uint64_t* a, int first, int last)
{
	const int mid = first + (last - first) / 2;
	{   // __move_median_to_first(first, first+1, mid, last-1)
		const int r = first, x = first + 1, y = mid, z = last - 1;
		if (qs_before(a[x], a[y]))
		{
			if (qs_before(a[y], a[z])) qs_swap(a, r, y);
			else if (qs_before(a[x], a[z])) qs_swap(a, r, z);
			else qs_swap(a, r, x);
		}
		else if (qs_before(a[x], a[z])) qs_swap(a, r, x);
		else if (qs_before(a[y], a[z])) qs_swap(a, r, z);
		else qs_swap(a, r, y);
	}
	int lo = first + 1, hi = last;
	const uint64_t pivot_key = a[first];   // the pivot stays at a[first] during the partition
	for (;;)
	{
		while (qs_before(a[lo], pivot_key)) ++lo;
		--hi;
		while (qs_before(pivot_key, a[hi])) --hi;
		if (!(lo < hi)) break;
		qs_swap(a, lo, hi);
		++lo;
	}
	return lo;
}

// The same partition step by a whole warp, with the same result (array contents and cut). The serial loop swaps the k-th element from the
// left that is not before the pivot (a "left stopper") with the k-th element from the right that the pivot is not before (a "right
// stopper") for as long as the former lies left of the latter. Both scans only ever read positions no swap has touched yet (a swap writes
// at or below lo and at or above hi), so the stoppers can be taken from the array as it is BEFORE the first swap: with Ls ascending and Rs
// descending over the original array, the swaps are (Ls[k], Rs[k]) for k < m, m = #{k : Ls[k] < Rs[k]}, and the cut is min(Ls[m], Rs[m - 1])
// (the scan for the (m+1)-th left stopper also stops at the element the last swap moved to Rs[m - 1]). rs: scratch for Rs, one 16-bit entry
// per element of the range (positions < 65536), indexed from `first` (ranges of concurrent calls are disjoint). All 32 lanes call it.
__device__ __forceinline__ int qs_partition_warp(uint64_t* a, int first, int last, uint16_t* rs)
{
	const int lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1u;
	if (lane == 0)
	{
		const int mid = first + (last - first) / 2;
		const int r = first, x = first + 1, y = mid, z = last - 1;
		if (qs_before(a[x], a[y]))
		{
			if (qs_before(a[y], a[z])) qs_swap(a, r, y);
			else if (qs_before(a[x], a[z])) qs_swap(a, r, z);
			else qs_swap(a, r, x);
		}
		else if (qs_before(a[x], a[z])) qs_swap(a, r, x);
		else if (qs_before(a[y], a[z])) qs_swap(a, r, z);
		else qs_swap(a, r, y);
	}
	__syncwarp();
	const uint32_t p = (uint32_t)(a[first] >> 32);
	// right stoppers, descending: positions j in [first, last) with !before(pivot, a[j]) <=> size(a[j]) >= p; j = first (the pivot) is the last one
	int cntR = 0;
	for (int base = last - 1; base >= first; base -= 32)
	{
		const int j = base - lane;
		const bool f = j >= first && (uint32_t)(a[j] >> 32) >= p;
		const unsigned b = __ballot_sync(0xffffffffu, f);
		if (f) rs[first + cntR + __popc(b & lt)] = (uint16_t)j;
		cntR += __popc(b);
	}
	__syncwarp();
	// left stoppers, ascending: positions i in (first, last) with !before(a[i], pivot) <=> size(a[i]) <= p
	int cntL = 0, cut = -1, lastR = 0x7fffffff;
	for (int base = first + 1; base < last; base += 32)
	{
		const int i = base + lane;
		const uint64_t v = i < last ? a[i] : 0ull;
		const bool f = i < last && (uint32_t)(v >> 32) <= p;
		const unsigned b = __ballot_sync(0xffffffffu, f);
		const int k = cntL + __popc(b & lt);
		const int r = (f && k < cntR) ? (int)rs[first + k] : -1;
		const bool sw = f && i < r;
		uint64_t vr = 0ull;
		if (sw) vr = a[r];
		const unsigned bs = __ballot_sync(0xffffffffu, sw);      // also orders every read above before the writes below
		if (sw) { a[i] = vr; a[r] = v; }
		// the swaps form a prefix of the left stoppers: the first stopper that does not swap ends the loop
		const unsigned fail = b & ~bs;
		if (bs) lastR = (int)rs[first + cntL + __popc(bs) - 1];
		if (fail)
		{
			const int src = __ffs(fail) - 1;
			cut = min(__shfl_sync(0xffffffffu, i, src), lastR);
			break;
		}
		cntL += __popc(b);
	}
	__syncwarp();
	return cut >= 0 ? cut : lastR;
}

// std::sort by one thread: __introsort_loop with an explicit stack, then __final_insertion_sort. Used for small inputs,
// where the rounds of the block-parallel version below cost more barriers than they save.
__device__ __noinline__ inline void qs_sort_serial(uint64_t* a, int n)
{
	if (n == 0) return;
	int lg = 0;
	for (int m = n; m > 1; m >>= 1) ++lg;
	int st_first[48], st_last[48], st_depth[48];
	int sp = 0;
	st_first[0] = 0; st_last[0] = n; st_depth[0] = 2 * lg; sp = 1;
	while (sp > 0)
	{
		--sp;
		int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
		while (last - first > 16)
		{
			if (depth == 0) { qs_heapsort(a, first, last); break; }
			--depth;
			const int cut = qs_partition(a, first, last);
			// the right part is sorted "recursively" before the left part continues; the ranges are disjoint, so deferring it
			// on the stack yields the same result
			st_first[sp] = cut; st_last[sp] = last; st_depth[sp] = depth; ++sp;
			last = cut;
		}
	}
	if (n > 16)
	{
		qs_insertion(a, 0, 16);
		for (int i = 16; i != n; ++i) qs_linear_insert(a, i);
	}
	else qs_insertion(a, 0, n);
}

