// TEST INFRASTRUCTURE — CPU oracle, not the product.
#pragma once
#include <algorithm>
#include <vector>

namespace stdsort {
// ---------------------------------------------------------------------------------------------
// libstdc++ (GCC 13) std::sort for the comparator "size descending" (:642-643), restated because
// its order among EQUAL sizes decides which nodes get split before the quota break (:666-667).
// SURVEY App. E; /usr/include/c++/13/bits/stl_algo.h:85-104,1792-1950.
// ---------------------------------------------------------------------------------------------
struct SortItem { int size; int node; };
inline bool before(const SortItem& a, const SortItem& b) { return a.size > b.size; }

inline void median_to_first(SortItem* r, SortItem* a, SortItem* b, SortItem* c)
{
	if (before(*a, *b))
	{
		if (before(*b, *c)) std::swap(*r, *b);
		else if (before(*a, *c)) std::swap(*r, *c);
		else std::swap(*r, *a);
	}
	else if (before(*a, *c)) std::swap(*r, *a);
	else if (before(*b, *c)) std::swap(*r, *c);
	else std::swap(*r, *b);
}

inline SortItem* partition_unguarded(SortItem* first, SortItem* last, SortItem* pivot)
{
	for (;;)
	{
		while (before(*first, *pivot)) ++first;
		--last;
		while (before(*pivot, *last)) --last;
		if (!(first < last)) return first;
		std::swap(*first, *last);
		++first;
	}
}

inline void sift_down(SortItem* a, int hole, int len, SortItem v)
{
	// __adjust_heap + __push_heap (stl_heap.h) for the heapsort fallback
	const int top = hole;
	int child = hole;
	while (child < (len - 1) / 2)
	{
		child = 2 * (child + 1);
		if (before(a[child], a[child - 1])) child--;
		a[hole] = a[child];
		hole = child;
	}
	if ((len & 1) == 0 && child == (len - 2) / 2)
	{
		child = 2 * (child + 1);
		a[hole] = a[child - 1];
		hole = child - 1;
	}
	int parent = (hole - 1) / 2;
	while (hole > top && before(a[parent], v))
	{
		a[hole] = a[parent];
		hole = parent;
		parent = (hole - 1) / 2;
	}
	a[hole] = v;
}

inline void heap_sort(SortItem* first, SortItem* last)
{
	// std::partial_sort(first, last, last): make_heap then sort_heap
	const int len = (int)(last - first);
	if (len < 2) return;
	for (int parent = (len - 2) / 2;; parent--)
	{
		sift_down(first, parent, len, first[parent]);
		if (parent == 0) break;
	}
	for (SortItem* end = last; end - first > 1;)
	{
		--end;
		SortItem v = *end;
		*end = *first;
		sift_down(first, 0, (int)(end - first), v);
	}
}

inline void intro_loop(SortItem* first, SortItem* last, int depth)
{
	while (last - first > 16)
	{
		if (depth == 0) { heap_sort(first, last); return; }
		--depth;
		SortItem* mid = first + (last - first) / 2;
		median_to_first(first, first + 1, mid, last - 1);
		SortItem* cut = partition_unguarded(first + 1, last, first);
		intro_loop(cut, last, depth);
		last = cut;
	}
}

inline void linear_insert_unguarded(SortItem* last)
{
	SortItem v = *last;
	SortItem* next = last - 1;
	while (before(v, *next)) { *last = *next; last = next; --next; }
	*last = v;
}

inline void insertion(SortItem* first, SortItem* last)
{
	if (first == last) return;
	for (SortItem* i = first + 1; i != last; ++i)
	{
		if (before(*i, *first))
		{
			SortItem v = *i;
			for (SortItem* p = i; p != first; --p) *p = *(p - 1);
			*first = v;
		}
		else linear_insert_unguarded(i);
	}
}

inline void libstdcxx_sort_desc(std::vector<SortItem>& a)
{
	const int n = (int)a.size();
	if (n == 0) return;
	int lg = 0;
	for (int m = n; m > 1; m >>= 1) ++lg;
	SortItem* first = a.data();
	SortItem* last = first + n;
	intro_loop(first, last, 2 * lg);
	if (n > 16)
	{
		insertion(first, first + 16);
		for (SortItem* i = first + 16; i != last; ++i) linear_insert_unguarded(i);
	}
	else insertion(first, last);
}
}  // namespace stdsort
