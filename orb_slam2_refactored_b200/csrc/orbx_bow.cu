// Bag-of-words transform on sm_100a: what Frame::ComputeBoW / KeyFrame::ComputeBoW call.
//
//   TemplatedVocabulary::loadFromTextFile                      Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.cpp:21-90   host (load_text)
//   TemplatedVocabulary::transform(feature, id, w, nid, lup)   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1220-1262 k_bow_descend
//   FORB::distance                                             Thirdparty/DBoW2/DBoW2/FORB.cpp:79-98                  inside k_bow_descend
//   TemplatedVocabulary::transform(features, v, fv, levelsup)  Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1129-1197 k_bow_finalize
//   BowVector::addWeight / addIfNotExist / normalize           Thirdparty/DBoW2/DBoW2/BowVector.cpp:32-87             inside k_bow_finalize
//   FeatureVector::addFeature                                  Thirdparty/DBoW2/DBoW2/FeatureVector.cpp:30-44         inside k_bow_finalize
//   L1Scoring::score (the reference's vocabulary is L1_NORM)   Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:24-58         k_bow_score_l1
//
// The tree lives in HBM as CSR children lists + one 32-byte descriptor per node (k = 10, L = 6: 1.1 M nodes, 36 MB — it stays in the
// 126 MB L2 across frames). Eight lanes walk one feature down the tree: lane j holds word j of the feature and of each child, the
// per-child popcounts are reduced with shuffles two children at a time (16-bit fields), and the strict `d < best_d` of the reference
// keeps the first minimum in children order. The two std::maps the reference fills feature by feature are rebuilt by a sort:
// (word, feature) and (node, feature) keys through a shared-memory bitonic network, run heads by a block scan. The floating-point
// parts are order-dependent in the reference and are kept in its order: a word's value is w added once per feature (v += w), the
// norm is one sequential pass over the words in ascending id (one thread, FP64), then every value is divided by it.
#include "orbx_internal.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

namespace {

constexpr int BW_THREADS = 1024;
constexpr int BW_MAX_FEATURES = 16384;       // one CTA sorts a frame's keys in shared memory (16384 x 8 B = 128 KB)

struct VocabDev
{
	int k, L, scoring, weighting;
	int64_t nnodes;              // including the root (node 0)
	int64_t nwords;
	const int4* node_info;       // [nnodes]: first slot in child_ids, number of children, id of the first child when the children's ids are
	                             // consecutive (the order HKmeansStep creates them in) else -1, Node::word_id (0 for nodes that are not words, as
	                             // in the reference): one 16-byte load per level instead of a chain of dependent loads
	const int32_t* child_ids;    // [nnodes - 1], push_back order of Node::children
	const uint32_t* desc;        // [nnodes][8]
	const double* weight;        // [nnodes]: Node::weight
};

// ---- per-feature tree walk ------------------------------------------------------------------------------------------------------
// 8 lanes per feature. feat_word / feat_node / feat_w: [frames][cap]. A feature index >= n[frame] writes nothing.
__global__ void __launch_bounds__(256) k_bow_descend(const VocabDev V, const uint8_t* __restrict__ desc, const int32_t* __restrict__ n, const int cap,
                                                     const int levelsup, int32_t* __restrict__ feat_word, int32_t* __restrict__ feat_node,
                                                     double* __restrict__ feat_w)
{
	const int frame = blockIdx.y;
	const int nf = min(n[frame], cap);
	const int i = blockIdx.x * 32 + (threadIdx.x >> 3), sub = threadIdx.x & 7;
	if (i >= nf) return;         // whole 8-lane groups leave: the shuffles below name only the group's own lanes
	const unsigned gmask = 0xffu << (threadIdx.x & 24);
	const uint32_t fw = reinterpret_cast<const uint32_t*>(desc + ((int64_t)frame * cap + i) * 32)[sub];
	const int nid_level = V.L - levelsup;
	int nid = 0;                 // :1232: root when nid_level <= 0. A leaf shallower than nid_level leaves *nid unset in the reference (the
	                             // caller's variable is uninitialised there); this implementation reports node 0 for that case.
	int node = 0, level = 0;
	int word = 0;
	for (;;)
	{
		const int4 info = __ldg(V.node_info + node);
		const int cs = info.x, nch = info.y, first = info.z;
		word = info.w;
		if (nch == 0) break;     // Node::isLeaf()
		++level;
		uint32_t best = 0xffffffffu;
		for (int c = 0; c < nch; c += 2)
		{
			const bool two = c + 1 < nch;
			const int id0 = first >= 0 ? first + c : __ldg(V.child_ids + cs + c);
			const int id1 = !two ? id0 : (first >= 0 ? first + c + 1 : __ldg(V.child_ids + cs + c + 1));
			uint32_t d = (uint32_t)__popc(fw ^ __ldg(V.desc + (int64_t)id0 * 8 + sub)) | ((uint32_t)__popc(fw ^ __ldg(V.desc + (int64_t)id1 * 8 + sub)) << 16);
			d += __shfl_xor_sync(gmask, d, 1);
			d += __shfl_xor_sync(gmask, d, 2);
			d += __shfl_xor_sync(gmask, d, 4);
			const uint32_t d0 = d & 0xffffu, d1 = d >> 16;
			if (d0 < best) { best = d0; node = id0; }                    // strict: the first minimum in children order wins (:1249)
			if (two && d1 < best) { best = d1; node = id1; }
		}
		if (level == nid_level) nid = node;
	}
	if (sub == 0)
	{
		const int64_t o = (int64_t)frame * cap + i;
		feat_word[o] = word;
		feat_node[o] = nid;
		feat_w[o] = __ldg(V.weight + node);
	}
}

// ---- block helpers ------------------------------------------------------------------------------------------------------------------
__device__ void bitonic_sort(uint64_t* a, int p)
{
	for (int k = 2; k <= p; k <<= 1)
		for (int j = k >> 1; j > 0; j >>= 1)
		{
			for (int t = threadIdx.x; t < (p >> 1); t += BW_THREADS)
			{
				const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
				const uint64_t x = a[lo], y = a[hi];
				const bool up = (lo & k) == 0;
				if ((x > y) == up) { a[lo] = y; a[hi] = x; }
			}
			__syncthreads();
		}
}

// exclusive scan of one int per thread over the block; `total` = sum. s_w: 32 ints of shared scratch.
__device__ int block_exscan(int v, int* s_w, int& total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		const int t = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= d) inc += t;
	}
	__syncthreads();
	if (lane == 31) s_w[warp] = inc;
	__syncthreads();
	int base = inc - v;
	total = 0;
	for (int w = 0; w < BW_THREADS / 32; w++)
	{
		const int t = s_w[w];
		if (w < warp) base += t;
		total += t;
	}
	return base;
}

// ---- per-frame BowVector + FeatureVector --------------------------------------------------------------------------------------------
// One CTA per frame. Outputs [frames][cap] (fv_start: [frames][cap + 1]); counts[frames] = (n_words, n_fv_nodes).
__global__ void __launch_bounds__(BW_THREADS) k_bow_finalize(const VocabDev V, const int32_t* __restrict__ n, const int cap, const int p2,
                                                             const int32_t* __restrict__ feat_word, const int32_t* __restrict__ feat_node,
                                                             const double* __restrict__ feat_w, int32_t* __restrict__ word_ids,
                                                             double* __restrict__ word_vals, uint32_t* __restrict__ fv_nodes,
                                                             int32_t* __restrict__ fv_start, uint32_t* __restrict__ fv_items, int2* __restrict__ counts)
{
	extern __shared__ __align__(16) uint64_t keys[];   // [p2]
	__shared__ int s_w[32];
	__shared__ double s_norm;
	const int frame = blockIdx.x;
	const int nf = min(n[frame], cap);
	const int64_t fo = (int64_t)frame * cap;
	const bool empty_voc = V.nwords == 0;              // :1137 empty(): both outputs stay empty

	// (a) BowVector: keys (word, feature) of the features whose word is not stopped (w > 0, :1159)
	for (int i = threadIdx.x; i < p2; i += BW_THREADS)
		keys[i] = (i < nf && !empty_voc && feat_w[fo + i] > 0.0) ? ((uint64_t)(uint32_t)feat_word[fo + i] << 32) | (uint32_t)i : ~0ull;
	__syncthreads();
	bitonic_sort(keys, p2);
	// kept = number of real keys = features that enter both maps
	int kept;
	{
		int c = 0;
		for (int i = threadIdx.x; i < p2; i += BW_THREADS) c += keys[i] != ~0ull;
		block_exscan(c, s_w, kept);
	}
	// run heads -> word slots. Each thread owns a contiguous chunk so that slots come out in ascending word order.
	const int chunk = (p2 + BW_THREADS - 1) / BW_THREADS, i0 = threadIdx.x * chunk, i1 = min(i0 + chunk, kept);
	int heads = 0;
	for (int i = i0; i < i1; i++) heads += (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32));
	int nwords;
	int slot = block_exscan(heads, s_w, nwords);
	const bool tf = V.weighting == 0 || V.weighting == 1;     // TF_IDF, TF: addWeight; IDF, BINARY: addIfNotExist (:1149, :1177)
	for (int i = i0; i < i1; i++)
		if (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32))
		{
			const uint32_t word = (uint32_t)(keys[i] >> 32);
			const double w = feat_w[fo + (uint32_t)keys[i]];     // the same for every feature of the word: Node::weight of the leaf
			double v = w;                                        // insert(id, w), then v += w per further feature, in feature order
			if (tf)
				for (int j = i + 1; j < kept && (uint32_t)(keys[j] >> 32) == word; j++) v = __dadd_rn(v, w);
			word_ids[fo + slot] = (int32_t)word;
			word_vals[fo + slot] = v;
			slot++;
		}
	__syncthreads();
	__threadfence_block();
	// normalisation (:1164-1170, :1196; BowVector::normalize): sequential in ascending word id, as std::map iterates
	const bool must = V.scoring != 5;                        // every scoring but DOT_PRODUCT normalises (ScoringObject.h:77-92)
	const bool l2 = V.scoring == 1;
	if (threadIdx.x == 0)
	{
		double norm = 0.0;
		if (must)
		{
			if (!l2) for (int i = 0; i < nwords; i++) norm = __dadd_rn(norm, fabs(word_vals[fo + i]));
			else
			{
				for (int i = 0; i < nwords; i++) { const double v = word_vals[fo + i]; norm = __dadd_rn(norm, __dmul_rn(v, v)); }
				norm = sqrt(norm);
			}
		}
		s_norm = norm;
	}
	__syncthreads();
	if (tf && !must && nwords > 0)
	{
		const double nd = (double)nwords;                    // :1166-1169
		for (int i = threadIdx.x; i < nwords; i += BW_THREADS) word_vals[fo + i] = __ddiv_rn(word_vals[fo + i], nd);
	}
	if (must && s_norm > 0.0)
		for (int i = threadIdx.x; i < nwords; i += BW_THREADS) word_vals[fo + i] = __ddiv_rn(word_vals[fo + i], s_norm);
	__syncthreads();

	// (b) FeatureVector: keys (node, feature) of the same features; ascending feature index inside a node = push_back order
	for (int i = threadIdx.x; i < p2; i += BW_THREADS)
		keys[i] = (i < nf && !empty_voc && feat_w[fo + i] > 0.0) ? ((uint64_t)(uint32_t)feat_node[fo + i] << 32) | (uint32_t)i : ~0ull;
	__syncthreads();
	bitonic_sort(keys, p2);
	heads = 0;
	for (int i = i0; i < i1; i++) heads += (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32));
	int nnodes;
	slot = block_exscan(heads, s_w, nnodes);
	int32_t* const st = fv_start + (int64_t)frame * (cap + 1);
	for (int i = i0; i < i1; i++)
	{
		if (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32))
		{
			fv_nodes[fo + slot] = (uint32_t)(keys[i] >> 32);
			st[slot] = i;
			slot++;
		}
		fv_items[fo + i] = (uint32_t)keys[i];
	}
	if (threadIdx.x == 0)
	{
		st[nnodes] = kept;
		counts[frame] = make_int2(nwords, nnodes);
	}
}

// ---- L1Scoring::score (ScoringObject.cpp:24-58) for pairs of sparse vectors: merge in ascending id, sequential FP64 sum ---------------
struct BowPair { const int32_t* ida; const double* va; int na; const int32_t* idb; const double* vb; int nb; };
__global__ void k_bow_score_l1(const BowPair* __restrict__ pairs, int npairs, double* __restrict__ out)
{
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= npairs) return;
	const BowPair P = pairs[p];
	int i = 0, j = 0;
	double score = 0.0;
	while (i < P.na && j < P.nb)
	{
		const int a = P.ida[i], b = P.idb[j];
		if (a == b)
		{
			const double vi = P.va[i], wi = P.vb[j];
			// score += fabs(vi - wi) - fabs(vi) - fabs(wi), left to right
			score = __dadd_rn(score, __dsub_rn(__dsub_rn(fabs(__dsub_rn(vi, wi)), fabs(vi)), fabs(wi)));
			++i; ++j;
		}
		else if (a < b)
		{
			// v1_it = v1.lower_bound(v2_it->first): first id >= b
			int lo = i, hi = P.na;
			while (lo < hi) { const int m = (lo + hi) >> 1; if (P.ida[m] < b) lo = m + 1; else hi = m; }
			i = lo;
		}
		else
		{
			int lo = j, hi = P.nb;
			while (lo < hi) { const int m = (lo + hi) >> 1; if (P.idb[m] < a) lo = m + 1; else hi = m; }
			j = lo;
		}
	}
	out[p] = __dmul_rn(-score, 0.5);     // score = -score/2.0
}

template <class T> struct Buf
{
	T* p = nullptr; size_t n = 0;
	cudaError_t ensure(size_t count)
	{
		if (count <= n) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr; n = 0;
		const cudaError_t e = cudaMalloc(&p, count * sizeof(T));
		if (e == cudaSuccess) n = count;
		return e;
	}
	~Buf() { if (p) cudaFree(p); }
};

}  // namespace

struct orbx_vocabulary_s
{
	int device = 0;
	cudaStream_t stream = nullptr;
	VocabDev V = {};
	Buf<int32_t> child_ids;
	Buf<int4> node_info;
	Buf<uint32_t> desc;
	Buf<double> weight;
	// scratch of the transform calls
	Buf<int32_t> feat_word, feat_node;
	Buf<double> feat_w;
	Buf<uint8_t> in_desc, out_block;
	uint8_t* h_in = nullptr; uint8_t* h_out = nullptr; size_t h_cap = 0;   // pinned staging of orbx_bow_transform
	Buf<BowPair> pairs; Buf<double> scores; Buf<int32_t> sc_ids; Buf<double> sc_vals;
};

#define BCU(x) do { const cudaError_t e_ = (x); if (e_ != cudaSuccess) return orbx_fail(ORBX_ERR_CUDA, cudaGetErrorString(e_)); } while (0)

namespace {

orbx_status build_vocabulary(int k, int L, int scoring, int weighting, int64_t nn, const int32_t* parent, const uint8_t* is_leaf,
                             const uint8_t* desc, const double* weights, int device, orbx_vocabulary* out)
{
	if (!out) return orbx_fail(ORBX_ERR_INVALID, "null output handle");
	*out = nullptr;
	// the checks of loadFromTextFile (TemplatedVocabulary.cpp:37)
	if (k < 0 || k > 20 || L < 1 || L > 10 || scoring < 0 || scoring > 5 || weighting < 0 || weighting > 3)
		return orbx_fail(ORBX_ERR_INVALID, "vocabulary header outside the ranges the reference accepts (k 0..20, L 1..10, scoring 0..5, weighting 0..3)");
	if (nn < 0 || nn > (int64_t)1 << 30 || (nn > 0 && (!parent || !is_leaf || !desc || !weights))) return orbx_fail(ORBX_ERR_INVALID, "bad node arrays");
	const char* why = nullptr;
	if (!orbx_device_usable(device, &why)) return orbx_fail(ORBX_ERR_CUDA, why);
	const int64_t total = nn + 1;
	// node i of the arrays is node id i + 1 (the file's line order); a parent must already exist (the reference indexes m_nodes[pid])
	std::vector<int32_t> cstart(total + 1, 0), cids(std::max<int64_t>(nn, 1)), wid(total, 0);
	std::vector<double> w(total, 0.0);
	std::vector<uint8_t> d(total * 32, 0);
	int64_t nwords = 0;
	for (int64_t i = 0; i < nn; i++)
	{
		if (parent[i] < 0 || parent[i] > i) return orbx_fail(ORBX_ERR_INVALID, "a node's parent id must refer to an earlier node");
		cstart[parent[i] + 1]++;
	}
	for (int64_t i = 0; i < total; i++) cstart[i + 1] += cstart[i];
	{
		std::vector<int32_t> cur(cstart.begin(), cstart.end() - 1);
		for (int64_t i = 0; i < nn; i++) cids[cur[parent[i]]++] = (int32_t)(i + 1);
	}
	for (int64_t i = 0; i < nn; i++)
	{
		w[i + 1] = weights[i];
		memcpy(&d[(i + 1) * 32], desc + i * 32, 32);
		if (is_leaf[i]) wid[i + 1] = (int32_t)nwords++;      // m_words.size() at the time the line is read (:77-83)
	}
	orbx_vocabulary_s* v = new orbx_vocabulary_s;
	v->device = device;
	BCU(cudaSetDevice(device));
	BCU(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
	BCU(v->child_ids.ensure(cids.size()));
	BCU(v->desc.ensure(total * 8)); BCU(v->weight.ensure(total));
	BCU(cudaMemcpy(v->child_ids.p, cids.data(), cids.size() * 4, cudaMemcpyHostToDevice));
	BCU(cudaMemcpy(v->desc.p, d.data(), total * 32, cudaMemcpyHostToDevice));
	BCU(cudaMemcpy(v->weight.p, w.data(), total * 8, cudaMemcpyHostToDevice));
	{
		std::vector<int4> info((size_t)total);
		for (int64_t i = 0; i < total; i++)
		{
			const int cs = cstart[i], nch = cstart[i + 1] - cstart[i];
			bool consecutive = nch > 0;
			for (int c = 1; c < nch; c++) consecutive = consecutive && cids[cs + c] == cids[cs] + c;
			info[i] = make_int4(cs, nch, consecutive ? cids[cs] : -1, wid[i]);
		}
		BCU(v->node_info.ensure(total));
		BCU(cudaMemcpy(v->node_info.p, info.data(), total * sizeof(int4), cudaMemcpyHostToDevice));
		v->V.node_info = v->node_info.p;
	}
	v->V.k = k; v->V.L = L; v->V.scoring = scoring; v->V.weighting = weighting; v->V.nnodes = total; v->V.nwords = nwords;
	v->V.child_ids = v->child_ids.p; v->V.desc = v->desc.p; v->V.weight = v->weight.p;
	*out = v;
	return ORBX_OK;
}

int next_pow2(int n) { int p = 32; while (p < n) p <<= 1; return p; }

orbx_status transform_device(orbx_vocabulary_s* v, const uint8_t* d_desc, const int32_t* d_n, int frames, int cap, int levelsup, int32_t* word_ids,
                             double* word_vals, uint32_t* fv_nodes, int32_t* fv_start, uint32_t* fv_items, int2* counts, int32_t* feat_word,
                             int32_t* feat_node, cudaStream_t st)
{
	if (cap > BW_MAX_FEATURES) return orbx_fail(ORBX_ERR_INVALID, "more than 16384 features per frame");
	const size_t tot = (size_t)frames * cap;
	BCU(v->feat_w.ensure(tot));
	if (!feat_word) { BCU(v->feat_word.ensure(tot)); feat_word = v->feat_word.p; }
	if (!feat_node) { BCU(v->feat_node.ensure(tot)); feat_node = v->feat_node.p; }
	dim3 g1((cap + 31) / 32, frames);
	k_bow_descend<<<g1, 256, 0, st>>>(v->V, d_desc, d_n, cap, levelsup, feat_word, feat_node, v->feat_w.p);
	const int p2 = next_pow2(cap);
	const size_t smem = (size_t)p2 * 8;
	static int attr[64] = {};
	if (v->device < 64 && attr[v->device] < (int)smem)
	{
		BCU(cudaFuncSetAttribute(k_bow_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		attr[v->device] = (int)smem;
	}
	k_bow_finalize<<<frames, BW_THREADS, smem, st>>>(v->V, d_n, cap, p2, feat_word, feat_node, v->feat_w.p, word_ids, word_vals, fv_nodes, fv_start, fv_items,
	                                                  counts);
	BCU(cudaGetLastError());
	return ORBX_OK;
}

}  // namespace

extern "C" {

orbx_status orbx_vocabulary_create(const orbx_vocabulary_desc* d, int device, orbx_vocabulary* out)
{
	if (!d) return orbx_fail(ORBX_ERR_INVALID, "null vocabulary description");
	return build_vocabulary(d->k, d->L, d->scoring, d->weighting, d->nnodes, d->parent, d->is_leaf, d->descriptors, d->weights, device, out);
}

// TemplatedVocabulary<FORB::TDescriptor, FORB>::loadFromTextFile — Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.cpp:21-90. Same tokenizer
// (strtok on ' ' + atoi over a 256-byte fgets buffer), including this fork's `int weight = getInt()`: a fractional weight is truncated.
orbx_status orbx_vocabulary_load_text(const char* path, int device, orbx_vocabulary* out)
{
	if (!path || !out) return orbx_fail(ORBX_ERR_INVALID, "null argument");
	*out = nullptr;
	FILE* fp = fopen(path, "r");
	if (!fp) return orbx_fail(ORBX_ERR_INVALID, "cannot open the vocabulary file");
	char buffer[256];
	if (!fgets(buffer, sizeof(buffer), fp)) { fclose(fp); return orbx_fail(ORBX_ERR_INVALID, "empty vocabulary file"); }
	int k = 0, L = 0, n1 = 0, n2 = 0;
	sscanf(buffer, "%d %d %d %d", &k, &L, &n1, &n2);
	std::vector<int32_t> parent; std::vector<uint8_t> leaf, desc; std::vector<double> weight;
	while (fgets(buffer, sizeof(buffer), fp))
	{
		char* tok = strtok(buffer, " ");
		auto get_int = [&]() { const int i = tok ? atoi(tok) : 0; tok = strtok(NULL, " "); return i; };
		parent.push_back(get_int());
		leaf.push_back(get_int() > 0);
		for (int i = 0; i < 32; i++) desc.push_back((uint8_t)get_int());
		weight.push_back((double)get_int());
	}
	fclose(fp);
	return build_vocabulary(k, L, n1, n2, (int64_t)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data(), device, out);
}

orbx_status orbx_vocabulary_info(orbx_vocabulary v, int* k, int* L, int* scoring, int* weighting, int64_t* nodes, int64_t* words)
{
	if (!v) return orbx_fail(ORBX_ERR_INVALID, "null vocabulary");
	if (k) *k = v->V.k; if (L) *L = v->V.L; if (scoring) *scoring = v->V.scoring; if (weighting) *weighting = v->V.weighting;
	if (nodes) *nodes = v->V.nnodes; if (words) *words = v->V.nwords;
	return ORBX_OK;
}

orbx_status orbx_vocabulary_destroy(orbx_vocabulary v)
{
	if (!v) return ORBX_OK;
	cudaSetDevice(v->device);
	if (v->stream) { cudaStreamSynchronize(v->stream); cudaStreamDestroy(v->stream); }
	if (v->h_in) cudaFreeHost(v->h_in);
	if (v->h_out) cudaFreeHost(v->h_out);
	delete v;
	return ORBX_OK;
}

orbx_status orbx_bow_transform_batch_device(orbx_vocabulary v, const uint8_t* d_desc, const int32_t* d_n, int frames, int cap, int levelsup,
                                            int32_t* d_word_ids, double* d_word_vals, uint32_t* d_fv_nodes, int32_t* d_fv_start, uint32_t* d_fv_items,
                                            int32_t* d_counts, int32_t* d_feat_word, int32_t* d_feat_node, void* stream)
{
	if (!v || !d_desc || !d_n || frames < 1 || cap < 1 || !d_word_ids || !d_word_vals || !d_fv_nodes || !d_fv_start || !d_fv_items || !d_counts)
		return orbx_fail(ORBX_ERR_INVALID, "null or empty argument");
	BCU(cudaSetDevice(v->device));
	return transform_device(v, d_desc, d_n, frames, cap, levelsup, d_word_ids, d_word_vals, d_fv_nodes, d_fv_start, d_fv_items,
	                        reinterpret_cast<int2*>(d_counts), d_feat_word, d_feat_node, stream ? (cudaStream_t)stream : v->stream);
}

orbx_status orbx_bow_transform(orbx_vocabulary v, const uint8_t* desc, int n, int levelsup, int32_t* word_ids, double* word_vals, int32_t* n_words,
                               uint32_t* fv_nodes, int32_t* fv_start, uint32_t* fv_items, int32_t* n_fv_nodes, int32_t* feat_word, int32_t* feat_node)
{
	if (!v || n < 0 || (n > 0 && !desc) || !word_ids || !word_vals || !n_words || !fv_nodes || !fv_start || !fv_items || !n_fv_nodes)
		return orbx_fail(ORBX_ERR_INVALID, "null argument");
	*n_words = 0; *n_fv_nodes = 0; fv_start[0] = 0;
	if (n == 0) return ORBX_OK;
	if (n > BW_MAX_FEATURES) return orbx_fail(ORBX_ERR_INVALID, "more than 16384 features per frame");
	BCU(cudaSetDevice(v->device));
	// One pinned staging buffer each way, one copy each way, one synchronisation: [n | descriptors] in,
	// [counts | word_vals | word_ids | fv_nodes | fv_start | fv_items | feat_word | feat_node] out.
	const size_t cap = (size_t)n, c16 = (cap * 4 + 15) & ~(size_t)15;
	const size_t in_bytes = 16 + cap * 32;
	const size_t o_vals = 16, o_ids = o_vals + ((cap * 8 + 15) & ~(size_t)15), o_nodes = o_ids + c16, o_start = o_nodes + c16, o_items = o_start + c16 + 16,
	             o_fw = o_items + c16, o_fn = o_fw + c16, out_bytes = o_fn + c16;
	if (v->h_cap < std::max(in_bytes, out_bytes))
	{
		if (v->h_in) cudaFreeHost(v->h_in);
		if (v->h_out) cudaFreeHost(v->h_out);
		v->h_in = v->h_out = nullptr; v->h_cap = 0;
		const size_t want = 2 * std::max(in_bytes, out_bytes);
		BCU(cudaMallocHost(&v->h_in, want)); BCU(cudaMallocHost(&v->h_out, want));
		v->h_cap = want;
	}
	BCU(v->in_desc.ensure(in_bytes)); BCU(v->out_block.ensure(out_bytes));
	cudaStream_t st = v->stream;
	memcpy(v->h_in, &n, 4);
	memcpy(v->h_in + 16, desc, cap * 32);
	BCU(cudaMemcpyAsync(v->in_desc.p, v->h_in, in_bytes, cudaMemcpyHostToDevice, st));
	uint8_t* o = v->out_block.p;
	const orbx_status s = transform_device(v, v->in_desc.p + 16, reinterpret_cast<const int32_t*>(v->in_desc.p), 1, (int)cap, levelsup,
	                                       reinterpret_cast<int32_t*>(o + o_ids), reinterpret_cast<double*>(o + o_vals), reinterpret_cast<uint32_t*>(o + o_nodes),
	                                       reinterpret_cast<int32_t*>(o + o_start), reinterpret_cast<uint32_t*>(o + o_items), reinterpret_cast<int2*>(o),
	                                       reinterpret_cast<int32_t*>(o + o_fw), reinterpret_cast<int32_t*>(o + o_fn), st);
	if (s != ORBX_OK) return s;
	BCU(cudaMemcpyAsync(v->h_out, o, out_bytes, cudaMemcpyDeviceToHost, st));
	BCU(cudaStreamSynchronize(st));
	const int2 c = *reinterpret_cast<const int2*>(v->h_out);
	*n_words = c.x; *n_fv_nodes = c.y;
	memcpy(word_ids, v->h_out + o_ids, (size_t)c.x * 4);
	memcpy(word_vals, v->h_out + o_vals, (size_t)c.x * 8);
	memcpy(fv_nodes, v->h_out + o_nodes, (size_t)c.y * 4);
	memcpy(fv_start, v->h_out + o_start, (size_t)(c.y + 1) * 4);
	memcpy(fv_items, v->h_out + o_items, (size_t)reinterpret_cast<const int32_t*>(v->h_out + o_start)[c.y] * 4);
	if (feat_word) memcpy(feat_word, v->h_out + o_fw, cap * 4);
	if (feat_node) memcpy(feat_node, v->h_out + o_fn, cap * 4);
	return ORBX_OK;
}

orbx_status orbx_bow_score_l1(orbx_vocabulary v, const int32_t* ids, const double* vals, const int32_t* offsets, const int32_t* pair_a,
                              const int32_t* pair_b, int npairs, double* scores)
{
	if (!v || npairs < 0 || (npairs > 0 && (!ids || !vals || !offsets || !pair_a || !pair_b || !scores))) return orbx_fail(ORBX_ERR_INVALID, "null argument");
	if (npairs == 0) return ORBX_OK;
	BCU(cudaSetDevice(v->device));
	int nvec = 0;
	for (int p = 0; p < npairs; p++)
	{
		if (pair_a[p] < 0 || pair_b[p] < 0) return orbx_fail(ORBX_ERR_INVALID, "negative vector index");
		nvec = std::max(nvec, std::max(pair_a[p], pair_b[p]) + 1);
	}
	const size_t total = (size_t)offsets[nvec];
	BCU(v->sc_ids.ensure(std::max<size_t>(total, 1))); BCU(v->sc_vals.ensure(std::max<size_t>(total, 1)));
	BCU(v->pairs.ensure(npairs)); BCU(v->scores.ensure(npairs));
	std::vector<BowPair> hp(npairs);
	for (int p = 0; p < npairs; p++)
	{
		const int a = pair_a[p], b = pair_b[p];
		hp[p] = BowPair{ v->sc_ids.p + offsets[a], v->sc_vals.p + offsets[a], offsets[a + 1] - offsets[a],
		                 v->sc_ids.p + offsets[b], v->sc_vals.p + offsets[b], offsets[b + 1] - offsets[b] };
	}
	cudaStream_t st = v->stream;
	if (total) { BCU(cudaMemcpyAsync(v->sc_ids.p, ids, total * 4, cudaMemcpyHostToDevice, st)); BCU(cudaMemcpyAsync(v->sc_vals.p, vals, total * 8, cudaMemcpyHostToDevice, st)); }
	BCU(cudaMemcpyAsync(v->pairs.p, hp.data(), (size_t)npairs * sizeof(BowPair), cudaMemcpyHostToDevice, st));
	k_bow_score_l1<<<(npairs + 127) / 128, 128, 0, st>>>(v->pairs.p, npairs, v->scores.p);
	BCU(cudaGetLastError());
	BCU(cudaMemcpyAsync(scores, v->scores.p, (size_t)npairs * 8, cudaMemcpyDeviceToHost, st));
	BCU(cudaStreamSynchronize(st));
	return ORBX_OK;
}

}  // extern "C"
