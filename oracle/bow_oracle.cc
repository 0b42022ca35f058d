// TEST INFRASTRUCTURE — CPU oracle, not the product. Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use it.
//
// Stand-alone restatement of the bag-of-words transform — what travels to the GPU box, where /root/reference does not
// exist. Checked against the reference's own DBoW2 text (oracle/_ref, rule bow_gen.cc) by tests/test_oracle_vs_ref.py and pinned by
// tests/golden/bow.npz (outputs of that reference text).
//
//   load_text      Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.cpp:21-90     (strtok/atoi tokenizer; the weight is read as an int, :67)
//   descend        Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1220-1262   (strict d < best_d; nid at level L - levelsup)
//   distance       Thirdparty/DBoW2/DBoW2/FORB.cpp:79-100
//   transform      Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1129-1197   + BowVector.cpp:32-87, FeatureVector.cpp:30-44
//   score (L1)     Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:24-58
//
// The two std::maps are restated in pass form (the form the CUDA kernels use): sort (word, feature) and (node, feature) keys, a word's
// value is w added once per feature in feature order, the norm is one sequential pass in ascending word id.
#define ORACLE_PREFIX orc_
#include "oracle_api.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct Voc
{
	int k = 0, L = 0, scoring = 0, weighting = 0;
	std::vector<std::vector<int>> children;   // per node, push_back order
	std::vector<uint8_t> desc;                 // 32 B per node
	std::vector<double> weight;
	std::vector<int> word_id;
	int nwords = 0;
};

int distance(const uint8_t* a, const uint8_t* b)
{
	int d = 0;
	for (int i = 0; i < 8; i++)
	{
		uint32_t x, y;
		memcpy(&x, a + 4 * i, 4); memcpy(&y, b + 4 * i, 4);
		d += __builtin_popcount(x ^ y);
	}
	return d;
}

void descend(const Voc& V, const uint8_t* f, int levelsup, int& word, double& w, int& nid)
{
	const int nid_level = V.L - levelsup;
	nid = 0;                        // root when nid_level <= 0; a leaf above nid_level leaves the caller's variable unset in the reference: 0 here
	int node = 0, level = 0;
	do
	{
		++level;
		const std::vector<int>& ch = V.children[node];
		node = ch[0];
		int best = distance(f, &V.desc[(size_t)node * 32]);
		for (size_t c = 1; c < ch.size(); c++)
		{
			const int d = distance(f, &V.desc[(size_t)ch[c] * 32]);
			if (d < best) { best = d; node = ch[c]; }
		}
		if (level == nid_level) nid = node;
	} while (!V.children[node].empty());
	word = V.word_id[node];
	w = V.weight[node];
}

int transform(const Voc& V, const uint8_t* desc, int n, int levelsup, int32_t* word_ids, double* word_vals, uint32_t* fv_nodes, int32_t* fv_start,
              uint32_t* fv_items, int32_t* n_fv_nodes)
{
	*n_fv_nodes = 0; fv_start[0] = 0;
	if (V.nwords == 0) return 0;
	std::vector<uint64_t> kw, kn;
	std::vector<double> wf((size_t)n);
	for (int i = 0; i < n; i++)
	{
		int word, nid; double w;
		descend(V, desc + (size_t)i * 32, levelsup, word, w, nid);
		wf[i] = w;
		if (w > 0)
		{
			kw.push_back(((uint64_t)(uint32_t)word << 32) | (uint32_t)i);
			kn.push_back(((uint64_t)(uint32_t)nid << 32) | (uint32_t)i);
		}
	}
	std::sort(kw.begin(), kw.end());
	std::sort(kn.begin(), kn.end());
	const bool tf = V.weighting == 0 || V.weighting == 1;
	int nw = 0;
	for (size_t i = 0; i < kw.size();)
	{
		const uint32_t word = (uint32_t)(kw[i] >> 32);
		const double w = wf[(uint32_t)kw[i]];
		double v = w;
		size_t j = i + 1;
		for (; j < kw.size() && (uint32_t)(kw[j] >> 32) == word; j++)
			if (tf) v += w;
		word_ids[nw] = (int32_t)word; word_vals[nw] = v; nw++;
		i = j;
	}
	const bool must = V.scoring != 5, l2 = V.scoring == 1;
	if (tf && !must && nw > 0)
		for (int i = 0; i < nw; i++) word_vals[i] /= (double)nw;
	if (must)
	{
		double norm = 0.0;
		if (!l2) for (int i = 0; i < nw; i++) norm += std::fabs(word_vals[i]);
		else
		{
			for (int i = 0; i < nw; i++) norm += word_vals[i] * word_vals[i];
			norm = std::sqrt(norm);
		}
		if (norm > 0.0)
			for (int i = 0; i < nw; i++) word_vals[i] /= norm;
	}
	int nn = 0;
	for (size_t i = 0; i < kn.size(); i++)
	{
		if (i == 0 || (kn[i] >> 32) != (kn[i - 1] >> 32)) { fv_nodes[nn] = (uint32_t)(kn[i] >> 32); fv_start[nn] = (int32_t)i; nn++; }
		fv_items[i] = (uint32_t)kn[i];
	}
	fv_start[nn] = (int32_t)kn.size();
	*n_fv_nodes = nn;
	return nw;
}

Voc* finish(Voc* v, const std::vector<int>& parent, const std::vector<uint8_t>& leaf)
{
	const size_t nn = parent.size();
	v->children.assign(nn + 1, std::vector<int>());
	v->word_id.assign(nn + 1, 0);
	for (size_t i = 0; i < nn; i++)
	{
		if (parent[i] < 0 || (size_t)parent[i] > i) { delete v; return nullptr; }
		v->children[(size_t)parent[i]].push_back((int)i + 1);
		if (leaf[i]) v->word_id[i + 1] = v->nwords++;
	}
	return v;
}

}  // namespace

extern "C" {

void* orc_voc_load_text(const char* path)
{
	FILE* fp = fopen(path, "r");
	if (!fp) return nullptr;
	char buffer[256];
	if (!fgets(buffer, sizeof(buffer), fp)) { fclose(fp); return nullptr; }
	Voc* v = new Voc;
	int n1 = 0, n2 = 0;
	sscanf(buffer, "%d %d %d %d", &v->k, &v->L, &n1, &n2);
	if (v->k < 0 || v->k > 20 || v->L < 1 || v->L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) { fclose(fp); delete v; return nullptr; }
	v->scoring = n1; v->weighting = n2;
	std::vector<int> parent; std::vector<uint8_t> leaf;
	v->desc.assign(32, 0); v->weight.assign(1, 0.0);
	while (fgets(buffer, sizeof(buffer), fp))
	{
		char* tok = strtok(buffer, " ");
		auto get_int = [&]() { const int i = atoi(tok); tok = strtok(NULL, " "); return i; };
		parent.push_back(get_int());
		leaf.push_back(get_int() > 0);
		for (int i = 0; i < 32; i++) v->desc.push_back((uint8_t)get_int());
		v->weight.push_back((double)get_int());
	}
	fclose(fp);
	return finish(v, parent, leaf);
}

// the same tree from arrays (entry i = node id i + 1, file order), for vocabularies too large to go through a text file in a test
void* orc_voc_create(int k, int L, int scoring, int weighting, int64_t nnodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc,
                     const double* weights)
{
	Voc* v = new Voc;
	v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting;
	v->desc.assign(32, 0); v->desc.insert(v->desc.end(), desc, desc + nnodes * 32);
	v->weight.assign(1, 0.0); v->weight.insert(v->weight.end(), weights, weights + nnodes);
	return finish(v, std::vector<int>(parent, parent + nnodes), std::vector<uint8_t>(is_leaf, is_leaf + nnodes));
}

void orc_voc_destroy(void* voc) { delete static_cast<Voc*>(voc); }

int orc_bow_transform(void* voc, const uint8_t* desc, int n, int levelsup, int32_t* word_ids, double* word_vals, uint32_t* fv_nodes, int32_t* fv_start,
                      uint32_t* fv_items, int32_t* n_fv_nodes)
{
	return transform(*static_cast<const Voc*>(voc), desc, n, levelsup, word_ids, word_vals, fv_nodes, fv_start, fv_items, n_fv_nodes);
}

// per-feature results of the tree walk (word, levelsup-ancestor), stopped or not
void orc_bow_descend(void* voc, const uint8_t* desc, int n, int levelsup, int32_t* feat_word, int32_t* feat_node)
{
	for (int i = 0; i < n; i++)
	{
		int word, nid; double w;
		descend(*static_cast<const Voc*>(voc), desc + (size_t)i * 32, levelsup, word, w, nid);
		feat_word[i] = word; feat_node[i] = nid;
	}
}

double orc_bow_score(void* voc, const int32_t* ida, const double* va, int na, const int32_t* idb, const double* vb, int nb)
{
	// L1Scoring::score: both vectors ascending; lower_bound on the other vector = skip to the first id >= the current one
	int i = 0, j = 0;
	double score = 0;
	while (i < na && j < nb)
	{
		if (ida[i] == idb[j]) { score += std::fabs(va[i] - vb[j]) - std::fabs(va[i]) - std::fabs(vb[j]); ++i; ++j; }
		else if ((uint32_t)ida[i] < (uint32_t)idb[j]) { while (i < na && (uint32_t)ida[i] < (uint32_t)idb[j]) ++i; }
		else { while (j < nb && (uint32_t)idb[j] < (uint32_t)ida[i]) ++j; }
	}
	return -score / 2.0;
}

double orc_time_bow_transform(void* voc, const uint8_t* desc, int n, int levelsup, int reps)
{
	std::vector<int32_t> wi((size_t)n + 1), st((size_t)n + 2); std::vector<double> wv((size_t)n + 1); std::vector<uint32_t> fn((size_t)n + 1), fi((size_t)n + 1);
	int32_t nfv;
	const auto t0 = std::chrono::steady_clock::now();
	for (int r = 0; r < reps; r++) transform(*static_cast<const Voc*>(voc), desc, n, levelsup, wi.data(), wv.data(), fn.data(), st.data(), fi.data(), &nfv);
	return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
}

}  // extern "C"
