// TEST INFRASTRUCTURE — CPU oracle, not the product.
//
// Stand-in declarations so that the reference's own text of
//   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:300-332    struct Node
//   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:465-500    createScoringObject
//   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1011-1015  empty
//   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1129-1197  transform(features, BowVector&, FeatureVector&, levelsup)
//   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1220-1262  transform(feature, word_id, weight, nid, levelsup)
//   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.cpp:7-90     GetInt, loadFromTextFile
//   Thirdparty/DBoW2/DBoW2/FORB.cpp:79-100                  FORB::distance
// compiles by line range (oracle/Makefile, rule bow_gen.cc). The real class template cannot be instantiated against the OpenCV
// stand-in: its virtual save/load members use cv::FileStorage. Only the members the ranges above touch are declared, under the same
// names (TemplatedVocabulary.h:47-433, FORB.h:22-75). BowVector.cpp, FeatureVector.cpp and ScoringObject.cpp (from line 14: the part
// after its #include of TemplatedVocabulary.h) are compiled as they are.
//
// Included INSIDE namespace DBoW2 of the generated TU.
#pragma once

using std::vector;
using std::endl;

class FORB
{
public:
	typedef cv::Mat TDescriptor;
	typedef const TDescriptor* pDescriptor;
	static const int L;
	static int distance(const TDescriptor& a, const TDescriptor& b);
};
const int FORB::L = 32;       // FORB.cpp:24

template <class TDescriptor, class F>
class TemplatedVocabulary
{
public:
	TemplatedVocabulary() : m_k(10), m_L(5), m_weighting(TF_IDF), m_scoring(L1_NORM), m_scoring_object(NULL) {}
	~TemplatedVocabulary() { delete m_scoring_object; }
	bool empty() const;
	void transform(const std::vector<TDescriptor>& features, BowVector& v, FeatureVector& fv, int levelsup) const;
	double score(const BowVector& a, const BowVector& b) const { return m_scoring_object->score(a, b); }   // :1201-1206
	bool loadFromTextFile(const std::string& filename);

	struct Node;
	void transform(const TDescriptor& feature, WordId& id, WordValue& weight, NodeId* nid = NULL, int levelsup = 0) const;
	void createScoringObject();

// (struct Node follows from the reference text, then the data members)
