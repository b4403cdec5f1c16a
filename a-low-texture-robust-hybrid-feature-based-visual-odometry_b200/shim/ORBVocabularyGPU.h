// Device side of Frame::ComputeBoW (reference src/Frame.cc:1692-1699): the one call it makes,
//
//     mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4);
//
// i.e. DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>::transform(features, BowVector&, FeatureVector&, levelsup)
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1137-1206, 1228-1270), with the same signature and results:
//
//     #include "shim/ORBVocabularyGPU.h"
//     hvo_shim::BowTransformerT<ORBVocabulary> bow(*mpORBvocabulary);         // once, after the vocabulary is loaded
//     bow.transform(vCurrentDesc, mBowVec, mFeatVec, 4);                       // instead of mpORBvocabulary->transform(...)
//
// The vocabulary object stays the reference's own (loading, scoring, the key-frame database keep using it); the constructor reads its tree
// (m_nodes, m_L: protected members, reached through a derived accessor, no change to DBoW2) once and hands it to the device as arrays
// (hvo_bow_set_vocabulary).  transform() descends the tree on the GPU (hvo_bow_transform: one warp per descriptor), and fills the two
// std::map-derived outputs in ascending key order, which is the order the reference's own insertions produce; BowVector values are the
// reference's doubles bit for bit (accumulated and L1-normalised in its order).  TF-IDF / L1 is what ORB-SLAM2's vocabulary uses
// (ORBvoc.txt header); other weighting / scoring types are refused at construction.
#ifndef HVO_SHIM_ORBVOCABULARYGPU_H
#define HVO_SHIM_ORBVOCABULARYGPU_H

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <utility>
#include <vector>

#include "hvo_capi.h"

namespace hvo_shim {

template <class TVoc>
class BowTransformerT {
    // m_nodes / m_L are protected in TemplatedVocabulary (TemplatedVocabulary.h:417-439); a derived class may form pointers to them
    struct Access : public TVoc {
        static auto nodes() -> decltype(&Access::m_nodes) { return &Access::m_nodes; }
        static auto levels() -> decltype(&Access::m_L) { return &Access::m_L; }
    };

public:
    explicit BowTransformerT(const TVoc& voc, int device = 0) : h_(nullptr), ok_(false) {
        if (hvo_bow_create(device, &h_) != HVO_OK) { std::fprintf(stderr, "BowTransformer: %s\n", hvo_last_error()); return; }
        const auto& nodes = voc.*(Access::nodes());
        const int L = voc.*(Access::levels());
        const int n = (int)nodes.size();
        std::vector<int32_t> child_start(n + 1, 0), child_ids, word(n, -1);
        std::vector<uint8_t> desc((size_t)n * 32, 0);
        std::vector<double> weight(n, 0.0);
        for (int i = 0; i < n; ++i) {
            for (size_t c = 0; c < nodes[i].children.size(); ++c) child_ids.push_back((int32_t)nodes[i].children[c]);
            child_start[i + 1] = (int32_t)child_ids.size();
            if (!nodes[i].descriptor.empty()) std::memcpy(&desc[(size_t)i * 32], nodes[i].descriptor.template ptr<uint8_t>(), 32);
            weight[i] = nodes[i].weight;
            if (i != 0 && nodes[i].isLeaf()) word[i] = (int32_t)nodes[i].word_id;
        }
        ok_ = n > 0 && hvo_bow_set_vocabulary(h_, n, child_start.data(), child_ids.empty() ? nullptr : child_ids.data(), desc.data(), weight.data(),
                                              word.data(), L) == HVO_OK;
        if (!ok_) std::fprintf(stderr, "BowTransformer: %s\n", n > 0 ? hvo_last_error() : "empty vocabulary");
    }
    ~BowTransformerT() { hvo_bow_destroy(h_); }
    BowTransformerT(const BowTransformerT&) = delete;
    BowTransformerT& operator=(const BowTransformerT&) = delete;
    bool valid() const { return ok_; }

    // TemplatedVocabulary::transform(features, v, fv, levelsup); BowVector / FeatureVector are the reference's own types (deduced)
    template <class BowVector, class FeatureVector>
    void transform(const std::vector<cv::Mat>& features, BowVector& v, FeatureVector& fv, int levelsup) const {
        v.clear();
        fv.clear();
        const int n = (int)features.size();
        if (!ok_ || n == 0) return;
        std::vector<uint8_t> desc((size_t)n * 32);
        for (int i = 0; i < n; ++i) std::memcpy(&desc[(size_t)i * 32], features[i].template ptr<uint8_t>(), 32);
        const int32_t offsets[2] = {0, n};
        std::vector<int32_t> node_of(n), words(n), order(n);
        std::vector<double> values(n);
        int32_t nwords = 0, nfeat = 0;
        if (hvo_bow_transform(h_, desc.data(), offsets, 1, levelsup, nullptr, node_of.data(), &nwords, words.data(), values.data(), order.data(), &nfeat) !=
            HVO_OK) {
            std::fprintf(stderr, "BowTransformer: %s\n", hvo_last_error());
            return;
        }
        typedef typename BowVector::key_type WordId;
        typedef typename BowVector::mapped_type WordValue;
        typedef typename FeatureVector::key_type NodeId;
        for (int k = 0; k < nwords; ++k) v.insert(v.end(), std::make_pair((WordId)words[k], (WordValue)values[k]));   // ascending word id
        for (int k = 0; k < nfeat; ++k) {                                                                               // ascending (node, feature)
            const NodeId nid = (NodeId)node_of[order[k]];
            typename FeatureVector::iterator it = fv.end();
            if (fv.empty() || (--it)->first != nid) it = fv.insert(fv.end(), std::make_pair(nid, typename FeatureVector::mapped_type()));
            it->second.push_back((unsigned int)order[k]);
        }
    }

private:
    hvo_bow* h_;
    bool ok_;
};

}  // namespace hvo_shim

#endif
