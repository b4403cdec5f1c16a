// TEST INFRASTRUCTURE — driver for the reference's own bag-of-words transform: Thirdparty/DBoW2/DBoW2/{TemplatedVocabulary.h, FORB.cpp,
// BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp} and Thirdparty/DBoW2/DUtils/{Random.cpp, Timestamp.cpp} compiled UNMODIFIED where
// they lie against the OpenCV stand-in (oracle/cvshim).  ORBvoc.bin is absent from the reference tree, so the vocabulary is built by
// the reference's own create() (hierarchical k-means++, srand seeded through DUtils::Random::SeedRandOnce(seed)) from the training
// descriptors of in.bin, dumped node by node, and then used for TemplatedVocabulary::transform(features, BowVector, FeatureVector,
// levelsup) — the call of Frame::ComputeBoW (src/Frame.cc:1692-1699, levelsup = 4).
//
//   ref_bow <in.bin> <out.bin>
//   in.bin : int32 magic 'BOWV', k, L, seed, levelsup, ntrain_images ; per image: int32 n, n x 32 B descriptors ;
//            int32 nframes ; per frame: int32 n, n x 32 B descriptors
//   out.bin: int32 n_nodes ; per node: int32 parent, word_id (-1 = inner node), n_children, children ids ; 32 B descriptor ; double weight ;
//            per frame: int32 nw ; nw x (int32 word id, double value) ; int32 nn ; per node: int32 node id, count, count x int32 features
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <opencv2/core/core.hpp>
#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"
#include "DUtils/Random.h"
#ifdef REF_BOW_USE_SHIM   // oracle/_ref/shim_bow: the same driver, the transform through the drop-in (shim/ORBVocabularyGPU.h -> C ABI -> CUDA)
#include "ORBVocabularyGPU.h"
#endif

using namespace DBoW2;

class Voc : public TemplatedVocabulary<FORB::TDescriptor, FORB> {   // exposes the protected tree for the dump
public:
    Voc(int k, int L) : TemplatedVocabulary<FORB::TDescriptor, FORB>(k, L, TF_IDF, L1_NORM) {}   // ORBvoc: TF-IDF weights, L1 scoring
    const std::vector<Node>& nodes() const { return m_nodes; }
};

static FILE *g_in, *g_out;
template <class T> static T get() { T v; if (std::fread(&v, sizeof(T), 1, g_in) != 1) { std::fprintf(stderr, "ref_bow: short input\n"); std::exit(4); } return v; }
template <class T> static void put(const T& v) { std::fwrite(&v, sizeof(T), 1, g_out); }
static std::vector<cv::Mat> get_descriptors() {
    const int n = get<int32_t>();
    std::vector<cv::Mat> d(n);
    for (int i = 0; i < n; ++i) {
        d[i] = cv::Mat(1, 32, CV_8UC1);
        if (std::fread(d[i].data, 1, 32, g_in) != 32) std::exit(4);
    }
    return d;
}

int main(int argc, char** argv) {
    if (argc != 3) { std::fprintf(stderr, "usage: ref_bow in.bin out.bin\n"); return 2; }
    g_in = std::fopen(argv[1], "rb");
    g_out = std::fopen(argv[2], "wb");
    if (!g_in || !g_out) return 2;
    if (get<int32_t>() != 0x424f5756) return 3;
    const int k = get<int32_t>(), L = get<int32_t>(), seed = get<int32_t>(), levelsup = get<int32_t>(), nimg = get<int32_t>();
    std::vector<std::vector<cv::Mat>> training(nimg);
    for (int i = 0; i < nimg; ++i) training[i] = get_descriptors();
    DUtils::Random::SeedRandOnce(seed);
    Voc voc(k, L);
    voc.create(training);
    const auto& nodes = voc.nodes();
    put<int32_t>((int32_t)nodes.size());
    for (size_t i = 0; i < nodes.size(); ++i) {
        const auto& nd = nodes[i];
        put<int32_t>((int32_t)nd.parent);
        put<int32_t>(nd.isLeaf() && i != 0 ? (int32_t)nd.word_id : -1);
        put<int32_t>((int32_t)nd.children.size());
        for (NodeId c : nd.children) put<int32_t>((int32_t)c);
        uint8_t zero[32] = {0};
        std::fwrite(nd.descriptor.empty() ? zero : nd.descriptor.data, 1, 32, g_out);
        put<double>(nd.weight);
    }
    const int nframes = get<int32_t>();
#ifdef REF_BOW_USE_SHIM
    hvo_shim::BowTransformerT<TemplatedVocabulary<FORB::TDescriptor, FORB>> gpu(voc);   // handed the reference's own vocabulary object
    if (!gpu.valid()) return 7;
#endif
    for (int f = 0; f < nframes; ++f) {
        std::vector<cv::Mat> feats = get_descriptors();
        BowVector bv;
        FeatureVector fv;
#ifdef REF_BOW_USE_SHIM
        gpu.transform(feats, bv, fv, levelsup);
#else
        voc.transform(feats, bv, fv, levelsup);
#endif
        put<int32_t>((int32_t)bv.size());
        for (const auto& e : bv) { put<int32_t>((int32_t)e.first); put<double>(e.second); }
        put<int32_t>((int32_t)fv.size());
        for (const auto& e : fv) {
            put<int32_t>((int32_t)e.first); put<int32_t>((int32_t)e.second.size());
            for (unsigned int x : e.second) put<int32_t>((int32_t)x);
        }
    }
    std::fclose(g_in);
    std::fclose(g_out);
    return 0;
}
