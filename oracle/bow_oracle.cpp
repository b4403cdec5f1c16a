// TEST INFRASTRUCTURE — CPU restatement of DBoW2::TemplatedVocabulary::transform(features, BowVector, FeatureVector, levelsup)
// (reference Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1137-1206 and :1228-1270, FORB::distance FORB.cpp:81-101, BowVector::addWeight /
// normalize BowVector.cpp:33-83) for a TF-IDF / L1 vocabulary given as arrays.  Not on the product path.  Pinned by executing the
// reference's own DBoW2 sources: oracle/_ref/ref_bow (oracle/ref_bow_main.cpp), golden tests/golden/bow_ref.npz.
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>

static int bow_distance(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

extern "C" {

// one frame: n descriptors.  word_of / node_of [n]; bow_words / bow_values [n] (first *nw filled); fv_order [n] (first *nfv filled)
void orc_bow_transform(const int32_t* child_start, const int32_t* child_ids, const uint8_t* node_desc, const double* node_weight,
                       const int32_t* node_word, int L, const uint8_t* desc, int n, int levelsup, int32_t* word_of, int32_t* node_of,
                       int32_t* bow_words, double* bow_values, int32_t* nw, int32_t* fv_order, int32_t* nfv) {
    std::map<unsigned, double> v;
    std::map<unsigned, std::vector<unsigned>> fv;
    const int nid_level = L - levelsup;
    for (int f = 0; f < n; ++f) {
        const uint8_t* q = desc + 32 * (size_t)f;
        int final_id = 0, current_level = 0, nid = 0;
        while (child_start[final_id] != child_start[final_id + 1]) {
            ++current_level;
            const int b = child_start[final_id], e = child_start[final_id + 1];
            int best = child_ids[b];
            double best_d = bow_distance(q, node_desc + 32 * (size_t)best);
            for (int c = b + 1; c < e; ++c) {
                const double d = bow_distance(q, node_desc + 32 * (size_t)child_ids[c]);
                if (d < best_d) { best_d = d; best = child_ids[c]; }
            }
            final_id = best;
            if (current_level == nid_level) nid = final_id;
        }
        const double w = node_weight[final_id];
        node_of[f] = nid;
        word_of[f] = w > 0 ? node_word[final_id] : -1;
        if (w > 0) {
            auto it = v.find((unsigned)node_word[final_id]);
            if (it != v.end()) it->second += w; else v[(unsigned)node_word[final_id]] = w;
            fv[(unsigned)nid].push_back((unsigned)f);
        }
    }
    double norm = 0.0;
    for (auto& e : v) norm += std::fabs(e.second);
    if (norm > 0.0) for (auto& e : v) e.second /= norm;
    int k = 0;
    for (auto& e : v) { bow_words[k] = (int32_t)e.first; bow_values[k] = e.second; ++k; }
    *nw = k;
    k = 0;
    for (auto& e : fv) for (unsigned x : e.second) fv_order[k++] = (int32_t)x;
    *nfv = k;
}

}  // extern "C"
