// CPU check of csrc/std_sort.cuh (the libstdc++ std::sort restatement the CUDA line kernels use to reproduce the reference's
// unstable response sort, include/auxiliar.h:47-52) against the real std::sort of this toolchain, on structs like the
// reference sorts (68-byte KeyLines compared by response only), with many ties.  Exit code 0 = identical permutations.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../a-low-texture-robust-hybrid-feature-based-visual-odometry_b200/csrc/std_sort.cuh"

struct KeyLineLike { float angle; int class_id; float pad[14]; float response; };  // compared by response only
struct ByResponse { bool operator()(const KeyLineLike& a, const KeyLineLike& b) const { return a.response > b.response; } };

int main() {
    std::mt19937 rng(7);
    long checked = 0;
    for (int trial = 0; trial < 3000; ++trial) {
        const int n = trial < 40 ? trial : 1 + (int)(rng() % (trial % 7 == 0 ? 5000 : 700));
        const int levels = 1 + (int)(rng() % (trial % 3 == 0 ? 4 : (trial % 3 == 1 ? 60 : 100000)));   // few levels = many ties
        std::vector<KeyLineLike> v(n);
        std::vector<float> key(n);
        for (int i = 0; i < n; ++i) {
            v[i].class_id = i;
            v[i].response = key[i] = (float)(rng() % levels) / 64.f;
        }
        if (trial % 11 == 0) std::sort(key.begin(), key.end());                    // presorted ascending: adversarial for median-of-3
        if (trial % 13 == 0) std::sort(key.begin(), key.end(), std::greater<float>());
        for (int i = 0; i < n; ++i) v[i].response = key[i];
        std::sort(v.begin(), v.end(), ByResponse());
        std::vector<uint16_t> idx(n);
        for (int i = 0; i < n; ++i) idx[i] = (uint16_t)i;
        const float* K = key.data();
        hvo::stdsort::sort(idx.data(), n, [K](uint16_t a, uint16_t b) { return K[a] > K[b]; });
        for (int i = 0; i < n; ++i)
            if ((int)idx[i] != v[i].class_id) { std::printf("mismatch: trial %d n %d at %d\n", trial, n, i); return 1; }
        checked += n;
        // the depth-limit fallback on its own: std::__partial_sort(first, last, last) = make_heap + sort_heap
        for (int i = 0; i < n; ++i) { v[i].class_id = i; v[i].response = key[i]; idx[i] = (uint16_t)i; }
        std::partial_sort(v.begin(), v.end(), v.end(), ByResponse());
        hvo::stdsort::heap_sort(idx.data(), idx.data() + n, [K](uint16_t a, uint16_t b) { return K[a] > K[b]; });
        for (int i = 0; i < n; ++i)
            if ((int)idx[i] != v[i].class_id) { std::printf("heap mismatch: trial %d n %d at %d\n", trial, n, i); return 1; }
    }
    std::printf("ok %ld elements\n", checked);
    return 0;
}
