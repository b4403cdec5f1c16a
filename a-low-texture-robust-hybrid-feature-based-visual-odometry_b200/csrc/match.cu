// Brute-force Hamming matching for 256-bit binary descriptors (ORB rBRIEF and LBD), sm_100a.
//
// Replaces the distance loops of ORBmatcher::DescriptorDistance (reference src/ORBmatcher.cc:1676-1692) /
// LSDmatcher::DescriptorDistance (src/LSDmatcher.cpp:1137-1153) and cv::BFMatcher(NORM_HAMMING).knnMatch(k=2)
// as called by LSDmatcher::matchNNR / FrameBFMatch (src/LSDmatcher.cpp:803-826, 942-966).
//
//   k_knn2_partial  one thread per query (8 x u32 in registers), the CTA's slice of the train set staged in
//                   shared memory and read as warp-wide broadcasts; XOR + __popc, running best / second with
//                   BFMatcher's tie rule (lower train index first)
//   k_knn2_merge    merges the per-slice (best, second) pairs in slice order
//
// Not a dense contraction in the north-star's sense: the integer pipe (POPC) is the roofline here.
#include <algorithm>
#include <climits>
#include <new>

#include "hvo_common.cuh"

namespace hvo {

static const int kKnnThreads = 128;
static const int kKnnTile = 512;  // train descriptors staged per shared-memory tile (16 KB)

struct Top2 { int d0, i0, d1, i1; };

__device__ __forceinline__ void top2_push(Top2& t, int d, int i) {
    // candidates arrive in increasing train index: strict '<' keeps the lower index on ties
    if (d < t.d0) { t.d1 = t.d0; t.i1 = t.i0; t.d0 = d; t.i0 = i; }
    else if (d < t.d1) { t.d1 = d; t.i1 = i; }
}

__global__ void __launch_bounds__(kKnnThreads) k_knn2_partial(const uint32_t* __restrict__ q, int nq,
                                                              const uint32_t* __restrict__ t, int nt, int slice_len,
                                                              int4* __restrict__ partial) {
    __shared__ uint4 s_t[kKnnTile * 2];
    const int qi = blockIdx.x * kKnnThreads + threadIdx.x;
    const int slice = blockIdx.y;
    const int t_begin = slice * slice_len, t_end = min(t_begin + slice_len, nt);
    uint32_t a[8];
    if (qi < nq) {
        const uint4 lo = __ldg(reinterpret_cast<const uint4*>(q) + 2 * qi), hi = __ldg(reinterpret_cast<const uint4*>(q) + 2 * qi + 1);
        a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w; a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = 0;
    }
    Top2 best = {257, -1, 257, -1};
    for (int base = t_begin; base < t_end; base += kKnnTile) {
        const int n = min(kKnnTile, t_end - base);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * n; i += kKnnThreads) s_t[i] = __ldg(reinterpret_cast<const uint4*>(t) + 2 * base + i);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const uint4 lo = s_t[2 * j], hi = s_t[2 * j + 1];
            const int d = __popc(a[0] ^ lo.x) + __popc(a[1] ^ lo.y) + __popc(a[2] ^ lo.z) + __popc(a[3] ^ lo.w) +
                          __popc(a[4] ^ hi.x) + __popc(a[5] ^ hi.y) + __popc(a[6] ^ hi.z) + __popc(a[7] ^ hi.w);
            top2_push(best, d, base + j);
        }
    }
    if (qi < nq) partial[(long long)slice * nq + qi] = make_int4(best.d0, best.i0, best.d1, best.i1);
}

__global__ void k_knn2_merge(const int4* __restrict__ partial, int nq, int nslices, int32_t* __restrict__ idx2,
                             int32_t* __restrict__ dist2) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    Top2 best = {257, -1, 257, -1};
    for (int s = 0; s < nslices; ++s) {  // slices are in increasing train-index order
        const int4 p = partial[(long long)s * nq + qi];
        if (p.y >= 0) top2_push(best, p.x, p.y);
        if (p.w >= 0) top2_push(best, p.z, p.w);
    }
    idx2[2 * qi] = best.i0; idx2[2 * qi + 1] = best.i1;
    dist2[2 * qi] = best.i0 >= 0 ? best.d0 : -1; dist2[2 * qi + 1] = best.i1 >= 0 ? best.d1 : -1;
}

// MapPoint / MapLine::ComputeDistinctiveDescriptors (reference src/MapPoint.cc:240-300, src/MapLine.cpp:331-400): per map element
// the descriptor with the least median Hamming distance to the other observations.  One warp per element: row i of the
// distance matrix goes into a 257-bin histogram in shared memory (distances are integers <= 256), the median
// sorted[int(0.5 * (n - 1))] is the first bin whose running count exceeds k; rows in order, strict '<' keeps the first minimum.
static const int kDistWarps = 4, kDistBins = 288;  // 9 bins per lane

__global__ void __launch_bounds__(kDistWarps * 32) k_distinctive(const uint4* __restrict__ desc, const int* __restrict__ off, int ngroups,
                                                                  int* __restrict__ best_idx, int* __restrict__ best_median) {
    __shared__ int s_hist[kDistWarps][kDistBins];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kDistWarps + w;
    if (g >= ngroups) return;
    const int b = off[g], n = off[g + 1] - b;
    int* hist = s_hist[w];
    int bestMedian = INT_MAX, bestIdx = n > 0 ? 0 : -1;
    const int k = (n - 1) >> 1;  // int(0.5 * (N - 1))
    for (int i = 0; i < n; ++i) {
        for (int t = lane; t < kDistBins; t += 32) hist[t] = 0;
        __syncwarp();
        const uint4 a0 = desc[2 * (size_t)(b + i)], a1 = desc[2 * (size_t)(b + i) + 1];
        for (int j = lane; j < n; j += 32) {
            const uint4 c0 = desc[2 * (size_t)(b + j)], c1 = desc[2 * (size_t)(b + j) + 1];
            const int d = __popc(a0.x ^ c0.x) + __popc(a0.y ^ c0.y) + __popc(a0.z ^ c0.z) + __popc(a0.w ^ c0.w) + __popc(a1.x ^ c1.x) +
                          __popc(a1.y ^ c1.y) + __popc(a1.z ^ c1.z) + __popc(a1.w ^ c1.w);
            atomicAdd(&hist[d], 1);
        }
        __syncwarp();
        int mine = 0;
#pragma unroll
        for (int t = 0; t < 9; ++t) mine += hist[lane * 9 + t];
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
        const unsigned reach = __ballot_sync(0xffffffffu, inc > k);
        const int src = __ffs(reach) - 1;  // first lane whose cumulative count exceeds k
        int median = 0;
        if (lane == src) {
            int c = inc - mine;
            for (int t = 0; t < 9; ++t) { c += hist[lane * 9 + t]; if (c > k) { median = lane * 9 + t; break; } }
        }
        median = __shfl_sync(0xffffffffu, median, src);
        if (median < bestMedian) { bestMedian = median; bestIdx = i; }
        __syncwarp();
    }
    if (lane == 0) { best_idx[g] = bestIdx; best_median[g] = n > 0 ? bestMedian : -1; }
}

}  // namespace hvo

using namespace hvo;

struct hvo_matcher {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t tev[2] = {nullptr, nullptr};
    int4* d_partial = nullptr;
    size_t partial_cap = 0;
    uint8_t *d_q = nullptr, *d_t = nullptr;
    int32_t *d_idx = nullptr, *d_dist = nullptr;
    size_t q_cap = 0, t_cap = 0;
    int sm_count = 148;
    int last_launches = 0;
};

static int knn2_launch(hvo_matcher* m, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int32_t* d_idx, int32_t* d_dist) {
    const int qblocks = div_up(nq, kKnnThreads);
    // enough slices to give every SM a few CTAs, each slice a multiple of the shared-memory tile
    int nslices = std::max(1, std::min(div_up(nt, kKnnTile), div_up(4 * m->sm_count, qblocks)));
    int slice_len = (int)align_up((size_t)div_up(nt, nslices), kKnnTile);
    nslices = div_up(nt, slice_len);
    const size_t need = (size_t)nslices * nq;
    if (need > m->partial_cap) {
        if (m->d_partial) cudaFree(m->d_partial);
        m->d_partial = nullptr; m->partial_cap = 0;
        HVO_CUDA(cudaMalloc(&m->d_partial, need * sizeof(int4)));
        m->partial_cap = need;
    }
    k_knn2_partial<<<dim3(qblocks, nslices), kKnnThreads, 0, m->stream>>>(reinterpret_cast<const uint32_t*>(d_q), nq,
                                                                          reinterpret_cast<const uint32_t*>(d_t), nt, slice_len, m->d_partial);
    k_knn2_merge<<<div_up(nq, 256), 256, 0, m->stream>>>(m->d_partial, nq, nslices, d_idx, d_dist);
    m->last_launches = 2;
    HVO_CUDA(cudaGetLastError());
    return HVO_OK;
}

extern "C" {

int hvo_matcher_create(int device, hvo_matcher** out) {
    HVO_CHECK_ARG(out, "null out");
    *out = nullptr;
    int ndev = 0;
    HVO_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev < 1) { set_error("no CUDA device: libhvofront has no CPU fallback"); return HVO_ERR_CUDA; }
    HVO_CHECK_ARG(device >= 0 && device < ndev, "device index out of range");
    hvo_matcher* m = new (std::nothrow) hvo_matcher();
    if (!m) { set_error("out of host memory"); return HVO_ERR_ARG; }
    m->device = device;
    HVO_CUDA(cudaSetDevice(device));
    HVO_CUDA(create_stream(&m->stream));
    for (auto& e : m->tev) HVO_CUDA(cudaEventCreate(&e));
    HVO_CUDA(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device));
    *out = m;
    return HVO_OK;
}

void hvo_matcher_destroy(hvo_matcher* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    void* bufs[] = {m->d_partial, m->d_q, m->d_t, m->d_idx, m->d_dist};
    for (void* b : bufs) if (b) cudaFree(b);
    for (auto& e : m->tev) if (e) cudaEventDestroy(e);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

int hvo_match_knn2_device(hvo_matcher* m, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int32_t* d_idx2, int32_t* d_dist2) {
    HVO_CHECK_ARG(m && d_q && d_t && d_idx2 && d_dist2, "null argument");
    HVO_CHECK_ARG(nq >= 1 && nt >= 1, "empty descriptor set");
    HVO_CHECK_ARG(((uintptr_t)d_q & 15) == 0 && ((uintptr_t)d_t & 15) == 0, "descriptor arrays must be 16-byte aligned");
    HVO_CUDA(cudaSetDevice(m->device));
    return knn2_launch(m, d_q, nq, d_t, nt, d_idx2, d_dist2);
}

int hvo_match_distinctive(hvo_matcher* m, const uint8_t* desc, const int32_t* offsets, int ngroups, int32_t* best_idx, int32_t* best_median) {
    HVO_CHECK_ARG(m && best_idx, "null argument");
    if (ngroups <= 0) return HVO_OK;
    HVO_CHECK_ARG(offsets, "null offsets");
    const int total = offsets[ngroups];
    HVO_CHECK_ARG(offsets[0] == 0 && total >= 0 && (total == 0 || desc), "bad offsets / null descriptors");
    for (int g = 0; g < ngroups; ++g) HVO_CHECK_ARG(offsets[g + 1] >= offsets[g], "offsets must be non-decreasing");
    HVO_CUDA(cudaSetDevice(m->device));
    uint8_t* d_desc = nullptr;
    int32_t *d_off = nullptr, *d_out = nullptr;
    HVO_CUDA(cudaMallocAsync(&d_desc, std::max<size_t>((size_t)total * 32, 32), m->stream));
    HVO_CUDA(cudaMallocAsync(&d_off, ((size_t)ngroups + 1) * 4, m->stream));
    HVO_CUDA(cudaMallocAsync(&d_out, (size_t)ngroups * 8, m->stream));
    if (total > 0) HVO_CUDA(cudaMemcpyAsync(d_desc, desc, (size_t)total * 32, cudaMemcpyHostToDevice, m->stream));
    HVO_CUDA(cudaMemcpyAsync(d_off, offsets, ((size_t)ngroups + 1) * 4, cudaMemcpyHostToDevice, m->stream));
    k_distinctive<<<div_up(ngroups, kDistWarps), kDistWarps * 32, 0, m->stream>>>(reinterpret_cast<const uint4*>(d_desc), d_off, ngroups, d_out,
                                                                               d_out + ngroups);
    HVO_CUDA(cudaGetLastError());
    HVO_CUDA(cudaMemcpyAsync(best_idx, d_out, (size_t)ngroups * 4, cudaMemcpyDeviceToHost, m->stream));
    if (best_median) HVO_CUDA(cudaMemcpyAsync(best_median, d_out + ngroups, (size_t)ngroups * 4, cudaMemcpyDeviceToHost, m->stream));
    HVO_CUDA(cudaFreeAsync(d_desc, m->stream));
    HVO_CUDA(cudaFreeAsync(d_off, m->stream));
    HVO_CUDA(cudaFreeAsync(d_out, m->stream));
    HVO_CUDA(cudaStreamSynchronize(m->stream));
    m->last_launches = 1;
    return HVO_OK;
}

int hvo_match_knn2(hvo_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx2, int32_t* dist2) {
    HVO_CHECK_ARG(m && idx2 && dist2, "null argument");
    if (nq <= 0) return HVO_OK;  // no queries: nothing to do (BFMatcher returns an empty list)
    HVO_CHECK_ARG(q, "null query descriptors");
    if (nt <= 0 || t == nullptr) {  // empty train set: no neighbours
        for (int i = 0; i < 2 * nq; ++i) { idx2[i] = -1; dist2[i] = -1; }
        return HVO_OK;
    }
    HVO_CUDA(cudaSetDevice(m->device));
    if ((size_t)nq > m->q_cap) {
        if (m->d_q) cudaFree(m->d_q);
        if (m->d_idx) cudaFree(m->d_idx);
        if (m->d_dist) cudaFree(m->d_dist);
        m->d_q = nullptr; m->d_idx = nullptr; m->d_dist = nullptr; m->q_cap = 0;
        HVO_CUDA(cudaMalloc(&m->d_q, (size_t)nq * 32));
        HVO_CUDA(cudaMalloc(&m->d_idx, (size_t)nq * 8));
        HVO_CUDA(cudaMalloc(&m->d_dist, (size_t)nq * 8));
        m->q_cap = nq;
    }
    if ((size_t)nt > m->t_cap) {
        if (m->d_t) cudaFree(m->d_t);
        m->d_t = nullptr; m->t_cap = 0;
        HVO_CUDA(cudaMalloc(&m->d_t, (size_t)nt * 32));
        m->t_cap = nt;
    }
    HVO_CUDA(cudaMemcpyAsync(m->d_q, q, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream));
    HVO_CUDA(cudaMemcpyAsync(m->d_t, t, (size_t)nt * 32, cudaMemcpyHostToDevice, m->stream));
    int st = knn2_launch(m, m->d_q, nq, m->d_t, nt, m->d_idx, m->d_dist);
    if (st != HVO_OK) return st;
    HVO_CUDA(cudaMemcpyAsync(idx2, m->d_idx, (size_t)nq * 8, cudaMemcpyDeviceToHost, m->stream));
    HVO_CUDA(cudaMemcpyAsync(dist2, m->d_dist, (size_t)nq * 8, cudaMemcpyDeviceToHost, m->stream));
    HVO_CUDA(cudaStreamSynchronize(m->stream));
    return HVO_OK;
}

int hvo_matcher_sync(hvo_matcher* m) {
    HVO_CHECK_ARG(m, "null handle");
    HVO_CUDA(cudaSetDevice(m->device));
    HVO_CUDA(cudaStreamSynchronize(m->stream));
    return HVO_OK;
}
int hvo_matcher_timer_start(hvo_matcher* m) {
    HVO_CHECK_ARG(m, "null handle");
    HVO_CUDA(cudaSetDevice(m->device));
    HVO_CUDA(cudaEventRecord(m->tev[0], m->stream));
    return HVO_OK;
}
int hvo_matcher_timer_stop(hvo_matcher* m, float* ms_out) {
    HVO_CHECK_ARG(m && ms_out, "null argument");
    HVO_CUDA(cudaSetDevice(m->device));
    HVO_CUDA(cudaEventRecord(m->tev[1], m->stream));
    HVO_CUDA(cudaEventSynchronize(m->tev[1]));
    HVO_CUDA(cudaEventElapsedTime(ms_out, m->tev[0], m->tev[1]));
    return HVO_OK;
}

/* ORBmatcher::DescriptorDistance / LSDmatcher::DescriptorDistance: host helper, 8 x 32-bit popcount */
int hvo_hamming_distance(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; i += 4) {
        uint32_t x, y;
        memcpy(&x, a + i, 4);
        memcpy(&y, b + i, 4);
        d += __builtin_popcount(x ^ y);
    }
    return d;
}

}  // extern "C"
