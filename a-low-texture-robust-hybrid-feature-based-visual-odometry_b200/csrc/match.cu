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

#include <vector>

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


// ---- LSDmatcher::FrameBFMatchNew's epipolar test on the nearest neighbour (src/LSDmatcher.cpp:983-1029) + mutualOverlap (:1033-1108) ----
// One thread per query line.  cv::Mat arithmetic as OpenCV evaluates it for CV_32F: F * p = float products summed in k order;
// Mat::cross in float; `m /= s` = m * (float)(1. / s); cv::norm of a difference = sqrt of the double sum of squares, narrowed to float.
struct F33 { float m[9]; };
__device__ __forceinline__ void epi_mul(const F33& F, float x, float y, float* o) {
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = __fadd_rn(__fadd_rn(__fmul_rn(F.m[3 * i], x), __fmul_rn(F.m[3 * i + 1], y)), __fmul_rn(F.m[3 * i + 2], 1.0f));
}
__device__ __forceinline__ void epi_cross(const float* a, const float* b, float* c) {   // a.cross(b)
    c[0] = __fsub_rn(__fmul_rn(a[1], b[2]), __fmul_rn(a[2], b[1]));
    c[1] = __fsub_rn(__fmul_rn(a[2], b[0]), __fmul_rn(a[0], b[2]));
    c[2] = __fsub_rn(__fmul_rn(a[0], b[1]), __fmul_rn(a[1], b[0]));
}
__device__ __forceinline__ float epi_dist(const float* a, const float* b) {
    const float dx = __fsub_rn(a[0], b[0]), dy = __fsub_rn(a[1], b[1]), dz = __fsub_rn(a[2], b[2]);
    return (float)sqrt(__dadd_rn(__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)), __dmul_rn((double)dz, (double)dz)));
}
__device__ float epi_mutual_overlap(const float (*pt)[3]) {
    float max_dist = 0.0f;
    int outer1 = 0, outer2 = 3, inner1, inner2;
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 4; ++j) {
            const float d = epi_dist(pt[i], pt[j]);
            if (d > max_dist) { max_dist = d; outer1 = i; outer2 = j; }
        }
    if (max_dist < 1.0f) return 0.0f;
    if (outer1 == 0) {
        if (outer2 == 1) { inner1 = 2; inner2 = 3; }
        else if (outer2 == 2) { inner1 = 1; inner2 = 3; }
        else { inner1 = 1; inner2 = 2; }
    } else if (outer1 == 1) {
        inner1 = 0;
        inner2 = outer2 == 2 ? 3 : 2;
    } else { inner1 = 0; inner2 = 1; }
    const float dx = __fsub_rn(pt[inner1][0], pt[inner2][0]), dy = __fsub_rn(pt[inner1][1], pt[inner2][1]), dz = __fsub_rn(pt[inner1][2], pt[inner2][2]);
    const double nrm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)), __dmul_rn((double)dz, (double)dz)));
    return (float)(nrm / (double)max_dist);
}
__global__ void k_lines_epipolar(const int32_t* __restrict__ idx2, const int32_t* __restrict__ dist2, const hvo_keyline* __restrict__ kls1, int n1,
                                 const hvo_keyline* __restrict__ kls2, const double* __restrict__ func2, F33 F, float th, float nnratio,
                                 int32_t* __restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n1) return;
    int res = -1;
    const int t = idx2[2 * q], t1 = idx2[2 * q + 1];
    if (t >= 0 && t1 >= 0) {
        const hvo_keyline a = kls1[q], b = kls2[t];
        float e1[3], e2[3], pt[4][3];
        epi_mul(F, a.startPointX, a.startPointY, e1);
        epi_mul(F, a.endPointX, a.endPointY, e2);
        const float l2[3] = {(float)func2[3 * t], (float)func2[3 * t + 1], (float)func2[3 * t + 2]};
        epi_cross(l2, e1, pt[0]);
        epi_cross(l2, e2, pt[1]);
        if (fabs((double)pt[0][2]) > 1e-12 && fabs((double)pt[1][2]) > 1e-12) {
            const float s0 = (float)(1.0 / (double)pt[0][2]), s1 = (float)(1.0 / (double)pt[1][2]);
#pragma unroll
            for (int i = 0; i < 3; ++i) { pt[0][i] = __fmul_rn(pt[0][i], s0); pt[1][i] = __fmul_rn(pt[1][i], s1); }
            pt[2][0] = b.startPointX; pt[2][1] = b.startPointY; pt[2][2] = 1.0f;
            pt[3][0] = b.endPointX; pt[3][1] = b.endPointY; pt[3][2] = 1.0f;
            const float score = epi_mutual_overlap(pt);
            const float d0 = (float)dist2[2 * q], d1 = (float)dist2[2 * q + 1];
            if (d0 < th && (double)score > 0.8 && d0 < __fmul_rn(nnratio, d1)) res = t;
        }
    }
    out[q] = res;
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
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = create_stream(&m->stream);
    for (auto& ev : m->tev) if (e == cudaSuccess) e = cudaEventCreate(&ev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {   // one exit: nothing of the partial handle is leaked
        set_error("hvo_matcher_create: %s", cudaGetErrorString(e));
        hvo_matcher_destroy(m);
        return HVO_ERR_CUDA;
    }
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

int hvo_match_lines_epipolar(hvo_matcher* m, const uint8_t* ldesc1, const hvo_keyline* kls1, int n1, const uint8_t* ldesc2, const hvo_keyline* kls2,
                             const double* kls2func, int n2, const float* F, float th, float nnratio, int32_t* line_matches) {
    HVO_CHECK_ARG(m && line_matches && F, "null argument");
    if (n1 <= 0) return HVO_OK;
    for (int i = 0; i < n1; ++i) line_matches[i] = -1;
    if (n2 < 2) return HVO_OK;          // knnMatch yields fewer than two neighbours: the reference's loop body never accepts
    HVO_CHECK_ARG(ldesc1 && kls1 && ldesc2 && kls2 && kls2func, "null descriptors / keylines / line functions");
    std::vector<int32_t> idx((size_t)n1 * 2), dist((size_t)n1 * 2);
    int st = hvo_match_knn2(m, ldesc1, n1, ldesc2, n2, idx.data(), dist.data());   // leaves idx / dist of this call in m->d_idx / m->d_dist
    if (st != HVO_OK) return st;
    hvo_keyline *d_k1 = nullptr, *d_k2 = nullptr;
    double* d_f = nullptr;
    int32_t* d_out = nullptr;
    cudaStream_t s = m->stream;
    HVO_CUDA(cudaMallocAsync(&d_k1, (size_t)n1 * sizeof(hvo_keyline), s));
    HVO_CUDA(cudaMallocAsync(&d_k2, (size_t)n2 * sizeof(hvo_keyline), s));
    HVO_CUDA(cudaMallocAsync(&d_f, (size_t)n2 * 3 * sizeof(double), s));
    HVO_CUDA(cudaMallocAsync(&d_out, (size_t)n1 * sizeof(int32_t), s));
    HVO_CUDA(cudaMemcpyAsync(d_k1, kls1, (size_t)n1 * sizeof(hvo_keyline), cudaMemcpyHostToDevice, s));
    HVO_CUDA(cudaMemcpyAsync(d_k2, kls2, (size_t)n2 * sizeof(hvo_keyline), cudaMemcpyHostToDevice, s));
    HVO_CUDA(cudaMemcpyAsync(d_f, kls2func, (size_t)n2 * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    F33 f;
    for (int i = 0; i < 9; ++i) f.m[i] = F[i];
    k_lines_epipolar<<<div_up(n1, 128), 128, 0, s>>>(m->d_idx, m->d_dist, d_k1, n1, d_k2, d_f, f, th, nnratio, d_out);
    HVO_CUDA(cudaGetLastError());
    HVO_CUDA(cudaMemcpyAsync(line_matches, d_out, (size_t)n1 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    HVO_CUDA(cudaFreeAsync(d_k1, s));
    HVO_CUDA(cudaFreeAsync(d_k2, s));
    HVO_CUDA(cudaFreeAsync(d_f, s));
    HVO_CUDA(cudaFreeAsync(d_out, s));
    HVO_CUDA(cudaStreamSynchronize(s));
    m->last_launches = 3;
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
