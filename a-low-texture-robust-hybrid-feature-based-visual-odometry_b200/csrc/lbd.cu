// LBD line-band descriptor for sm_100a.  Replaces cv::line_descriptor::BinaryDescriptor::compute(image, keylines,
// descriptors) as called by LINEextractor::operator() (reference src/LineExtractor.cpp:361-363) and by
// Frame::cullingLine (src/Frame.cc:1094-1096); algorithm = Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp
// (:350-398 gradients, :539-687 computeImpl, :1026-1372 computeLBD, :401-412 binaryConversion).
//
//   k_lbd_grad      fused 5x5 sigma-1 Q8 Gaussian (cv::GaussianBlur fixed point) + Sobel 3x3 -> int16 dx, dy;
//                   one shared-memory tile with a 3-px halo, nothing but dx/dy goes to HBM
//   k_lbd_describe  one CTA per line: thread <-> row of the 63-row line support region, each walks its row
//                   sequentially (float32, reference operation order, so the binary tests cannot flip), then
//                   72 threads fold rows into the 9 bands in row order, then mean/std, the two normalisations
//                   and the 32 x 8 band-pair comparisons
#include <cmath>
#include <new>

#include "hvo_common.cuh"

namespace hvo {

static const int kLbdBands = 9, kLbdBandW = 7, kLbdH = 63;
static const int kGW = 64, kGH = 32;  // gradient tile

struct LbdWeights { float G[63]; float L[21]; };

__constant__ int8_t c_comb[64] = {0, 1, 0, 2, 0, 3, 0, 4, 0, 5, 0, 6, 1, 2, 1, 3, 1, 4, 1, 5, 1, 6, 2, 3, 2, 4, 2, 5, 2, 6, 2, 7,
                                  2, 8, 3, 4, 3, 5, 3, 6, 3, 7, 3, 8, 4, 5, 4, 6, 4, 7, 4, 8, 5, 6, 5, 7, 5, 8, 6, 7, 6, 8, 7, 8};

__global__ void __launch_bounds__(256) k_lbd_grad(const uint8_t* __restrict__ gray, int w, int h, long long frame_px,
                                                  uint32_t* __restrict__ dxy) {
    __shared__ __align__(4) uint8_t raw[(kGH + 6) * (kGW + 8)];
    __shared__ uint16_t hb[(kGH + 6) * (kGW + 2)];
    __shared__ uint8_t bl[(kGH + 2) * (kGW + 4)];
    const int tx = blockIdx.x * kGW, ty = blockIdx.y * kGH, f = blockIdx.z, tid = threadIdx.x;
    const uint8_t* img = gray + (long long)f * frame_px;
    // raw tile: rows ty-3..ty+kGH+2, cols tx-3..tx+kGW+2 (reflect-101; the blurred image's own reflection equals
    // the blur of the reflected input because the kernel is symmetric)
    // Interior tiles (no reflection; tx is a multiple of 64, so tx - 4 is word aligned when the rows are) come in as aligned
    // 32-bit words: raw col c then holds image col tx - 4 + c, one byte to the right of the general layout (`sh`).
    const bool interior = tx >= 4 && ty >= 3 && tx + kGW + 4 <= w && ty + kGH + 3 <= h && (w & 3) == 0 && (frame_px & 3) == 0 &&
                          (reinterpret_cast<uintptr_t>(gray) & 3) == 0;
    const int sh = interior ? 1 : 0;
    if (interior) {
        const uint8_t* base = img + (long long)(ty - 3) * w + (tx - 4);
        for (int i = tid; i < (kGH + 6) * ((kGW + 8) / 4); i += 256) {
            const int r = i / ((kGW + 8) / 4), c = i - r * ((kGW + 8) / 4);
            reinterpret_cast<uint32_t*>(raw + r * (kGW + 8))[c] = __ldg(reinterpret_cast<const uint32_t*>(base + (long long)r * w) + c);
        }
    } else {
        for (int i = tid; i < (kGH + 6) * (kGW + 6); i += 256) {
            const int r = i / (kGW + 6), c = i - r * (kGW + 6);
            const int y = min(max(reflect101(ty - 3 + r, h), 0), h - 1), x = min(max(reflect101(tx - 3 + c, w), 0), w - 1);
            raw[r * (kGW + 8) + c] = __ldg(img + (long long)y * w + x);
        }
    }
    __syncthreads();
    // horizontal 5-tap (14,62,104,62,14): blurred cols tx-1..tx+kGW  <->  index 0..kGW+1
    for (int i = tid; i < (kGH + 6) * (kGW + 2); i += 256) {
        const int r = i / (kGW + 2), c = i - r * (kGW + 2);
        const uint8_t* p = raw + r * (kGW + 8) + c + sh;  // taps at raw cols c..c+4  (centre c+2 <-> x = tx-1+c)
        hb[r * (kGW + 2) + c] = (uint16_t)(14 * (p[0] + p[4]) + 62 * (p[1] + p[3]) + 104 * p[2]);
    }
    __syncthreads();
    // vertical 5-tap: blurred rows ty-1..ty+kGH  <-> index 0..kGH+1 ; hb row r <-> y = ty-3+r
    for (int i = tid; i < (kGH + 2) * (kGW + 2); i += 256) {
        const int r = i / (kGW + 2), c = i - r * (kGW + 2);
        const uint16_t* p = hb + r * (kGW + 2) + c;  // rows r..r+4 (centre r+2 <-> y = ty-1+r)
        const uint32_t acc = 14u * (p[0] + p[4 * (kGW + 2)]) + 62u * (p[kGW + 2] + p[3 * (kGW + 2)]) + 104u * p[2 * (kGW + 2)];
        bl[r * (kGW + 4) + c] = (uint8_t)((acc + 32768u) >> 16);
    }
    __syncthreads();
    // The blurred image is reflected at the IMAGE border for Sobel.  Tile halo pixels that fall outside the image
    // were computed from reflected coordinates, which is the same value (see above).
    for (int i = tid; i < kGH * kGW; i += 256) {
        const int r = i / kGW, c = i - r * kGW;
        const int x = tx + c, y = ty + r;
        if (x >= w || y >= h) continue;
        const uint8_t* p = bl + (r + 1) * (kGW + 4) + (c + 1);
        const int S = kGW + 4;
        const int gx = (p[-S + 1] - p[-S - 1]) + 2 * (p[1] - p[-1]) + (p[S + 1] - p[S - 1]);
        const int gy = (p[S - 1] - p[-S - 1]) + 2 * (p[S] - p[-S]) + (p[S + 1] - p[-S + 1]);
        const long long o = (long long)f * frame_px + (long long)y * w + x;
        dxy[o] = ((uint32_t)gx & 0xffffu) | ((uint32_t)gy << 16);   // dx | dy << 16: one load per visited pixel in k_lbd_describe
    }
}

struct KeyLineDev {  // cv::line_descriptor::KeyLine POD, 68 bytes
    float angle;
    int class_id, octave;
    float pt_x, pt_y, response, size;
    float startPointX, startPointY, endPointX, endPointY;
    float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
    float lineLength;
    int numOfPixels;
};

__global__ void __launch_bounds__(64) k_lbd_describe(const uint32_t* __restrict__ dxyI, int w, int h,
                                                     long long frame_px, const KeyLineDev* __restrict__ kls,
                                                     const int32_t* __restrict__ counts, int max_lines, LbdWeights W,
                                                     float* __restrict__ raw) {
    __shared__ float rows[kLbdH][8];  // per row: pL, nL, pO, nO and their squares (after the global weight)
    const int line = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    if (line >= counts[f]) return;
    const KeyLineDev kl = kls[(long long)f * max_lines + line];
    const uint32_t* gp = dxyI + (long long)f * frame_px;
    const int len = (short)kl.numOfPixels;
    const float dL0 = (float)cos((double)kl.angle), dL1 = (float)sin((double)kl.angle);
    if (tid < kLbdH) {
        const int imageWidth = w - 1, imageHeight = h - 1;
        const short halfWidth = (short)((len - 1) / 2), halfHeight = (short)((kLbdH - 1) / 2);
        const float midX = 0.5f * (kl.sPointInOctaveX + kl.ePointInOctaveX);
        const float midY = 0.5f * (kl.sPointInOctaveY + kl.ePointInOctaveY);
        const float dO0 = -dL1, dO1 = dL0;
        float sCorX = -dL0 * halfWidth + dL1 * halfHeight + midX;
        float sCorY = -dL1 * halfWidth - dL0 * halfHeight + midY;
        for (int k = 0; k < tid; ++k) { sCorX -= dL1; sCorY += dL0; }  // the reference advances the row origin by repeated addition
        float pL = 0, nL = 0, pO = 0, nO = 0;
#pragma unroll 4
        for (int wID = 0; wID < len; ++wID) {
            int t = (int)(short)roundf(sCorX);
            const int xCor = t < 0 ? 0 : (t > imageWidth ? imageWidth : t);
            t = (int)(short)roundf(sCorY);
            const int yCor = t < 0 ? 0 : (t > imageHeight ? imageHeight : t);
            const uint32_t g2 = __ldg(gp + yCor * w + xCor);
            const int gx = (int)(short)(g2 & 0xffffu), gy = (int)g2 >> 16;
            const float gDL = (float)gx * dL0 + (float)gy * dL1;
            const float gDO = (float)gx * dO0 + (float)gy * dO1;
            if (gDL > 0) pL += gDL; else nL -= gDL;
            if (gDO > 0) pO += gDO; else nO -= gDO;
            sCorX += dL0;
            sCorY += dL1;
        }
        const float c = W.G[tid];
        pL = c * pL; nL = c * nL; pO = c * pO; nO = c * nO;
        rows[tid][0] = pL; rows[tid][1] = nL; rows[tid][2] = pO; rows[tid][3] = nO;
        rows[tid][4] = pL * pL; rows[tid][5] = nL * nL; rows[tid][6] = pO * pO; rows[tid][7] = nO * nO;
    }
    __syncthreads();
    for (int e = tid; e < 72; e += 64) {
        // band b, quantity q: rows contribute in increasing row order: rows of band b-1 (as "band below" of that row,
        // weight L[r % 7]), band b (L[r % 7 + 7]), band b+1 (as "band above", L[r % 7 + 14])
        const int b = e >> 3, q = e & 7;
        float acc = 0;
        const int r0 = max(0, (b - 1) * kLbdBandW), r1 = min(kLbdH, (b + 2) * kLbdBandW);
        for (int r = r0; r < r1; ++r) {
            const int rb = r / kLbdBandW;
            const float c = W.L[r % kLbdBandW + (rb == b ? kLbdBandW : (rb == b + 1 ? 2 * kLbdBandW : 0))];
            acc += (q < 4) ? c * rows[r][q] : c * c * rows[r][q];
        }
        // the 72 band sums of this line; the serial normalisation runs lane-parallel over lines in k_lbd_finish
        raw[((long long)f * max_lines + line) * 72 + e] = acc;
    }
}

// Second half of computeLBD (binary_descriptor_custom.cpp:1300-1372): mean / standard deviation per band, the two normalisations, the 0.4
// clamp, the final normalisation and the 32 comparison bytes.  Every sum is sequential in the reference, so it is one thread per line (the
// lanes of a warp work on 32 different lines) instead of one thread of a 96-thread CTA with the other 95 waiting at a barrier.
__global__ void __launch_bounds__(128) k_lbd_finish(const float* __restrict__ raw, const int32_t* __restrict__ counts, int max_lines, int nframes,
                                                    uint8_t* __restrict__ desc, float* __restrict__ fdesc) {
    __shared__ float sd[72][128];   // d[k] of thread t at sd[k][t]: conflict-free, dynamically indexable
    const long long i = (long long)blockIdx.x * 128 + threadIdx.x;
    if (i >= (long long)nframes * max_lines) return;
    const int f = (int)(i / max_lines), line = (int)(i - (long long)f * max_lines);
    if (line >= counts[f]) return;
    const float* des = raw + i * 72;
    float(*d)[128] = reinterpret_cast<float(*)[128]>(&sd[0][threadIdx.x]);   // d[k][0] == sd[k][tid]
    const float invN2 = (float)(1.0 / (kLbdBandW * 2.0)), invN3 = (float)(1.0 / (kLbdBandW * 3.0));
    for (int b = 0; b < kLbdBands; ++b) {
        const float invN = (b == 0 || b == kLbdBands - 1) ? invN2 : invN3;
        const float4 m4 = *reinterpret_cast<const float4*>(des + 8 * b), s4 = *reinterpret_cast<const float4*>(des + 8 * b + 4);
        const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float temp = mm[q] * invN;
            d[8 * b + q][0] = temp;
            d[8 * b + 4 + q][0] = sqrtf(ss[q] * invN - temp * temp);
        }
    }
    float tempM = 0, tempS = 0;
    for (int b = 0; b < kLbdBands; ++b) {
#pragma unroll
        for (int q = 0; q < 4; ++q) tempM += d[8 * b + q][0] * d[8 * b + q][0];
#pragma unroll
        for (int q = 4; q < 8; ++q) tempS += d[8 * b + q][0] * d[8 * b + q][0];
    }
    tempM = 1.f / sqrtf(tempM);
    tempS = 1.f / sqrtf(tempS);
    float temp = 0;
    for (int b = 0; b < kLbdBands; ++b)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float v = d[8 * b + q][0] * (q < 4 ? tempM : tempS);
            if (v > 0.4f) v = 0.4f;  // (double)x > 0.4  <=>  x >= 0.4f, and the clamp value is 0.4f
            d[8 * b + q][0] = v;
        }
    for (int k = 0; k < 72; ++k) temp += d[k][0] * d[k][0];
    temp = 1.f / sqrtf(temp);
    for (int k = 0; k < 72; ++k) d[k][0] = d[k][0] * temp;
    uint32_t words[8];
#pragma unroll
    for (int t = 0; t < 32; ++t) {
        const int i1 = 8 * c_comb[2 * t], i2 = 8 * c_comb[2 * t + 1];
        unsigned v = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) v |= (d[i1 + b][0] > d[i2 + b][0] ? 1u : 0u) << b;
        if ((t & 3) == 0) words[t >> 2] = v; else words[t >> 2] |= v << (8 * (t & 3));
    }
    uint4* o = reinterpret_cast<uint4*>(desc + i * 32);
    o[0] = make_uint4(words[0], words[1], words[2], words[3]);
    o[1] = make_uint4(words[4], words[5], words[6], words[7]);
    if (fdesc != nullptr)
        for (int k = 0; k < 72; ++k) fdesc[i * 72 + k] = d[k][0];
}

}  // namespace hvo

using namespace hvo;

struct hvo_lbd {
    int device = 0, width = 0, height = 0, max_batch = 0, max_lines = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t tev[2] = {nullptr, nullptr};
    LbdWeights W;
    uint8_t* d_gray = nullptr;
    uint32_t* d_dxy = nullptr;   // Sobel dx | dy << 16 of the blurred frame
    KeyLineDev* d_kl = nullptr;
    int32_t* d_counts = nullptr;
    uint8_t* d_desc = nullptr;
    float* d_fdesc = nullptr;
    float* d_raw = nullptr;   // [B][max_lines][72] band sums between k_lbd_describe and k_lbd_finish
    int last_launches = 0;
};

static int lbd_run_stream(hvo_lbd* h, cudaStream_t stream, const uint8_t* d_gray, int nframes, const KeyLineDev* d_kl,
                          const int32_t* d_counts, uint8_t* d_desc, float* d_fdesc) {
    const long long fpx = (long long)h->width * h->height;
    timeline_mark(stream, "k_lbd_grad");
    k_lbd_grad<<<dim3(div_up(h->width, kGW), div_up(h->height, kGH), nframes), 256, 0, stream>>>(d_gray, h->width, h->height, fpx,
                                                                                              h->d_dxy);
    timeline_mark(stream, "k_lbd_describe");
    k_lbd_describe<<<dim3(h->max_lines, nframes), 64, 0, stream>>>(h->d_dxy, h->width, h->height, fpx, d_kl, d_counts,
                                                                   h->max_lines, h->W, h->d_raw);
    timeline_mark(stream, "k_lbd_finish");
    k_lbd_finish<<<div_up(nframes * h->max_lines, 128), 128, 0, stream>>>(h->d_raw, d_counts, h->max_lines, nframes, d_desc, d_fdesc);
    h->last_launches = 3;
    HVO_CUDA(cudaGetLastError());
    return HVO_OK;
}
static int lbd_run(hvo_lbd* h, const uint8_t* d_gray, int nframes, const KeyLineDev* d_kl, const int32_t* d_counts, uint8_t* d_desc,
                   float* d_fdesc) {
    return lbd_run_stream(h, h->stream, d_gray, nframes, d_kl, d_counts, d_desc, d_fdesc);
}

// Internal (lsd.cu): LBD of device-resident keylines on the caller's stream (the line extractor chains it after LSD).
namespace hvo {
int lbd_compute_on_stream(hvo_lbd* h, cudaStream_t stream, const uint8_t* d_gray, int nframes, const hvo_keyline* d_keylines,
                          const int32_t* d_counts, uint8_t* d_desc) {
    return lbd_run_stream(h, stream, d_gray, nframes, reinterpret_cast<const KeyLineDev*>(d_keylines), d_counts, d_desc, nullptr);
}
}  // namespace hvo

extern "C" {

int hvo_lbd_create(int width, int height, int max_batch, int max_lines, int device, hvo_lbd** out) {
    HVO_CHECK_ARG(out, "null out");
    *out = nullptr;
    HVO_CHECK_ARG(width >= 8 && height >= 8 && width <= 32767 && height <= 32767, "image size out of range");
    HVO_CHECK_ARG(max_batch >= 1 && max_lines >= 1, "max_batch / max_lines < 1");
    int ndev = 0;
    HVO_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev < 1) { set_error("no CUDA device: libhvofront has no CPU fallback"); return HVO_ERR_CUDA; }
    HVO_CHECK_ARG(device >= 0 && device < ndev, "device index out of range");
    hvo_lbd* h = new (std::nothrow) hvo_lbd();
    if (!h) { set_error("out of host memory"); return HVO_ERR_ARG; }
    h->device = device; h->width = width; h->height = height; h->max_batch = max_batch; h->max_lines = max_lines;
    // weights: binary_descriptor_custom.cpp:228-258 (integer divisions kept)
    {
        double u = (kLbdBandW * 3 - 1) / 2, sigma = (kLbdBandW * 2 + 1) / 2, inv = -1 / (2 * sigma * sigma);
        for (int i = 0; i < 21; ++i) { const double d = i - u; h->W.L[i] = (float)std::exp(d * d * inv); }
        u = (kLbdBands * kLbdBandW - 1) / 2; sigma = u; inv = -1 / (2 * sigma * sigma);
        for (int i = 0; i < 63; ++i) { const double d = i - u; h->W.G[i] = (float)std::exp(d * d * inv); }
    }
    int st = HVO_OK;
    do {
#define HVO_TRY(call) if ((call) != cudaSuccess) { set_error("%s: %s", #call, cudaGetErrorString(cudaGetLastError())); st = HVO_ERR_CUDA; break; }
        HVO_TRY(cudaSetDevice(device));
        pin_carveout(k_lbd_grad); pin_carveout(k_lbd_describe);
        HVO_TRY(create_stream(&h->stream));
        HVO_TRY(cudaEventCreate(&h->tev[0]));
        HVO_TRY(cudaEventCreate(&h->tev[1]));
        const size_t B = (size_t)max_batch, px = (size_t)width * height;
        HVO_TRY(cudaMalloc(&h->d_gray, B * px));
        HVO_TRY(cudaMalloc(&h->d_dxy, B * px * 4));
        HVO_TRY(cudaMalloc(&h->d_kl, B * max_lines * sizeof(KeyLineDev)));
        HVO_TRY(cudaMalloc(&h->d_counts, B * sizeof(int32_t)));
        HVO_TRY(cudaMalloc(&h->d_desc, B * max_lines * 32));
        HVO_TRY(cudaMalloc(&h->d_fdesc, B * max_lines * 72 * sizeof(float)));
        HVO_TRY(cudaMalloc(&h->d_raw, B * max_lines * 72 * sizeof(float)));
#undef HVO_TRY
    } while (0);
    if (st != HVO_OK) { hvo_lbd_destroy(h); return st; }
    *out = h;
    return HVO_OK;
}

void hvo_lbd_destroy(hvo_lbd* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    void* bufs[] = {h->d_gray, h->d_dxy, h->d_kl, h->d_counts, h->d_desc, h->d_fdesc, h->d_raw};
    for (void* b : bufs) if (b) cudaFree(b);
    for (auto& e : h->tev) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int hvo_lbd_compute_batch(hvo_lbd* h, const uint8_t* gray, int nframes, const hvo_keyline* keylines, const int32_t* counts,
                          uint8_t* desc, float* fdesc) {
    HVO_CHECK_ARG(h && gray && keylines && counts && desc, "null argument");
    HVO_CHECK_ARG(nframes >= 1 && nframes <= h->max_batch, "nframes out of range for this handle");
    for (int f = 0; f < nframes; ++f) HVO_CHECK_ARG(counts[f] >= 0 && counts[f] <= h->max_lines, "line count out of range");
    HVO_CUDA(cudaSetDevice(h->device));
    const size_t n = (size_t)nframes, px = (size_t)h->width * h->height, ml = (size_t)h->max_lines;
    HVO_CUDA(cudaMemcpyAsync(h->d_gray, gray, n * px, cudaMemcpyHostToDevice, h->stream));
    HVO_CUDA(cudaMemcpyAsync(h->d_kl, keylines, n * ml * sizeof(KeyLineDev), cudaMemcpyHostToDevice, h->stream));
    HVO_CUDA(cudaMemcpyAsync(h->d_counts, counts, n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    int st = lbd_run(h, h->d_gray, nframes, h->d_kl, h->d_counts, h->d_desc, fdesc ? h->d_fdesc : nullptr);
    if (st != HVO_OK) return st;
    HVO_CUDA(cudaMemcpyAsync(desc, h->d_desc, n * ml * 32, cudaMemcpyDeviceToHost, h->stream));
    if (fdesc) HVO_CUDA(cudaMemcpyAsync(fdesc, h->d_fdesc, n * ml * 72 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    HVO_CUDA(cudaStreamSynchronize(h->stream));
    return HVO_OK;
}

int hvo_lbd_compute(hvo_lbd* h, const uint8_t* gray, size_t stride, const hvo_keyline* keylines, int n, uint8_t* desc) {
    HVO_CHECK_ARG(h && desc, "null argument");
    if (n == 0) { set_error("keypoint list is empty"); return HVO_OK; }  // the reference prints and returns (binary_descriptor_custom.cpp:556-560)
    HVO_CHECK_ARG(gray && keylines, "null argument");
    HVO_CHECK_ARG(n > 0 && n <= h->max_lines, "line count out of range for this handle");
    HVO_CHECK_ARG(stride >= (size_t)h->width, "stride smaller than width");
    HVO_CUDA(cudaSetDevice(h->device));
    const int32_t cnt = n;
    HVO_CUDA(cudaMemcpy2DAsync(h->d_gray, h->width, gray, stride, h->width, h->height, cudaMemcpyHostToDevice, h->stream));
    HVO_CUDA(cudaMemcpyAsync(h->d_kl, keylines, (size_t)n * sizeof(KeyLineDev), cudaMemcpyHostToDevice, h->stream));
    HVO_CUDA(cudaMemcpyAsync(h->d_counts, &cnt, sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    int st = lbd_run(h, h->d_gray, 1, h->d_kl, h->d_counts, h->d_desc, nullptr);
    if (st != HVO_OK) return st;
    HVO_CUDA(cudaMemcpyAsync(desc, h->d_desc, (size_t)n * 32, cudaMemcpyDeviceToHost, h->stream));
    HVO_CUDA(cudaStreamSynchronize(h->stream));
    return HVO_OK;
}

int hvo_lbd_compute_batch_device(hvo_lbd* h, const uint8_t* d_gray, int nframes, const hvo_keyline* d_keylines,
                                 const int32_t* d_counts, uint8_t* d_desc) {
    HVO_CHECK_ARG(h && d_gray && d_keylines && d_counts && d_desc, "null argument");
    HVO_CHECK_ARG(nframes >= 1 && nframes <= h->max_batch, "nframes out of range for this handle");
    HVO_CUDA(cudaSetDevice(h->device));
    return lbd_run(h, d_gray, nframes, reinterpret_cast<const KeyLineDev*>(d_keylines), d_counts, d_desc, nullptr);
}

int hvo_lbd_get_gradients(hvo_lbd* h, int frame, int16_t* dx, int16_t* dy) {
    HVO_CHECK_ARG(h && dx && dy, "null argument");
    HVO_CHECK_ARG(frame >= 0 && frame < h->max_batch, "frame out of range");
    HVO_CUDA(cudaSetDevice(h->device));
    const size_t px = (size_t)h->width * h->height;
    // the device keeps dx | dy << 16 per pixel: strided 2-byte copies split it
    HVO_CUDA(cudaMemcpy2DAsync(dx, 2, reinterpret_cast<const char*>(h->d_dxy + frame * px), 4, 2, px, cudaMemcpyDeviceToHost, h->stream));
    HVO_CUDA(cudaMemcpy2DAsync(dy, 2, reinterpret_cast<const char*>(h->d_dxy + frame * px) + 2, 4, 2, px, cudaMemcpyDeviceToHost, h->stream));
    HVO_CUDA(cudaStreamSynchronize(h->stream));
    return HVO_OK;
}

int hvo_lbd_sync(hvo_lbd* h) {
    HVO_CHECK_ARG(h, "null handle");
    HVO_CUDA(cudaSetDevice(h->device));
    HVO_CUDA(cudaStreamSynchronize(h->stream));
    return HVO_OK;
}
int hvo_lbd_timer_start(hvo_lbd* h) {
    HVO_CHECK_ARG(h, "null handle");
    HVO_CUDA(cudaSetDevice(h->device));
    HVO_CUDA(cudaEventRecord(h->tev[0], h->stream));
    return HVO_OK;
}
int hvo_lbd_timer_stop(hvo_lbd* h, float* ms_out) {
    HVO_CHECK_ARG(h && ms_out, "null argument");
    HVO_CUDA(cudaSetDevice(h->device));
    HVO_CUDA(cudaEventRecord(h->tev[1], h->stream));
    HVO_CUDA(cudaEventSynchronize(h->tev[1]));
    HVO_CUDA(cudaEventElapsedTime(ms_out, h->tev[0], h->tev[1]));
    return HVO_OK;
}

}  // extern "C"
