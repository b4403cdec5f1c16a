// Frame-level front-end for sm_100a: the extraction part of Frame::Frame(imGray, imDepth, ...) of the reference
// (src/Frame.cc:188-233), where three std::threads run ExtractORBNDepth (ORB + ComputeStereoFromRGBD, :874-884),
// ExtractLSD (LINEextractor::operator(), :895-903) and ComputePlanes (PlaneDetection + surface normals, :2104-2212)
// on the same frame.  Here the frame (gray + raw 16-bit depth) is uploaded once and the three pipelines run on three
// CUDA streams chained by events to a master stream, so a batch of frames is one call and one device-timed region.
#include <algorithm>
#include <cstdlib>
#include <new>

#include "hvo_common.cuh"

struct hvo_orb;
struct hvo_line;
struct hvo_plane;
struct hvo_normals;
namespace hvo {
cudaStream_t orb_stream(hvo_orb* h);
cudaStream_t line_stream(hvo_line* h);
cudaStream_t plane_stream(hvo_plane* h);
cudaStream_t normals_stream(hvo_normals* h);
int* orb_error_flag(hvo_orb* h);
const int* line_segment_counts(hvo_line* h);
int line_segment_cap(hvo_line* h);
const int32_t* plane_status(hvo_plane* h);
}  // namespace hvo
using namespace hvo;

enum { ST_ORB = 1, ST_LINE = 2, ST_PLANE = 4, ST_NORMALS = 8 };
static const int kMaxLanes = 8;

// Device-side capacity faults of the pipelines (ORB quadtree arena, LSD segment buffer, plane refinement queue) are folded,
// on the pipeline's own stream right after its kernels, into one sticky record per frame handle: fault[0] = lowest frame index
// (within its call) that overflowed, fault[1] = OR of the pipeline codes.  hvo_frame_sync / hvo_frame_timer_stop /
// hvo_frame_extract_batch read it back and return HVO_ERR_OVERFLOW ("reported, never silent", hvo_capi.h).
enum { FAULT_ORB = 1, FAULT_LINE = 2, FAULT_PLANE = 4 };
__global__ void k_frame_fold_status(const int* __restrict__ flags, int n, int stride_is_zero, int threshold, int base, int code, int* __restrict__ fault) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) {
        if (flags[stride_is_zero ? 0 : f] > threshold) {
            atomicMin(&fault[0], base + f);
            atomicOr(&fault[1], code);
        }
    }
}

// Compact host outputs: 4-bit plane labels and normal-only surface normals, packed on the device before the download.
__global__ void k_pack_membership4(const uint8_t* __restrict__ m8, uint8_t* __restrict__ m4, long long npairs8) {
    // 16 labels in, 8 bytes out per thread
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npairs8; i += (long long)gridDim.x * blockDim.x) {
        const uint4 v = reinterpret_cast<const uint4*>(m8)[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[2] = {0u, 0u};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // bytes b0 b1 b2 b3 (255 -> 15) -> two output bytes (b0 | b1 << 4), (b2 | b3 << 4)
            const uint32_t n = w[k] & 0x0f0f0f0fu;                      // 255 & 15 = 15 = none; labels are < 15
            const uint32_t packed = (n & 0xfu) | ((n >> 4) & 0xf0u) | ((n >> 8) & 0xf00u) | ((n >> 12) & 0xf000u);
            o[k >> 1] |= packed << (16 * (k & 1));
        }
        reinterpret_cast<uint2*>(m4)[i] = make_uint2(o[0], o[1]);
    }
}
__global__ void k_pack_normals3(const float* __restrict__ n8, float* __restrict__ n3, long long rows) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(n8)[2 * i];
        n3[3 * i] = v.x; n3[3 * i + 1] = v.y; n3[3 * i + 2] = v.z;
    }
}

// One lane = the staging of one chunk of `cap` frames in flight: input buffers, output buffers, an upload stream and one download
// stream per pipeline.  The pipeline handles (kernels' scratch and streams) belong to lane 0 and are shared by every lane: the
// kernels of consecutive chunks run one after the other anyway (see lane_launch), so a second copy of the 15 MB / frame of scratch
// would buy nothing, while sharing it lets a chunk be as large as the ordered kernels want (4096+ frames) inside 180 GB.
struct FrameLane {
    int cap = 0;
    bool owns = false;          // lane 0 owns the pipeline handles
    cudaStream_t down[4] = {nullptr, nullptr, nullptr, nullptr};  // host API: downloads of pipeline i (ORB, lines, planes, normals)
    cudaEvent_t fork_depth = nullptr;  // host API: the chunk's depth has arrived (the plane pipeline starts on it; gray follows)
    hvo_orb* orb = nullptr;
    hvo_line* line = nullptr;
    hvo_plane* plane = nullptr;
    hvo_normals* normals = nullptr;
    cudaStream_t up = nullptr;  // host API: uploads of this lane's chunks
    cudaEvent_t fork = nullptr, join[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t cdone[4] = {nullptr, nullptr, nullptr, nullptr};  // kernels of a pipeline done (recorded before its downloads)
    uint8_t* d_gray = nullptr;
    uint16_t* d_depth = nullptr;
    hvo_frame_outputs d_out;  // device staging of every output (host API)
};

struct hvo_frame {
    hvo_frame_params p;
    int device = 0, width = 0, height = 0, max_batch = 0, nlanes = 0;
    FrameLane lane[kMaxLanes];
    cudaStream_t stream = nullptr;  // master: fork/join of the device API, timing
    cudaEvent_t fork = nullptr, tev[2] = {nullptr, nullptr};
    int orb_cap = 0, max_lines = 0, normals_count = 0, last_launches = 0;
    int last_lane = -1;  // lane of the most recent chunk: the next chunk's kernels start after its kernels
    int next_lane = 0;   // host API: lane of the next chunk (round-robin across calls)
    int* d_fault = nullptr;   // {first faulty frame, pipeline codes}, sticky until read
    int* h_fault = nullptr;   // pinned copy
};

static const int kNoFault = 0x7fffffff;

// rows [off, off + n) of every output array
static hvo_frame_outputs outputs_at(const hvo_frame* h, const hvo_frame_outputs& o, size_t off) {
    const size_t c = (size_t)h->orb_cap, l = (size_t)h->max_lines, px = (size_t)h->width * h->height;
    hvo_frame_outputs r = o;
    if (o.kps) r.kps = o.kps + off * c;
    if (o.desc) r.desc = o.desc + off * c * 32;
    if (o.kp_counts) r.kp_counts = o.kp_counts + off;
    if (o.kp_depth) r.kp_depth = o.kp_depth + off * c;
    if (o.kp_uright) r.kp_uright = o.kp_uright + off * c;
    if (o.keylines) r.keylines = o.keylines + off * l;
    if (o.line_desc) r.line_desc = o.line_desc + off * l * 32;
    if (o.linevec3) r.linevec3 = o.linevec3 + off * l * 3;
    if (o.line_counts) r.line_counts = o.line_counts + off;
    if (o.n_planes) r.n_planes = o.n_planes + off;
    if (o.planes7) r.planes7 = o.planes7 + off * (size_t)h->p.max_planes * 7;
    if (o.membership) r.membership = o.membership + off * px;
    if (o.membership8) r.membership8 = o.membership8 + off * px;
    if (o.membership4) r.membership4 = o.membership4 + off * (px / 2);
    if (o.normals8) r.normals8 = o.normals8 + off * (size_t)h->normals_count * 8;
    if (o.normals3) r.normals3 = o.normals3 + off * (size_t)h->normals_count * 3;
    return r;
}

// The pipelines of one lane on n frames: wait for `start`, run, optionally copy each stage's results back on its own
// stream, record the lane's join events.  Launch order = placement order: the ordered (latency-bound) pipelines first.
// `after`: lane whose kernels must have finished first (host API: chunks compute one after the other, so the ordered kernels
// of two chunks never slow each other down, while the copies of the neighbouring chunks run beside the kernels).
//
// Two schedules (measured on B200, whole front-end, device-resident): SIDE BY SIDE (every pipeline starts at once on its own stream)
// is what a small batch wants - a single frame takes max(planes, lines) instead of their sum - but each of the ordered kernels fills
// the register file with one warp / CTA per frame by itself, and from about 2000 frames on running them beside each other only makes
// them slower (4736 frames: 16.2 K frames/s side by side, 18.2 K one after the other; ORB / normals beside the ordered kernels:
// 16.5-17.3 K).  So a chunk of >= kSerialFrames frames runs its pipelines ONE AFTER THE OTHER (planes, lines, ORB, normals), each at
// its saturating batch, and the download of a pipeline's results runs beside the kernels of the next one.
// HVO_FRAME_SERIAL=0 / 1 forces one schedule (tuning aid).  Also measured for 1024-frame chunks and rejected: planes || lines first and
// ORB || normals afterwards (87 instead of 65-79 ms), a third priority level that puts planes above lines (no change).
static const int kSerialFrames = 2048;
static int lane_launch(hvo_frame* h, FrameLane& L, cudaEvent_t start, cudaEvent_t start_depth, const uint8_t* d_gray, const uint16_t* d_depth, int n,
                       int base, const hvo_frame_outputs& o, const hvo_frame_outputs* host, int* launches, FrameLane* after = nullptr) {
    const size_t N = (size_t)n, px = (size_t)h->width * h->height;
    static const int serial_env = [] { const char* e = getenv("HVO_FRAME_SERIAL"); return e ? atoi(e) : -1; }();
    const bool serial = serial_env >= 0 ? serial_env == 1 : n >= kSerialFrames;
    // side by side on one lane: every pipeline stream is already in order behind its own previous chunk
    if (after == &L && !serial) after = nullptr;
    cudaEvent_t prev_done = nullptr;   // serial schedule: the pipeline launched before this one
    auto wait_start = [&](cudaStream_t s, bool depth_only) -> int {
        HVO_CUDA(cudaStreamWaitEvent(s, (depth_only && start_depth) ? start_depth : start, 0));
        if (serial && prev_done) {
            HVO_CUDA(cudaStreamWaitEvent(s, prev_done, 0));
        } else if (after) {   // (serial: only the first pipeline waits for the previous chunk, the others follow it)
            for (int i = 0; i < 4; ++i)
                if (h->p.stages & (i == 0 ? ST_ORB : i == 1 ? ST_LINE : i == 2 ? ST_PLANE : ST_NORMALS)) HVO_CUDA(cudaStreamWaitEvent(s, after->cdone[i], 0));
        }
        return HVO_OK;
    };
    // ---- phase 1: the kernels of every pipeline are queued before any copy.  (A device-to-host copy into pageable memory
    // blocks the calling thread until the stream gets there; queued behind the plane kernels it would keep the other
    // pipelines from even starting.) ----
    if (h->p.stages & ST_PLANE) {
        cudaStream_t s = plane_stream(L.plane);
        if (int ws = wait_start(s, true)) return ws;
        int st = o.membership8 ? hvo_plane_detect_batch_device_u8(L.plane, d_depth, n, o.n_planes, o.planes7, h->p.max_planes, o.membership, o.membership8)
                               : hvo_plane_detect_batch_device(L.plane, d_depth, n, o.n_planes, o.planes7, h->p.max_planes, o.membership);
        if (st != HVO_OK) return st;
        *launches += hvo_plane_last_launches(L.plane) + 1;
        if (o.membership4 && o.membership8) {
            const long long n16 = (long long)(N * px / 16);
            k_pack_membership4<<<(int)std::min<long long>((n16 + 255) / 256, 148 * 16), 256, 0, s>>>(o.membership8, o.membership4, n16);
            ++*launches;
        }
        k_frame_fold_status<<<(n + 255) / 256, 256, 0, s>>>(plane_status(L.plane), n, 0, 0, base, FAULT_PLANE, h->d_fault);
        HVO_CUDA(cudaEventRecord(L.cdone[2], s));
        prev_done = L.cdone[2];
    }
    if (h->p.stages & ST_LINE) {
        cudaStream_t s = line_stream(L.line);
        if (int ws = wait_start(s, false)) return ws;
        int st = hvo_line_extract_batch_device(L.line, d_gray, n, o.keylines, o.line_desc, o.linevec3, o.line_counts);
        if (st != HVO_OK) return st;
        *launches += hvo_line_last_launches(L.line) + 1;
        k_frame_fold_status<<<(n + 255) / 256, 256, 0, s>>>(line_segment_counts(L.line), n, 0, line_segment_cap(L.line), base, FAULT_LINE, h->d_fault);
        HVO_CUDA(cudaEventRecord(L.cdone[1], s));
        prev_done = L.cdone[1];
    }
    if (h->p.stages & ST_ORB) {
        cudaStream_t s = orb_stream(L.orb);
        if (int ws = wait_start(s, false)) return ws;
        hvo_rgbd_params rg{h->p.depth_factor, h->p.bf, h->p.distorted};
        int st = hvo_orb_extract_batch_device(L.orb, d_gray, n, o.kps, o.desc, o.kp_counts, d_depth, &rg, o.kp_depth, o.kp_uright);
        if (st != HVO_OK) return st;
        *launches += hvo_orb_last_launches(L.orb) + 1;
        k_frame_fold_status<<<1, 32, 0, s>>>(orb_error_flag(L.orb), 1, 1, 0, base, FAULT_ORB, h->d_fault);   // one flag per chunk
        HVO_CUDA(cudaMemsetAsync(orb_error_flag(L.orb), 0, sizeof(int), s));
        HVO_CUDA(cudaEventRecord(L.cdone[0], s));
        prev_done = L.cdone[0];
    }
    if (h->p.stages & ST_NORMALS) {
        cudaStream_t s = normals_stream(L.normals);
        if (int ws = wait_start(s, false)) return ws;
        int st = hvo_normals_compute_batch_device(L.normals, d_depth, n, o.normals8);
        if (st != HVO_OK) return st;
        *launches += 5;
        if (o.normals3) {
            const long long rows = (long long)(N * (size_t)h->normals_count);
            k_pack_normals3<<<(int)std::min<long long>((rows + 255) / 256, 148 * 16), 256, 0, s>>>(o.normals8, o.normals3, rows);
            ++*launches;
        }
        HVO_CUDA(cudaEventRecord(L.cdone[3], s));
        prev_done = L.cdone[3];
    }
    // ---- phase 2: results back to the host on the lane's download streams (one per pipeline, each behind its pipeline's kernels:
    // the shared pipeline streams stay free for the next chunk), then the join events ----
    auto down_stream = [&](int i, cudaStream_t pipe, cudaStream_t* out) -> int {
        *out = pipe;
        if (host) { HVO_CUDA(cudaStreamWaitEvent(L.down[i], L.cdone[i], 0)); *out = L.down[i]; }
        return HVO_OK;
    };
    if (h->p.stages & ST_PLANE) {
        cudaStream_t s;
        if (int ds = down_stream(2, plane_stream(L.plane), &s)) return ds;
        if (host) {
            HVO_CUDA(cudaMemcpyAsync(host->n_planes, o.n_planes, N * 4, cudaMemcpyDeviceToHost, s));
            HVO_CUDA(cudaMemcpyAsync(host->planes7, o.planes7, N * (size_t)h->p.max_planes * 56, cudaMemcpyDeviceToHost, s));
            if (host->membership) HVO_CUDA(cudaMemcpyAsync(host->membership, o.membership, N * px * 4, cudaMemcpyDeviceToHost, s));
            if (host->membership8) HVO_CUDA(cudaMemcpyAsync(host->membership8, o.membership8, N * px, cudaMemcpyDeviceToHost, s));
            if (host->membership4) HVO_CUDA(cudaMemcpyAsync(host->membership4, o.membership4, N * (px / 2), cudaMemcpyDeviceToHost, s));
        }
        timeline_mark(s, "end_planes");
        HVO_CUDA(cudaEventRecord(L.join[2], s));
    }
    if (h->p.stages & ST_LINE) {
        cudaStream_t s;
        if (int ds = down_stream(1, line_stream(L.line), &s)) return ds;
        if (host) {
            const size_t c = (size_t)h->max_lines;
            HVO_CUDA(cudaMemcpyAsync(host->line_counts, o.line_counts, N * 4, cudaMemcpyDeviceToHost, s));
            HVO_CUDA(cudaMemcpyAsync(host->keylines, o.keylines, N * c * sizeof(hvo_keyline), cudaMemcpyDeviceToHost, s));
            HVO_CUDA(cudaMemcpyAsync(host->line_desc, o.line_desc, N * c * 32, cudaMemcpyDeviceToHost, s));
            if (host->linevec3) HVO_CUDA(cudaMemcpyAsync(host->linevec3, o.linevec3, N * c * 24, cudaMemcpyDeviceToHost, s));
        }
        timeline_mark(s, "end_lines");
        HVO_CUDA(cudaEventRecord(L.join[1], s));
    }
    if (h->p.stages & ST_ORB) {
        cudaStream_t s;
        if (int ds = down_stream(0, orb_stream(L.orb), &s)) return ds;
        if (host) {
            const size_t c = (size_t)h->orb_cap;
            HVO_CUDA(cudaMemcpyAsync(host->kp_counts, o.kp_counts, N * 4, cudaMemcpyDeviceToHost, s));
            HVO_CUDA(cudaMemcpyAsync(host->kps, o.kps, N * c * sizeof(hvo_keypoint), cudaMemcpyDeviceToHost, s));
            HVO_CUDA(cudaMemcpyAsync(host->desc, o.desc, N * c * 32, cudaMemcpyDeviceToHost, s));
            if (host->kp_depth) HVO_CUDA(cudaMemcpyAsync(host->kp_depth, o.kp_depth, N * c * 4, cudaMemcpyDeviceToHost, s));
            if (host->kp_uright) HVO_CUDA(cudaMemcpyAsync(host->kp_uright, o.kp_uright, N * c * 4, cudaMemcpyDeviceToHost, s));
        }
        timeline_mark(s, "end_orb");
        HVO_CUDA(cudaEventRecord(L.join[0], s));
    }
    if (h->p.stages & ST_NORMALS) {
        cudaStream_t s;
        if (int ds = down_stream(3, normals_stream(L.normals), &s)) return ds;
        if (host && host->normals8) HVO_CUDA(cudaMemcpyAsync(host->normals8, o.normals8, N * (size_t)h->normals_count * 32, cudaMemcpyDeviceToHost, s));
        if (host && host->normals3) HVO_CUDA(cudaMemcpyAsync(host->normals3, o.normals3, N * (size_t)h->normals_count * 12, cudaMemcpyDeviceToHost, s));
        timeline_mark(s, "end_normals");
        HVO_CUDA(cudaEventRecord(L.join[3], s));
    }
    return HVO_OK;
}

static int wait_joins(hvo_frame* h, cudaStream_t s, FrameLane& L) {
    for (int i = 0; i < 4; ++i)
        if (h->p.stages & (i == 0 ? ST_ORB : i == 1 ? ST_LINE : i == 2 ? ST_PLANE : ST_NORMALS)) HVO_CUDA(cudaStreamWaitEvent(s, L.join[i], 0));
    return HVO_OK;
}

// After the master stream has been synchronised: read the sticky fault record, reset it, report.
static int frame_check_fault(hvo_frame* h) {
    HVO_CUDA(cudaMemcpyAsync(h->h_fault, h->d_fault, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    HVO_CUDA(cudaStreamSynchronize(h->stream));
    if (h->h_fault[1] == 0) return HVO_OK;
    const int frame = h->h_fault[0], code = h->h_fault[1];
    h->h_fault[0] = kNoFault; h->h_fault[1] = 0;
    HVO_CUDA(cudaMemcpyAsync(h->d_fault, h->h_fault, 2 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    HVO_CUDA(cudaStreamSynchronize(h->stream));
    set_error("device-side capacity overflow in frame %d of its call:%s%s%s (the outputs of that frame are truncated)", frame,
              (code & FAULT_ORB) ? " ORB quadtree arena" : "", (code & FAULT_LINE) ? " LSD segment buffer" : "",
              (code & FAULT_PLANE) ? " plane refinement queue" : "");
    return HVO_ERR_OVERFLOW;
}

extern "C" {

int hvo_membership4_expand(const uint8_t* labels4, int npixels, int32_t* labels) {
    HVO_CHECK_ARG(labels4 && labels && npixels >= 0 && (npixels & 1) == 0, "bad argument");
    for (int i = 0; i < npixels / 2; ++i) {
        const int a = labels4[i] & 15, b = labels4[i] >> 4;
        labels[2 * i] = a == 15 ? -1 : a;
        labels[2 * i + 1] = b == 15 ? -1 : b;
    }
    return HVO_OK;
}

int hvo_normals3_expand(const float* normals3, const uint16_t* depth16, int width, int height, float fx, float fy, float cx, float cy,
                        float depth_factor, float* normals8) {
    HVO_CHECK_ARG(normals3 && depth16 && normals8 && width > 0 && height > 0, "bad argument");
    const int cw = (width + 2) / 3, ch = (height + 2) / 3, ow = cw / 2, oh = ch / 2;
    for (int i = 0; i < ow * oh; ++i) {
        const int m = 2 * (i / ow) + 1, n = 2 * (i % ow) + 1;
        // the subsampled cloud of Frame::ComputePlanes (Frame.cc:2158-2171), float arithmetic as on the device (no FMA)
        const float z = (float)depth16[(size_t)(3 * m) * width + 3 * n] * depth_factor;
        float* o = normals8 + (size_t)i * 8;
        o[0] = normals3[3 * i]; o[1] = normals3[3 * i + 1]; o[2] = normals3[3 * i + 2];
        o[3] = ((float)(3 * n) - cx) * z / fx;
        o[4] = ((float)(3 * m) - cy) * z / fy;
        o[5] = z;
        o[6] = (float)(n * 3); o[7] = (float)(m * 3);
    }
    return HVO_OK;
}

int hvo_frame_create(const hvo_frame_params* p, int width, int height, int max_batch, int device, hvo_frame** out) {
    HVO_CHECK_ARG(out, "null out");
    *out = nullptr;
    HVO_CHECK_ARG(p, "null params");
    HVO_CHECK_ARG((p->stages & 15) != 0, "no stage selected");
    HVO_CHECK_ARG(max_batch >= 1 && p->max_planes >= 1, "max_batch / max_planes < 1");
    HVO_CHECK_ARG(p->lanes >= 0 && p->lanes <= kMaxLanes, "lanes out of range (0..8)");
    int ndev = 0;
    HVO_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev < 1) { set_error("no CUDA device: libhvofront has no CPU fallback"); return HVO_ERR_CUDA; }
    HVO_CHECK_ARG(device >= 0 && device < ndev, "device index out of range");
    hvo_frame* h = new (std::nothrow) hvo_frame();
    if (!h) { set_error("out of host memory"); return HVO_ERR_ARG; }
    h->p = *p; h->device = device; h->width = width; h->height = height; h->max_batch = max_batch;
    // default: two lanes once a chunk still holds >= 1024 frames (measured on B200: the ordered kernels, one warp / CTA per
    // frame, need about that many frames per launch; 4 x 256 is 25 % slower than 1 x 1024)
    int nl = p->lanes > 0 ? p->lanes : (max_batch >= 2048 ? 2 : 1);
    if (const char* e = getenv("HVO_FRAME_LANES")) nl = std::min(kMaxLanes, std::max(1, atoi(e)));  // tuning aid
    nl = std::min(nl, max_batch);
    h->nlanes = nl;
    const int cap = (max_batch + nl - 1) / nl;
    int st = HVO_OK;
    // lines and planes are chains of long latency-bound kernels (one warp / CTA per frame): they get the high-priority
    // streams so they are resident from the start; ORB and normals (streaming kernels with large grids) fill in around them
    int prio_lo = 0, prio_hi = 0;
    HVO_CUDA(cudaSetDevice(device));
    HVO_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    const char* env_prio = getenv("HVO_FRAME_PRIORITIES");  // tuning aid: "0" = all streams at the default priority
    if (env_prio && env_prio[0] == '0') prio_lo = prio_hi = 0;
    for (int li = 0; li < nl && st == HVO_OK; ++li) {
        FrameLane& L = h->lane[li];
        L.cap = cap;
        L.d_out = hvo_frame_outputs{};
        if (li > 0) {   // the pipelines (scratch + streams) are lane 0's
            L.orb = h->lane[0].orb; L.line = h->lane[0].line; L.plane = h->lane[0].plane; L.normals = h->lane[0].normals;
            continue;
        }
        L.owns = true;
        set_next_stream_priority(prio_lo);
        if (st == HVO_OK && (p->stages & ST_ORB)) st = hvo_orb_create(&p->orb, width, height, cap, device, &L.orb);
        set_next_stream_priority(prio_hi);
        if (st == HVO_OK && (p->stages & ST_LINE)) st = hvo_line_create(&p->line, width, height, cap, device, &L.line);
        if (st == HVO_OK && L.line) st = hvo_line_set_culling(L.line, p->line_cull);
        if (st == HVO_OK && (p->stages & ST_PLANE)) {
            hvo_plane_params pp{p->fx, p->fy, p->cx, p->cy, p->depth_factor};
            st = hvo_plane_create(&pp, width, height, cap, device, &L.plane);
        }
        set_next_stream_priority(prio_lo);
        if (st == HVO_OK && (p->stages & ST_NORMALS)) {
            hvo_normals_params np{p->fx, p->fy, p->cx, p->cy, p->depth_factor, 0.05f, 10.0f};  // Frame.cc:2179-2180
            st = hvo_normals_create(&np, width, height, cap, device, &L.normals);
        }
    }
    set_next_stream_priority(0);
    if (st != HVO_OK) { hvo_frame_destroy(h); return st; }
    const FrameLane& L0 = h->lane[0];
    h->orb_cap = L0.orb ? hvo_orb_capacity(L0.orb) : 0;
    h->max_lines = L0.line ? hvo_line_max_lines(L0.line) : 0;
    h->normals_count = L0.normals ? hvo_normals_count(L0.normals) : 0;
    do {
#define HVO_TRY(call) if ((call) != cudaSuccess) { set_error("%s: %s", #call, cudaGetErrorString(cudaGetLastError())); st = HVO_ERR_CUDA; break; }
        HVO_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        HVO_TRY(cudaEventCreateWithFlags(&h->fork, cudaEventDisableTiming));
        for (auto& e : h->tev) HVO_TRY(cudaEventCreate(&e));
        if (st != HVO_OK) break;
        HVO_TRY(cudaMalloc(&h->d_fault, 2 * sizeof(int)));
        HVO_TRY(cudaHostAlloc(&h->h_fault, 2 * sizeof(int), cudaHostAllocDefault));
        h->h_fault[0] = kNoFault; h->h_fault[1] = 0;
        HVO_TRY(cudaMemcpy(h->d_fault, h->h_fault, 2 * sizeof(int), cudaMemcpyHostToDevice));
        const size_t B = (size_t)cap, px = (size_t)width * height;
        for (int li = 0; li < nl && st == HVO_OK; ++li) {
            FrameLane& L = h->lane[li];
            HVO_TRY(cudaStreamCreateWithFlags(&L.up, cudaStreamNonBlocking));
            HVO_TRY(cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming));
            HVO_TRY(cudaEventCreateWithFlags(&L.fork_depth, cudaEventDisableTiming));
            for (auto& d : L.down) HVO_TRY(cudaStreamCreateWithFlags(&d, cudaStreamNonBlocking));
            if (st != HVO_OK) break;
            for (auto& e : L.join) HVO_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            if (st != HVO_OK) break;
            for (auto& e : L.cdone) HVO_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            if (st != HVO_OK) break;
            HVO_TRY(cudaMalloc(&L.d_gray, B * px));
            HVO_TRY(cudaMalloc(&L.d_depth, B * px * 2));
            hvo_frame_outputs& o = L.d_out;
            if (L.orb) {
                const size_t c = (size_t)h->orb_cap;
                HVO_TRY(cudaMalloc(&o.kps, B * c * sizeof(hvo_keypoint)));
                HVO_TRY(cudaMalloc(&o.desc, B * c * 32));
                HVO_TRY(cudaMalloc(&o.kp_counts, B * 4));
                HVO_TRY(cudaMalloc(&o.kp_depth, B * c * 4));
                HVO_TRY(cudaMalloc(&o.kp_uright, B * c * 4));
            }
            if (L.line) {
                const size_t c = (size_t)h->max_lines;
                HVO_TRY(cudaMalloc(&o.keylines, B * c * sizeof(hvo_keyline)));
                HVO_TRY(cudaMalloc(&o.line_desc, B * c * 32));
                HVO_TRY(cudaMalloc(&o.linevec3, B * c * 24));
                HVO_TRY(cudaMalloc(&o.line_counts, B * 4));
            }
            if (L.plane) {
                HVO_TRY(cudaMalloc(&o.n_planes, B * 4));
                HVO_TRY(cudaMalloc(&o.planes7, B * (size_t)p->max_planes * 56));
                HVO_TRY(cudaMalloc(&o.membership, B * px * 4));
                HVO_TRY(cudaMalloc(&o.membership8, B * px));
                HVO_TRY(cudaMalloc(&o.membership4, B * (px / 2) + 16));
            }
            if (L.normals) {
                HVO_TRY(cudaMalloc(&o.normals8, B * (size_t)h->normals_count * 32));
                HVO_TRY(cudaMalloc(&o.normals3, B * (size_t)h->normals_count * 12));
            }
        }
#undef HVO_TRY
    } while (0);
    if (st != HVO_OK) { hvo_frame_destroy(h); return st; }
    *out = h;
    return HVO_OK;
}

void hvo_frame_destroy(hvo_frame* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (int li = 0; li < kMaxLanes; ++li) {
        FrameLane& L = h->lane[li];
        if (L.owns) {
            if (L.orb) hvo_orb_destroy(L.orb);
            if (L.line) hvo_line_destroy(L.line);
            if (L.plane) hvo_plane_destroy(L.plane);
            if (L.normals) hvo_normals_destroy(L.normals);
        }
        for (auto& d : L.down) if (d) cudaStreamDestroy(d);
        if (L.fork_depth) cudaEventDestroy(L.fork_depth);
        hvo_frame_outputs& o = L.d_out;
        void* bufs[] = {L.d_gray, L.d_depth, o.kps, o.desc, o.kp_counts, o.kp_depth, o.kp_uright, o.keylines, o.line_desc, o.linevec3,
                        o.line_counts, o.n_planes, o.planes7, o.membership, o.membership8, o.normals8, o.membership4, o.normals3};
        for (void* b : bufs) if (b) cudaFree(b);
        if (L.fork) cudaEventDestroy(L.fork);
        for (auto& e : L.join) if (e) cudaEventDestroy(e);
        for (auto& e : L.cdone) if (e) cudaEventDestroy(e);
        if (L.up) cudaStreamDestroy(L.up);
    }
    if (h->d_fault) cudaFree(h->d_fault);
    if (h->h_fault) cudaFreeHost(h->h_fault);
    if (h->fork) cudaEventDestroy(h->fork);
    for (auto& e : h->tev) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int hvo_frame_capacities(const hvo_frame* h, int* orb_capacity, int* max_lines, int* normals_count) {
    HVO_CHECK_ARG(h, "null handle");
    if (orb_capacity) *orb_capacity = h->orb_cap;
    if (max_lines) *max_lines = h->max_lines;
    if (normals_count) *normals_count = h->normals_count;
    return HVO_OK;
}

int hvo_frame_lanes(const hvo_frame* h, int* lanes, int* chunk) {
    HVO_CHECK_ARG(h, "null handle");
    if (lanes) *lanes = h->nlanes;
    if (chunk) *chunk = h->lane[0].cap;
    return HVO_OK;
}

static int frame_check_outputs(const hvo_frame* h, const hvo_frame_outputs* o, bool device) {
    HVO_CHECK_ARG(o, "null outputs");
    if (h->p.stages & ST_ORB) HVO_CHECK_ARG(o->kps && o->desc && o->kp_counts, "ORB outputs missing");
    if (h->p.stages & ST_LINE) HVO_CHECK_ARG(o->keylines && o->line_desc && o->line_counts, "line outputs missing");
    if (h->p.stages & ST_PLANE) HVO_CHECK_ARG(o->n_planes && o->planes7 && (o->membership || (!device && (o->membership8 || o->membership4))), "plane outputs missing");
    if (h->p.stages & ST_NORMALS) HVO_CHECK_ARG(o->normals8 || (!device && o->normals3), "normals output missing");
    if (o->membership4) {
        HVO_CHECK_ARG(h->p.max_planes <= 15, "membership4 needs max_planes <= 15");
        HVO_CHECK_ARG(((size_t)h->width * h->height) % 16 == 0, "membership4 needs width * height divisible by 16");
        if (device) HVO_CHECK_ARG(o->membership8, "membership4 on the device needs membership8 as well");
    }
    return HVO_OK;
}

int hvo_frame_extract_batch_device(hvo_frame* h, const uint8_t* d_gray, const uint16_t* d_depth16, int nframes, const hvo_frame_outputs* d_out) {
    HVO_CHECK_ARG(h && d_gray && d_depth16, "null argument");
    HVO_CHECK_ARG(nframes >= 1 && nframes <= h->max_batch, "nframes out of range for this handle");
    int st = frame_check_outputs(h, d_out, true);
    if (st != HVO_OK) return st;
    if (h->p.stages & ST_ORB) HVO_CHECK_ARG(d_out->kp_depth && d_out->kp_uright, "kp_depth / kp_uright missing");
    if (h->p.stages & ST_LINE) HVO_CHECK_ARG(d_out->linevec3, "linevec3 missing");
    HVO_CUDA(cudaSetDevice(h->device));
    const size_t px = (size_t)h->width * h->height;
    HVO_CUDA(cudaEventRecord(h->fork, h->stream));
    int launches = 0, used = 0;
    // the batch is spread evenly over the lanes; the chunks compute one after the other (measured: two chunks whose ordered
    // kernels overlap are slower than the same chunks back to back)
    // as few chunks as the lane capacity allows, of equal size (the ordered kernels are latency-bound: one warp / CTA per frame, so
    // a chunk of 1024 is faster than two of 512)
    const int nchunks = (nframes + h->lane[0].cap - 1) / h->lane[0].cap;
    const int per = (nframes + nchunks - 1) / nchunks;
    for (int off = 0; off < nframes; off += per, ++used) {
        FrameLane& L = h->lane[used];
        const int n = std::min(per, nframes - off);
        st = lane_launch(h, L, h->fork, nullptr, d_gray + (size_t)off * px, d_depth16 + (size_t)off * px, n, off, outputs_at(h, *d_out, (size_t)off), nullptr, &launches,
                         h->last_lane >= 0 ? &h->lane[h->last_lane] : nullptr);
        if (st != HVO_OK) return st;
        h->last_lane = used;
    }
    for (int li = 0; li < used; ++li) {
        st = wait_joins(h, h->stream, h->lane[li]);
        if (st != HVO_OK) return st;
    }
    h->last_launches = launches;
    return HVO_OK;
}

int hvo_frame_extract_batch_async(hvo_frame* h, const uint8_t* gray, const uint16_t* depth16, int nframes, const hvo_frame_outputs* out) {
    HVO_CHECK_ARG(h && gray && depth16, "null argument");
    HVO_CHECK_ARG(nframes >= 1, "nframes < 1");
    int st = frame_check_outputs(h, out, false);
    if (st != HVO_OK) return st;
    HVO_CUDA(cudaSetDevice(h->device));
    const size_t px = (size_t)h->width * h->height;
    // chunks: as few as the lane capacity allows, of equal size.  Chunks go round-robin through the lanes ACROSS calls
    // (next_lane), so that the upload of a one-chunk call overlaps the kernels of the previous call on the other lane.
    const int cap = h->lane[0].cap;
    const int nchunks = (nframes + cap - 1) / cap;
    const int per = (nframes + nchunks - 1) / nchunks;
    int launches = 0, k = 0;
    bool lane_used[kMaxLanes] = {false};
    for (int off = 0; off < nframes; off += per, ++k) {
        const int li = h->next_lane;
        h->next_lane = (h->next_lane + 1) % h->nlanes;
        lane_used[li] = true;
        FrameLane& L = h->lane[li];
        const int n = std::min(per, nframes - off);
        st = wait_joins(h, L.up, L);  // the lane's previous chunk (of this or an earlier call) must be done with the staging buffers
        if (st != HVO_OK) return st;
        // depth first: the plane pipeline (first in the serial schedule) needs nothing else, the gray upload runs beside its kernels
        HVO_CUDA(cudaMemcpyAsync(L.d_depth, depth16 + (size_t)off * px, (size_t)n * px * 2, cudaMemcpyHostToDevice, L.up));
        HVO_CUDA(cudaEventRecord(L.fork_depth, L.up));
        HVO_CUDA(cudaMemcpyAsync(L.d_gray, gray + (size_t)off * px, (size_t)n * px, cudaMemcpyHostToDevice, L.up));
        HVO_CUDA(cudaEventRecord(L.fork, L.up));
        hvo_frame_outputs d = L.d_out;
        if (!out->membership8 && !out->membership4) d.membership8 = nullptr;
        if (!out->membership4) d.membership4 = nullptr;
        if (!out->normals3) d.normals3 = nullptr;
        const hvo_frame_outputs hostk = outputs_at(h, *out, (size_t)off);
        st = lane_launch(h, L, L.fork, L.fork_depth, L.d_gray, L.d_depth, n, off, d, &hostk, &launches, h->last_lane >= 0 ? &h->lane[h->last_lane] : nullptr);
        if (st != HVO_OK) return st;
        h->last_lane = li;
    }
    for (int li = 0; li < h->nlanes; ++li) {
        if (!lane_used[li]) continue;
        st = wait_joins(h, h->stream, h->lane[li]);
        if (st != HVO_OK) return st;
    }
    h->last_launches = launches;
    return HVO_OK;
}

int hvo_frame_extract_batch(hvo_frame* h, const uint8_t* gray, const uint16_t* depth16, int nframes, const hvo_frame_outputs* out) {
    const int st = hvo_frame_extract_batch_async(h, gray, depth16, nframes, out);
    if (st != HVO_OK) return st;
    HVO_CUDA(cudaStreamSynchronize(h->stream));
    return frame_check_fault(h);
}

int hvo_frame_last_launches(const hvo_frame* h) { return h ? h->last_launches : 0; }
int hvo_frame_sync(hvo_frame* h) {
    HVO_CHECK_ARG(h, "null handle");
    HVO_CUDA(cudaSetDevice(h->device));
    HVO_CUDA(cudaStreamSynchronize(h->stream));
    return frame_check_fault(h);
}
int hvo_frame_timer_start(hvo_frame* h) {
    HVO_CHECK_ARG(h, "null handle");
    HVO_CUDA(cudaSetDevice(h->device));
    HVO_CUDA(cudaEventRecord(h->tev[0], h->stream));
    return HVO_OK;
}
int hvo_frame_timer_stop(hvo_frame* h, float* ms_out) {
    HVO_CHECK_ARG(h && ms_out, "null argument");
    HVO_CUDA(cudaSetDevice(h->device));
    HVO_CUDA(cudaEventRecord(h->tev[1], h->stream));
    HVO_CUDA(cudaEventSynchronize(h->tev[1]));
    HVO_CUDA(cudaEventElapsedTime(ms_out, h->tev[0], h->tev[1]));
    return frame_check_fault(h);
}

}  // extern "C"
