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

// One lane = an independent copy of the three pipelines (own handles, streams, scratch) for chunks of `cap` frames.
struct FrameLane {
    int cap = 0;
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
    if (o.normals8) r.normals8 = o.normals8 + off * (size_t)h->normals_count * 8;
    return r;
}

// The pipelines of one lane on n frames: wait for `start`, run, optionally copy each stage's results back on its own
// stream, record the lane's join events.  Launch order = placement order: the ordered (latency-bound) pipelines first.
// `after`: lane whose kernels must have finished first (host API: chunks compute one after the other, so the ordered kernels
// of two chunks never slow each other down, while the copies of the neighbouring chunks run beside the kernels).
static int lane_launch(hvo_frame* h, FrameLane& L, cudaEvent_t start, const uint8_t* d_gray, const uint16_t* d_depth, int n, int base,
                       const hvo_frame_outputs& o, const hvo_frame_outputs* host, int* launches, FrameLane* after = nullptr) {
    const size_t N = (size_t)n, px = (size_t)h->width * h->height;
    if (after == &L) after = nullptr;  // one lane: every pipeline stream is already in order behind its own previous chunk
    auto wait_start = [&](cudaStream_t s) -> int {
        HVO_CUDA(cudaStreamWaitEvent(s, start, 0));
        if (after)
            for (int i = 0; i < 4; ++i)
                if (h->p.stages & (i == 0 ? ST_ORB : i == 1 ? ST_LINE : i == 2 ? ST_PLANE : ST_NORMALS)) HVO_CUDA(cudaStreamWaitEvent(s, after->cdone[i], 0));
        return HVO_OK;
    };
    // ---- phase 1: the kernels of every pipeline are queued before any copy.  (A device-to-host copy into pageable memory
    // blocks the calling thread until the stream gets there; queued behind the plane kernels it would keep the other
    // pipelines from even starting.) ----
    if (h->p.stages & ST_PLANE) {
        cudaStream_t s = plane_stream(L.plane);
        if (int ws = wait_start(s)) return ws;
        int st = o.membership8 ? hvo_plane_detect_batch_device_u8(L.plane, d_depth, n, o.n_planes, o.planes7, h->p.max_planes, o.membership, o.membership8)
                               : hvo_plane_detect_batch_device(L.plane, d_depth, n, o.n_planes, o.planes7, h->p.max_planes, o.membership);
        if (st != HVO_OK) return st;
        *launches += hvo_plane_last_launches(L.plane) + 1;
        k_frame_fold_status<<<(n + 255) / 256, 256, 0, s>>>(plane_status(L.plane), n, 0, 0, base, FAULT_PLANE, h->d_fault);
        HVO_CUDA(cudaEventRecord(L.cdone[2], s));
    }
    if (h->p.stages & ST_LINE) {
        cudaStream_t s = line_stream(L.line);
        if (int ws = wait_start(s)) return ws;
        int st = hvo_line_extract_batch_device(L.line, d_gray, n, o.keylines, o.line_desc, o.linevec3, o.line_counts);
        if (st != HVO_OK) return st;
        *launches += hvo_line_last_launches(L.line) + 1;
        k_frame_fold_status<<<(n + 255) / 256, 256, 0, s>>>(line_segment_counts(L.line), n, 0, line_segment_cap(L.line), base, FAULT_LINE, h->d_fault);
        HVO_CUDA(cudaEventRecord(L.cdone[1], s));
    }
    if (h->p.stages & ST_ORB) {
        cudaStream_t s = orb_stream(L.orb);
        if (int ws = wait_start(s)) return ws;
        hvo_rgbd_params rg{h->p.depth_factor, h->p.bf, h->p.distorted};
        int st = hvo_orb_extract_batch_device(L.orb, d_gray, n, o.kps, o.desc, o.kp_counts, d_depth, &rg, o.kp_depth, o.kp_uright);
        if (st != HVO_OK) return st;
        *launches += hvo_orb_last_launches(L.orb) + 1;
        k_frame_fold_status<<<1, 32, 0, s>>>(orb_error_flag(L.orb), 1, 1, 0, base, FAULT_ORB, h->d_fault);   // one flag per chunk
        HVO_CUDA(cudaMemsetAsync(orb_error_flag(L.orb), 0, sizeof(int), s));
        HVO_CUDA(cudaEventRecord(L.cdone[0], s));
    }
    if (h->p.stages & ST_NORMALS) {
        cudaStream_t s = normals_stream(L.normals);
        if (int ws = wait_start(s)) return ws;
        int st = hvo_normals_compute_batch_device(L.normals, d_depth, n, o.normals8);
        if (st != HVO_OK) return st;
        *launches += 5;
        HVO_CUDA(cudaEventRecord(L.cdone[3], s));
    }
    // ---- phase 2: results back to the host on each pipeline's own stream (shortest pipelines first), then the join events ----
    if (h->p.stages & ST_NORMALS) {
        cudaStream_t s = normals_stream(L.normals);
        if (host) HVO_CUDA(cudaMemcpyAsync(host->normals8, o.normals8, N * (size_t)h->normals_count * 32, cudaMemcpyDeviceToHost, s));
        timeline_mark(s, "end_normals");
        HVO_CUDA(cudaEventRecord(L.join[3], s));
    }
    if (h->p.stages & ST_ORB) {
        cudaStream_t s = orb_stream(L.orb);
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
    if (h->p.stages & ST_LINE) {
        cudaStream_t s = line_stream(L.line);
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
    if (h->p.stages & ST_PLANE) {
        cudaStream_t s = plane_stream(L.plane);
        if (host) {
            HVO_CUDA(cudaMemcpyAsync(host->n_planes, o.n_planes, N * 4, cudaMemcpyDeviceToHost, s));
            HVO_CUDA(cudaMemcpyAsync(host->planes7, o.planes7, N * (size_t)h->p.max_planes * 56, cudaMemcpyDeviceToHost, s));
            if (host->membership) HVO_CUDA(cudaMemcpyAsync(host->membership, o.membership, N * px * 4, cudaMemcpyDeviceToHost, s));
            if (host->membership8) HVO_CUDA(cudaMemcpyAsync(host->membership8, o.membership8, N * px, cudaMemcpyDeviceToHost, s));
        }
        timeline_mark(s, "end_planes");
        HVO_CUDA(cudaEventRecord(L.join[2], s));
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
            }
            if (L.normals) HVO_TRY(cudaMalloc(&o.normals8, B * (size_t)h->normals_count * 32));
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
        if (L.orb) hvo_orb_destroy(L.orb);
        if (L.line) hvo_line_destroy(L.line);
        if (L.plane) hvo_plane_destroy(L.plane);
        if (L.normals) hvo_normals_destroy(L.normals);
        hvo_frame_outputs& o = L.d_out;
        void* bufs[] = {L.d_gray, L.d_depth, o.kps, o.desc, o.kp_counts, o.kp_depth, o.kp_uright, o.keylines, o.line_desc, o.linevec3,
                        o.line_counts, o.n_planes, o.planes7, o.membership, o.membership8, o.normals8};
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
    if (h->p.stages & ST_PLANE) HVO_CHECK_ARG(o->n_planes && o->planes7 && (o->membership || (!device && o->membership8)), "plane outputs missing");
    if (h->p.stages & ST_NORMALS) HVO_CHECK_ARG(o->normals8, "normals output missing");
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
    const int per = (nframes + h->nlanes - 1) / h->nlanes;
    for (int off = 0; off < nframes; off += per, ++used) {
        FrameLane& L = h->lane[used];
        const int n = std::min(per, nframes - off);
        st = lane_launch(h, L, h->fork, d_gray + (size_t)off * px, d_depth16 + (size_t)off * px, n, off, outputs_at(h, *d_out, (size_t)off), nullptr, &launches,
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
    // chunk size: the lane capacity, or less when the batch is small so every lane gets work
    const int cap = h->lane[0].cap;
    const int per = std::min(cap, (nframes + h->nlanes - 1) / h->nlanes);
    int launches = 0, k = 0;
    for (int off = 0; off < nframes; off += per, ++k) {
        FrameLane& L = h->lane[k % h->nlanes];
        const int n = std::min(per, nframes - off);
        st = wait_joins(h, L.up, L);  // the lane's previous chunk (of this or an earlier call) must be done with the staging buffers
        if (st != HVO_OK) return st;
        HVO_CUDA(cudaMemcpyAsync(L.d_gray, gray + (size_t)off * px, (size_t)n * px, cudaMemcpyHostToDevice, L.up));
        HVO_CUDA(cudaMemcpyAsync(L.d_depth, depth16 + (size_t)off * px, (size_t)n * px * 2, cudaMemcpyHostToDevice, L.up));
        HVO_CUDA(cudaEventRecord(L.fork, L.up));
        hvo_frame_outputs d = L.d_out;
        if (!out->membership8) d.membership8 = nullptr;
        const hvo_frame_outputs hostk = outputs_at(h, *out, (size_t)off);
        st = lane_launch(h, L, L.fork, L.d_gray, L.d_depth, n, off, d, &hostk, &launches, h->last_lane >= 0 ? &h->lane[h->last_lane] : nullptr);
        if (st != HVO_OK) return st;
        h->last_lane = k % h->nlanes;
    }
    for (int li = 0; li < std::min(k, h->nlanes); ++li) {
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
