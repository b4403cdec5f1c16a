// Frame-level front-end for sm_100a: the extraction part of Frame::Frame(imGray, imDepth, ...) of the reference
// (src/Frame.cc:188-233), where three std::threads run ExtractORBNDepth (ORB + ComputeStereoFromRGBD, :874-884),
// ExtractLSD (LINEextractor::operator(), :895-903) and ComputePlanes (PlaneDetection + surface normals, :2104-2212)
// on the same frame.  Here the frame (gray + raw 16-bit depth) is uploaded once and the three pipelines run on three
// CUDA streams chained by events to a master stream, so a batch of frames is one call and one device-timed region.
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
}  // namespace hvo
using namespace hvo;

enum { ST_ORB = 1, ST_LINE = 2, ST_PLANE = 4, ST_NORMALS = 8 };

struct hvo_frame {
    hvo_frame_params p;
    int device = 0, width = 0, height = 0, max_batch = 0;
    hvo_orb* orb = nullptr;
    hvo_line* line = nullptr;
    hvo_plane* plane = nullptr;
    hvo_normals* normals = nullptr;
    cudaStream_t stream = nullptr;  // master: uploads, fork/join, timing
    cudaEvent_t fork = nullptr, join[4] = {nullptr, nullptr, nullptr, nullptr}, tev[2] = {nullptr, nullptr};
    uint8_t* d_gray = nullptr;
    uint16_t* d_depth = nullptr;
    hvo_frame_outputs d_out;  // device staging of every output (host API)
    int orb_cap = 0, max_lines = 0, normals_count = 0, last_launches = 0;
};

static int frame_launch(hvo_frame* h, const uint8_t* d_gray, const uint16_t* d_depth, int n, const hvo_frame_outputs& o,
                        const hvo_frame_outputs* host /* non-null: copy each stage's results back on its own stream */) {
    const size_t N = (size_t)n, px = (size_t)h->width * h->height;
    HVO_CUDA(cudaEventRecord(h->fork, h->stream));
    int launches = 0;
    // launch order = placement order: the ordered (latency-bound) pipelines first
    if (h->p.stages & ST_PLANE) {
        cudaStream_t s = plane_stream(h->plane);
        HVO_CUDA(cudaStreamWaitEvent(s, h->fork, 0));
        int st = hvo_plane_detect_batch_device(h->plane, d_depth, n, o.n_planes, o.planes7, h->p.max_planes, o.membership);
        if (st != HVO_OK) return st;
        launches += hvo_plane_last_launches(h->plane);
        if (host) {
            HVO_CUDA(cudaMemcpyAsync(host->n_planes, o.n_planes, N * 4, cudaMemcpyDeviceToHost, s));
            HVO_CUDA(cudaMemcpyAsync(host->planes7, o.planes7, N * (size_t)h->p.max_planes * 56, cudaMemcpyDeviceToHost, s));
            HVO_CUDA(cudaMemcpyAsync(host->membership, o.membership, N * px * 4, cudaMemcpyDeviceToHost, s));
        }
        HVO_CUDA(cudaEventRecord(h->join[2], s));
        HVO_CUDA(cudaStreamWaitEvent(h->stream, h->join[2], 0));
    }
    if (h->p.stages & ST_LINE) {
        cudaStream_t s = line_stream(h->line);
        HVO_CUDA(cudaStreamWaitEvent(s, h->fork, 0));
        int st = hvo_line_extract_batch_device(h->line, d_gray, n, o.keylines, o.line_desc, o.linevec3, o.line_counts);
        if (st != HVO_OK) return st;
        launches += hvo_line_last_launches(h->line);
        if (host) {
            const size_t c = (size_t)h->max_lines;
            HVO_CUDA(cudaMemcpyAsync(host->line_counts, o.line_counts, N * 4, cudaMemcpyDeviceToHost, s));
            HVO_CUDA(cudaMemcpyAsync(host->keylines, o.keylines, N * c * sizeof(hvo_keyline), cudaMemcpyDeviceToHost, s));
            HVO_CUDA(cudaMemcpyAsync(host->line_desc, o.line_desc, N * c * 32, cudaMemcpyDeviceToHost, s));
            if (host->linevec3) HVO_CUDA(cudaMemcpyAsync(host->linevec3, o.linevec3, N * c * 24, cudaMemcpyDeviceToHost, s));
        }
        HVO_CUDA(cudaEventRecord(h->join[1], s));
        HVO_CUDA(cudaStreamWaitEvent(h->stream, h->join[1], 0));
    }
    if (h->p.stages & ST_ORB) {
        cudaStream_t s = orb_stream(h->orb);
        HVO_CUDA(cudaStreamWaitEvent(s, h->fork, 0));
        hvo_rgbd_params rg{h->p.depth_factor, h->p.bf};
        int st = hvo_orb_extract_batch_device(h->orb, d_gray, n, o.kps, o.desc, o.kp_counts, d_depth, &rg, o.kp_depth, o.kp_uright);
        if (st != HVO_OK) return st;
        launches += hvo_orb_last_launches(h->orb);
        if (host) {
            const size_t c = (size_t)h->orb_cap;
            HVO_CUDA(cudaMemcpyAsync(host->kp_counts, o.kp_counts, N * 4, cudaMemcpyDeviceToHost, s));
            HVO_CUDA(cudaMemcpyAsync(host->kps, o.kps, N * c * sizeof(hvo_keypoint), cudaMemcpyDeviceToHost, s));
            HVO_CUDA(cudaMemcpyAsync(host->desc, o.desc, N * c * 32, cudaMemcpyDeviceToHost, s));
            if (host->kp_depth) HVO_CUDA(cudaMemcpyAsync(host->kp_depth, o.kp_depth, N * c * 4, cudaMemcpyDeviceToHost, s));
            if (host->kp_uright) HVO_CUDA(cudaMemcpyAsync(host->kp_uright, o.kp_uright, N * c * 4, cudaMemcpyDeviceToHost, s));
        }
        HVO_CUDA(cudaEventRecord(h->join[0], s));
        HVO_CUDA(cudaStreamWaitEvent(h->stream, h->join[0], 0));
    }
    if (h->p.stages & ST_NORMALS) {
        cudaStream_t s = normals_stream(h->normals);
        HVO_CUDA(cudaStreamWaitEvent(s, h->fork, 0));
        int st = hvo_normals_compute_batch_device(h->normals, d_depth, n, o.normals8);
        if (st != HVO_OK) return st;
        launches += 5;
        if (host) HVO_CUDA(cudaMemcpyAsync(host->normals8, o.normals8, N * (size_t)h->normals_count * 32, cudaMemcpyDeviceToHost, s));
        HVO_CUDA(cudaEventRecord(h->join[3], s));
        HVO_CUDA(cudaStreamWaitEvent(h->stream, h->join[3], 0));
    }
    h->last_launches = launches;
    return HVO_OK;
}

extern "C" {

int hvo_frame_create(const hvo_frame_params* p, int width, int height, int max_batch, int device, hvo_frame** out) {
    HVO_CHECK_ARG(out, "null out");
    *out = nullptr;
    HVO_CHECK_ARG(p, "null params");
    HVO_CHECK_ARG((p->stages & 15) != 0, "no stage selected");
    HVO_CHECK_ARG(max_batch >= 1 && p->max_planes >= 1, "max_batch / max_planes < 1");
    int ndev = 0;
    HVO_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev < 1) { set_error("no CUDA device: libhvofront has no CPU fallback"); return HVO_ERR_CUDA; }
    HVO_CHECK_ARG(device >= 0 && device < ndev, "device index out of range");
    hvo_frame* h = new (std::nothrow) hvo_frame();
    if (!h) { set_error("out of host memory"); return HVO_ERR_ARG; }
    h->p = *p; h->device = device; h->width = width; h->height = height; h->max_batch = max_batch;
    h->d_out = hvo_frame_outputs{};
    int st = HVO_OK;
    // lines and planes are chains of long latency-bound kernels (one warp / CTA per frame): they get the high-priority
    // streams so they are resident from the start; ORB and normals (streaming kernels with large grids) fill in around them
    int prio_lo = 0, prio_hi = 0;
    HVO_CUDA(cudaSetDevice(device));
    HVO_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    const char* env_prio = getenv("HVO_FRAME_PRIORITIES");  // tuning aid: "0" = all streams at the default priority
    if (env_prio && env_prio[0] == '0') prio_lo = prio_hi = 0;
    set_next_stream_priority(prio_lo);
    if (st == HVO_OK && (p->stages & ST_ORB)) st = hvo_orb_create(&p->orb, width, height, max_batch, device, &h->orb);
    set_next_stream_priority(prio_hi);
    if (st == HVO_OK && (p->stages & ST_LINE)) st = hvo_line_create(&p->line, width, height, max_batch, device, &h->line);
    if (st == HVO_OK && h->line) st = hvo_line_set_culling(h->line, p->line_cull);
    if (st == HVO_OK && (p->stages & ST_PLANE)) {
        hvo_plane_params pp{p->fx, p->fy, p->cx, p->cy, p->depth_factor};
        st = hvo_plane_create(&pp, width, height, max_batch, device, &h->plane);
    }
    set_next_stream_priority(prio_lo);
    if (st == HVO_OK && (p->stages & ST_NORMALS)) {
        hvo_normals_params np{p->fx, p->fy, p->cx, p->cy, p->depth_factor, 0.05f, 10.0f};  // Frame.cc:2179-2180
        st = hvo_normals_create(&np, width, height, max_batch, device, &h->normals);
    }
    set_next_stream_priority(0);
    if (st != HVO_OK) { hvo_frame_destroy(h); return st; }
    h->orb_cap = h->orb ? hvo_orb_capacity(h->orb) : 0;
    h->max_lines = h->line ? hvo_line_max_lines(h->line) : 0;
    h->normals_count = h->normals ? hvo_normals_count(h->normals) : 0;
    do {
#define HVO_TRY(call) if ((call) != cudaSuccess) { set_error("%s: %s", #call, cudaGetErrorString(cudaGetLastError())); st = HVO_ERR_CUDA; break; }
        HVO_TRY(cudaSetDevice(device));
        HVO_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        HVO_TRY(cudaEventCreateWithFlags(&h->fork, cudaEventDisableTiming));
        for (auto& e : h->join) HVO_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        if (st != HVO_OK) break;
        for (auto& e : h->tev) HVO_TRY(cudaEventCreate(&e));
        if (st != HVO_OK) break;
        const size_t B = (size_t)max_batch, px = (size_t)width * height;
        HVO_TRY(cudaMalloc(&h->d_gray, B * px));
        HVO_TRY(cudaMalloc(&h->d_depth, B * px * 2));
        hvo_frame_outputs& o = h->d_out;
        if (h->orb) {
            const size_t c = (size_t)h->orb_cap;
            HVO_TRY(cudaMalloc(&o.kps, B * c * sizeof(hvo_keypoint)));
            HVO_TRY(cudaMalloc(&o.desc, B * c * 32));
            HVO_TRY(cudaMalloc(&o.kp_counts, B * 4));
            HVO_TRY(cudaMalloc(&o.kp_depth, B * c * 4));
            HVO_TRY(cudaMalloc(&o.kp_uright, B * c * 4));
        }
        if (h->line) {
            const size_t c = (size_t)h->max_lines;
            HVO_TRY(cudaMalloc(&o.keylines, B * c * sizeof(hvo_keyline)));
            HVO_TRY(cudaMalloc(&o.line_desc, B * c * 32));
            HVO_TRY(cudaMalloc(&o.linevec3, B * c * 24));
            HVO_TRY(cudaMalloc(&o.line_counts, B * 4));
        }
        if (h->plane) {
            HVO_TRY(cudaMalloc(&o.n_planes, B * 4));
            HVO_TRY(cudaMalloc(&o.planes7, B * (size_t)p->max_planes * 56));
            HVO_TRY(cudaMalloc(&o.membership, B * px * 4));
        }
        if (h->normals) HVO_TRY(cudaMalloc(&o.normals8, B * (size_t)h->normals_count * 32));
#undef HVO_TRY
    } while (0);
    if (st != HVO_OK) { hvo_frame_destroy(h); return st; }
    *out = h;
    return HVO_OK;
}

void hvo_frame_destroy(hvo_frame* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->orb) hvo_orb_destroy(h->orb);
    if (h->line) hvo_line_destroy(h->line);
    if (h->plane) hvo_plane_destroy(h->plane);
    if (h->normals) hvo_normals_destroy(h->normals);
    hvo_frame_outputs& o = h->d_out;
    void* bufs[] = {h->d_gray, h->d_depth, o.kps, o.desc, o.kp_counts, o.kp_depth, o.kp_uright, o.keylines, o.line_desc, o.linevec3,
                    o.line_counts, o.n_planes, o.planes7, o.membership, o.normals8};
    for (void* b : bufs) if (b) cudaFree(b);
    if (h->fork) cudaEventDestroy(h->fork);
    for (auto& e : h->join) if (e) cudaEventDestroy(e);
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

static int frame_check_outputs(const hvo_frame* h, const hvo_frame_outputs* o) {
    HVO_CHECK_ARG(o, "null outputs");
    if (h->p.stages & ST_ORB) HVO_CHECK_ARG(o->kps && o->desc && o->kp_counts, "ORB outputs missing");
    if (h->p.stages & ST_LINE) HVO_CHECK_ARG(o->keylines && o->line_desc && o->line_counts, "line outputs missing");
    if (h->p.stages & ST_PLANE) HVO_CHECK_ARG(o->n_planes && o->planes7 && o->membership, "plane outputs missing");
    if (h->p.stages & ST_NORMALS) HVO_CHECK_ARG(o->normals8, "normals output missing");
    return HVO_OK;
}

int hvo_frame_extract_batch_device(hvo_frame* h, const uint8_t* d_gray, const uint16_t* d_depth16, int nframes, const hvo_frame_outputs* d_out) {
    HVO_CHECK_ARG(h && d_gray && d_depth16, "null argument");
    HVO_CHECK_ARG(nframes >= 1 && nframes <= h->max_batch, "nframes out of range for this handle");
    int st = frame_check_outputs(h, d_out);
    if (st != HVO_OK) return st;
    if (h->p.stages & ST_ORB) HVO_CHECK_ARG(d_out->kp_depth && d_out->kp_uright, "kp_depth / kp_uright missing");
    if (h->p.stages & ST_LINE) HVO_CHECK_ARG(d_out->linevec3, "linevec3 missing");
    HVO_CUDA(cudaSetDevice(h->device));
    return frame_launch(h, d_gray, d_depth16, nframes, *d_out, nullptr);
}

int hvo_frame_extract_batch(hvo_frame* h, const uint8_t* gray, const uint16_t* depth16, int nframes, const hvo_frame_outputs* out) {
    HVO_CHECK_ARG(h && gray && depth16, "null argument");
    HVO_CHECK_ARG(nframes >= 1 && nframes <= h->max_batch, "nframes out of range for this handle");
    int st = frame_check_outputs(h, out);
    if (st != HVO_OK) return st;
    HVO_CUDA(cudaSetDevice(h->device));
    const size_t N = (size_t)nframes, px = (size_t)h->width * h->height;
    HVO_CUDA(cudaMemcpyAsync(h->d_gray, gray, N * px, cudaMemcpyHostToDevice, h->stream));
    HVO_CUDA(cudaMemcpyAsync(h->d_depth, depth16, N * px * 2, cudaMemcpyHostToDevice, h->stream));
    st = frame_launch(h, h->d_gray, h->d_depth, nframes, h->d_out, out);
    if (st != HVO_OK) return st;
    HVO_CUDA(cudaStreamSynchronize(h->stream));
    return HVO_OK;
}

int hvo_frame_last_launches(const hvo_frame* h) { return h ? h->last_launches : 0; }
int hvo_frame_sync(hvo_frame* h) {
    HVO_CHECK_ARG(h, "null handle");
    HVO_CUDA(cudaSetDevice(h->device));
    HVO_CUDA(cudaStreamSynchronize(h->stream));
    return HVO_OK;
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
    return HVO_OK;
}

}  // extern "C"
