// Plane extraction (PEAC / agglomerative hierarchical clustering) for sm_100a.
// Replaces PlaneDetection::readDepthImage + runPlaneDetection (reference src/PlaneExtractor.cpp:26-66) and the
// ahc::PlaneFitter they drive (include/peac/AHCPlaneFitter.hpp, AHCPlaneSeg.hpp, AHCParamSet.hpp).
//
//   k_plane_blocks   (device) depth back-projection fused with the 10x10-block plane seeds: validity (missing data,
//                    right/down depth discontinuity), the nine second-order sums, centre, PCA normal, MSE.  The point
//                    cloud (7.4 MB / frame in the reference) is never materialised.  One thread per block walks its
//                    100 pixels in the reference's row-major order, so the double-precision sums are bit-identical.
//   host             the graph part (edges, min-MSE merge, block erosion, pixel flood fill, last merge, membership
//                    scan) is sequential and order-defined: it runs on the host from the block
//                    statistics, one frame per worker thread for batches.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iterator>
#include <limits>
#include <map>
#include <new>
#include <queue>
#include <thread>
#include <vector>

#include "hvo_common.cuh"

namespace hvo {

struct BlockOut {  // 96 bytes per block
    double s[9];   // sx sy sz sxx syy szz sxy syz sxz
    int N, queued; // queued: mse < T_mse(INIT) && !nouse
    double mse;
};

struct PlaneCam { double factor, fx, fy, cx, cy; };

__host__ __device__ inline void jacobi_eig33(const double K[3][3], double s[3], double V[3][3]) {
    double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = K[i][j];
    for (int sweep = 0; sweep < 60; ++sweep) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const double g = 100.0 * fabs(a[p][q]);
                if (sweep > 3 && fabs(a[p][p]) + g == fabs(a[p][p]) && fabs(a[q][q]) + g == fabs(a[q][q])) { a[p][q] = a[q][p] = 0.0; continue; }
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; ++k) { const double x = a[k][p], y = a[k][q]; a[k][p] = c * x - sn * y; a[k][q] = sn * x + c * y; }
                for (int k = 0; k < 3; ++k) { const double x = a[p][k], y = a[q][k]; a[p][k] = c * x - sn * y; a[q][k] = sn * x + c * y; }
                for (int k = 0; k < 3; ++k) { const double x = v[k][p], y = v[k][q]; v[k][p] = c * x - sn * y; v[k][q] = sn * x + c * y; }
            }
    }
    int o[3] = {0, 1, 2};
    const double d[3] = {a[0][0], a[1][1], a[2][2]};
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (d[o[j]] < d[o[i]]) { const int t = o[i]; o[i] = o[j]; o[j] = t; }
    for (int i = 0; i < 3; ++i) {
        s[i] = d[o[i]];
        for (int k = 0; k < 3; ++k) V[k][i] = v[k][o[i]];
    }
}

// Stats::compute (AHCPlaneSeg.hpp:125-163)
__host__ __device__ inline void stats_compute(const double s[9], int N, double center[3], double normal[3], double& mse, double& curv) {
    const double sc = 1.0 / N;
    center[0] = s[0] * sc; center[1] = s[1] * sc; center[2] = s[2] * sc;
    double K[3][3] = {{s[3] - s[0] * s[0] * sc, s[6] - s[0] * s[1] * sc, s[8] - s[0] * s[2] * sc},
                      {0, s[4] - s[1] * s[1] * sc, s[7] - s[1] * s[2] * sc},
                      {0, 0, s[5] - s[2] * s[2] * sc}};
    K[1][0] = K[0][1]; K[2][0] = K[0][2]; K[2][1] = K[1][2];
    double sv[3], V[3][3];
    jacobi_eig33(K, sv, V);
    const double sgn = (V[0][0] * center[0] + V[1][0] * center[1] + V[2][0] * center[2] <= 0) ? 1.0 : -1.0;
    normal[0] = sgn * V[0][0]; normal[1] = sgn * V[1][0]; normal[2] = sgn * V[2][0];
    mse = sv[0] * sc;
    curv = sv[0] / (sv[0] + sv[1] + sv[2]);
}

__global__ void __launch_bounds__(128) k_plane_blocks(const uint16_t* __restrict__ depth, int w, int h, PlaneCam cam, int Nw, int Nh,
                                                      BlockOut* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (b >= Nw * Nh) return;
    const int by = b / Nw, bx = b - by * Nw;
    const uint16_t* D = depth + (long long)f * w * h;
    double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int N = 0;
    bool valid = true;
    for (int ic = 0; ic < 10 && valid; ++ic) {
        const int i = by * 10 + ic;
        for (int jc = 0; jc < 10; ++jc) {
            const int j = bx * 10 + jc;
            const double z = (double)D[(long long)i * w + j] * cam.factor;
            if (z == 0) { valid = false; break; }                       // INIT_STRICT: one missing pixel rejects the block
            const double tdz = 0.04 * fabs(z) + 0.02;                   // ParamSet::T_dz
            if (j + 1 < w) { const double zn = (double)D[(long long)i * w + j + 1] * cam.factor; if (zn != 0 && fabs(z - zn) > tdz) { valid = false; break; } }
            if (i + 1 < h) { const double zn = (double)D[(long long)(i + 1) * w + j] * cam.factor; if (zn != 0 && fabs(z - zn) > tdz) { valid = false; break; } }
            const double x = ((double)j - cam.cx) * z / cam.fx, y = ((double)i - cam.cy) * z / cam.fy;
            s[0] += x; s[1] += y; s[2] += z;
            s[3] += x * x; s[4] += y * y; s[5] += z * z;
            s[6] += x * y; s[7] += y * z; s[8] += x * z;
            ++N;
        }
    }
    BlockOut o;
    if (!valid) { N = 0; for (int k = 0; k < 9; ++k) s[k] = 0; }
    for (int k = 0; k < 9; ++k) o.s[k] = s[k];
    o.N = N;
    o.queued = 0;
    o.mse = 0;
    if (N >= 4) {
        double c[3], n[3], mse, curv;
        stats_compute(s, N, c, n, mse, curv);
        const double t = 1.6e-6 * c[2] * c[2] + 5;  // ParamSet::T_mse(P_INIT): pow(depthSigma*z*z + stdTol_init, 2)
        o.mse = mse;
        o.queued = (mse < t * t) ? 1 : 0;
    }
    out[(long long)f * Nw * Nh + b] = o;
}

// ------------------------------------------------------------------------------------------------------------
// host graph stage (array-based; node ids grow with creation so "set order" == creation order)
// ------------------------------------------------------------------------------------------------------------
struct HNode {
    double s[9];
    int N, rid;
    double mse, center[3], normal[3];
    bool nouse;
    std::vector<int> nbs;  // sorted ascending, unique
};

static void nb_insert(std::vector<int>& v, int x) {
    auto it = std::lower_bound(v.begin(), v.end(), x);
    if (it == v.end() || *it != x) v.insert(it, x);
}
static void nb_erase(std::vector<int>& v, int x) {
    auto it = std::lower_bound(v.begin(), v.end(), x);
    if (it != v.end() && *it == x) v.erase(it);
}

struct HostAhc {
    int width, height, Nw, Nh, minSupport = 3000, maxStep = 100000;
    PlaneCam cam;
    const uint16_t* depth;
    std::vector<HNode> nodes;
    std::vector<int> extracted, parent, ssize, membershipImg, blkMap;
    std::vector<std::pair<int, int>> rfQueue;
    const double th_merge = std::cos(M_PI / 180.0 * 60.0), th_refine = std::cos(M_PI / 180.0 * 30.0);

    struct QCmp {
        const std::vector<HNode>* n;
        bool operator()(int a, int b) const { return (*n)[b].mse < (*n)[a].mse; }
    };
    typedef std::priority_queue<int, std::vector<int>, QCmp> MinQ;

    int Find(int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; }
    void Union(int x, int y) {
        const int xr = Find(x), yr = Find(y);
        if (xr == yr) return;
        if (ssize[xr] < ssize[yr]) { parent[xr] = yr; ssize[yr] += ssize[xr]; }
        else { parent[yr] = xr; ssize[xr] += ssize[yr]; }
    }
    double sim(const HNode& a, const HNode& b) const {
        return std::abs(a.normal[0] * b.normal[0] + a.normal[1] * b.normal[1] + a.normal[2] * b.normal[2]);
    }
    static double t_ang_init(double z) {
        const double z_near = 500, z_far = 4000, a_near = M_PI / 180.0 * 15.0, a_far = M_PI / 180.0 * 90.0;
        double cz = std::max(z, z_near);
        cz = std::min(cz, z_far);
        const double factor = (a_far - a_near) / (z_far - z_near);
        return std::cos(factor * cz + a_near - factor * z_near);
    }
    static double t_mse_merge(double z) { return std::pow(1.6e-6 * z * z + 8, 2); }
    void connect(int a, int b) { nb_insert(nodes[a].nbs, b); nb_insert(nodes[b].nbs, a); }
    void isolate(int a) {
        for (int nb : nodes[a].nbs) nb_erase(nodes[nb].nbs, a);
        nodes[a].nbs.clear();
    }
    bool point(int row, int col, double pt[3]) const {
        const double z = (double)depth[(size_t)row * width + col] * cam.factor;
        if (z == 0) return false;
        pt[0] = ((double)col - cam.cx) * z / cam.fx;
        pt[1] = ((double)row - cam.cy) * z / cam.fy;
        pt[2] = z;
        return true;
    }

    void cluster(MinQ& q) {
        int step = 0;
        while (!q.empty() && step <= maxStep) {
            const int p = q.top();
            q.pop();
            if (nodes[p].nouse) continue;
            int best = -1, best_nb = -1;
            HNode bestNode;
            const std::vector<int> nbs = nodes[p].nbs;
            for (int nb : nbs) {
                if (sim(nodes[p], nodes[nb]) < th_merge) continue;
                HNode m;
                for (int k = 0; k < 9; ++k) m.s[k] = nodes[p].s[k] + nodes[nb].s[k];
                m.N = nodes[p].N + nodes[nb].N;
                m.rid = nodes[p].N >= nodes[nb].N ? nodes[p].rid : nodes[nb].rid;
                m.nouse = false;
                double curv;
                stats_compute(m.s, m.N, m.center, m.normal, m.mse, curv);
                if (best < 0 || bestNode.mse > m.mse || (bestNode.mse == m.mse && bestNode.N < m.mse)) {  // N vs mse: AHCPlaneFitter.hpp:1045
                    best = 1; best_nb = nb; bestNode = m;
                }
            }
            if (best >= 0 && bestNode.mse < t_mse_merge(bestNode.center[2])) {
                const int nb = best_nb;
                Union(nodes[p].rid, nodes[nb].rid);
                std::vector<int> u;
                std::set_union(nodes[p].nbs.begin(), nodes[p].nbs.end(), nodes[nb].nbs.begin(), nodes[nb].nbs.end(), std::back_inserter(u));
                nb_erase(u, p);
                nb_erase(u, nb);
                bestNode.nbs = u;
                nodes.push_back(bestNode);
                const int id = (int)nodes.size() - 1;
                q.push(id);
                isolate(p);
                isolate(nb);
                for (int x : u) nb_insert(nodes[x].nbs, id);
                nodes[p].nouse = nodes[nb].nouse = true;
            } else {
                if (nodes[p].N >= minSupport) extracted.push_back(p);
                isolate(p);
            }
            ++step;
        }
        while (!q.empty()) {
            const int p = q.top();
            q.pop();
            if (nodes[p].N >= minSupport) extracted.push_back(p);
            isolate(p);
        }
        std::sort(extracted.begin(), extracted.end(), [this](int a, int b) { return nodes[b].N < nodes[a].N; });
    }

    static int valid4(int i, int j, int H, int W, int nbs[4]) {
        const int id = i * W + j;
        int c = 0;
        if (j > 0) nbs[c++] = id - 1;
        if (j < W - 1) nbs[c++] = id + 1;
        if (i > 0) nbs[c++] = id - W;
        if (i < H - 1) nbs[c++] = id + W;
        return c;
    }

    int run(const BlockOut* blocks, int32_t* membership, double* planes7, int max_planes) {
        nodes.clear(); extracted.clear(); rfQueue.clear();
        nodes.reserve(3 * (size_t)Nw * Nh);
        parent.resize(Nw * Nh); ssize.assign(Nw * Nh, 1);
        for (int i = 0; i < Nw * Nh; ++i) parent[i] = i;
        QCmp cmp{&nodes};
        MinQ q(cmp);
        // nodes of the initial graph (AHCPlaneFitter.hpp:786-826)
        std::vector<int> G(Nw * Nh, -1);
        for (int b = 0; b < Nw * Nh; ++b) {
            if (!blocks[b].queued) continue;
            HNode n;
            std::memcpy(n.s, blocks[b].s, sizeof(n.s));
            n.N = blocks[b].N; n.rid = b; n.nouse = false;
            double curv;
            stats_compute(n.s, n.N, n.center, n.normal, n.mse, curv);
            nodes.push_back(n);
            G[b] = (int)nodes.size() - 1;
            q.push(G[b]);
        }
        // edges (AHCPlaneFitter.hpp:896-954)
        for (int i = 0; i < Nh; ++i)
            for (int j = 1; j < Nw; j += 2) {
                const int c = i * Nw + j;
                if (G[c - 1] < 0) { --j; continue; }
                if (G[c] < 0) continue;
                if (j < Nw - 1 && G[c + 1] < 0) { ++j; continue; }
                const double th = t_ang_init(nodes[G[c]].center[2]);
                if ((j < Nw - 1 && sim(nodes[G[c - 1]], nodes[G[c + 1]]) >= th) || (j == Nw - 1 && sim(nodes[G[c]], nodes[G[c - 1]]) >= th)) {
                    connect(G[c], G[c - 1]);
                    if (j < Nw - 1) connect(G[c], G[c + 1]);
                } else {
                    --j;
                }
            }
        for (int j = 0; j < Nw; ++j)
            for (int i = 1; i < Nh; i += 2) {
                const int c = i * Nw + j;
                if (G[c - Nw] < 0) { --i; continue; }
                if (G[c] < 0) continue;
                if (i < Nh - 1 && G[c + Nw] < 0) { ++i; continue; }
                const double th = t_ang_init(nodes[G[c]].center[2]);
                if ((i < Nh - 1 && sim(nodes[G[c - Nw]], nodes[G[c + Nw]]) >= th) || (i == Nh - 1 && sim(nodes[G[c]], nodes[G[c - Nw]]) >= th)) {
                    connect(G[c], G[c - Nw]);
                    if (i < Nh - 1) connect(G[c], G[c + Nw]);
                } else {
                    --i;
                }
            }
        cluster(q);

        // refineDetails: block erosion + seeds (AHCPlaneFitter.hpp:485-587)
        std::map<int, int> rid2plid;
        for (int plid = 0; plid < (int)extracted.size(); ++plid) rid2plid.insert(std::make_pair(nodes[extracted[plid]].rid, plid));
        membershipImg.assign((size_t)width * height, -1);
        blkMap.assign(Nw * Nh, -1);
        std::vector<char> isValid(extracted.size(), 0);
        for (int i = 0, blkid = 0; i < Nh; ++i)
            for (int j = 0; j < Nw; ++j, ++blkid) {
                const int setid = Find(blkid);
                if (ssize[setid] * 100 >= minSupport) {
                    int nb4[4];
                    const int nn = valid4(i, j, Nh, Nw, nb4);
                    bool same = true;
                    for (int k = 0; k < nn; ++k)
                        if (Find(nb4[k]) != setid) { same = false; break; }
                    const int plid = rid2plid[setid];
                    if (same && plid < (int)isValid.size()) {
                        blkMap[blkid] = plid;
                        for (int y = i * 10; y < (i + 1) * 10; ++y)
                            for (int x = j * 10; x < (j + 1) * 10; ++x) membershipImg[(size_t)y * width + x] = plid;
                        isValid[plid] = 1;
                    }
                }
                if (blkMap[blkid] < 0) {
                    if (i > 0 && blkMap[blkid - Nw] >= 0) {
                        const int u = blkMap[blkid - Nw], sp = (i * 10 - 1) * width + j * 10;
                        for (int k = 1; k < 10; ++k) rfQueue.push_back(std::make_pair(sp + k, u));
                    }
                    if (j > 0 && blkMap[blkid - 1] >= 0) {
                        const int l = blkMap[blkid - 1], sp = (i * 10) * width + j * 10 - 1;
                        for (int k = 0; k < 9; ++k) rfQueue.push_back(std::make_pair(sp + k * width, l));
                    }
                } else {
                    const int plid = blkMap[blkid];
                    if (i > 0 && blkMap[blkid - Nw] != plid) {
                        const int sp = (i * 10) * width + j * 10;
                        for (int k = 0; k < 9; ++k) rfQueue.push_back(std::make_pair(sp + k, plid));
                    }
                    if (j > 0 && blkMap[blkid - 1] != plid) {
                        const int sp = (i * 10) * width + j * 10;
                        for (int k = 1; k < 10; ++k) rfQueue.push_back(std::make_pair(sp + k * width, plid));
                    }
                }
            }
        // pixel-level region growing (AHCPlaneFitter.hpp:428-476)
        {
            std::vector<float> distMap((size_t)height * width, std::numeric_limits<float>::max());
            for (size_t k = 0; k < rfQueue.size(); ++k) {
                const int sIdx = rfQueue[k].first, sy = sIdx / width, sx = sIdx - sy * width, plid = rfQueue[k].second;
                const HNode& pl = nodes[extracted[plid]];
                int nb4[4];
                const int nn = valid4(sy, sx, height, width, nb4);
                for (int it = 0; it < nn; ++it) {
                    const int cIdx = nb4[it];
                    int& trail = membershipImg[cIdx];
                    if (trail <= -6) continue;
                    if (trail >= 0 && trail == plid) continue;
                    const int cy = cIdx / width, cx = cIdx - cy * width;
                    const int by = cy / 10, bx = cx / 10;
                    const int blkid = (by < Nh && bx < Nw) ? by * Nw + bx : -1;
                    if (blkid >= 0 && blkMap[blkid] >= 0) continue;
                    double pt[3];
                    float cdist = -1;
                    bool ok = point(cy, cx, pt);
                    if (ok) {
                        cdist = (float)std::abs(pl.normal[0] * (pt[0] - pl.center[0]) + pl.normal[1] * (pt[1] - pl.center[1]) +
                                                pl.normal[2] * (pt[2] - pl.center[2]));
                        ok = std::pow((double)cdist, 2) < 9 * pl.mse + 1e-5;
                    }
                    if (ok) {
                        if (trail >= 0 && sim(pl, nodes[extracted[trail]]) >= th_refine) connect(extracted[trail], extracted[plid]);
                        float& od = distMap[cIdx];
                        if (cdist < od) { trail = plid; od = cdist; rfQueue.push_back(std::make_pair(cIdx, plid)); }
                        else if (trail < 0) trail -= 1;
                    } else if (trail < 0) {
                        trail -= 1;
                    }
                }
            }
        }
        // last merge among the refined planes, then relabel (AHCPlaneFitter.hpp:317-371)
        std::vector<int> old;
        extracted.swap(old);
        MinQ q2(cmp);
        for (size_t i = 0; i < old.size(); ++i)
            if (isValid[i]) q2.push(old[i]);
        cluster(q2);
        std::vector<int> plidmap(old.size(), -1);
        for (size_t i = 0; i < old.size(); ++i) {
            if (!isValid[i]) continue;
            const int r = Find(nodes[old[i]].rid);
            for (size_t j = 0; j < extracted.size(); ++j)
                if (r == nodes[extracted[j]].rid) { plidmap[i] = (int)j; break; }
        }
        for (size_t i = 0; i < membershipImg.size(); ++i) {
            const int plid = membershipImg[i];
            membership[i] = (plid >= 0 && plidmap[plid] >= 0) ? plidmap[plid] : -1;
        }
        const int n = (int)extracted.size();
        for (int i = 0; i < n && i < max_planes; ++i) {
            const HNode& p = nodes[extracted[i]];
            double* o = planes7 + 7 * (size_t)i;
            for (int k = 0; k < 3; ++k) { o[k] = p.normal[k]; o[3 + k] = p.center[k]; }
            o[6] = p.N;
        }
        return n;
    }
};

}  // namespace hvo

using namespace hvo;

struct hvo_plane {
    int device = 0, width = 0, height = 0, max_batch = 0, Nw = 0, Nh = 0;
    PlaneCam cam;
    cudaStream_t stream = nullptr;
    cudaEvent_t tev[2] = {nullptr, nullptr};
    uint16_t* d_depth = nullptr;
    BlockOut* d_blocks = nullptr;
    BlockOut* h_blocks = nullptr;  // pinned
    int host_threads = 1;
};

extern "C" {

int hvo_plane_create(const hvo_plane_params* p, int width, int height, int max_batch, int device, hvo_plane** out) {
    HVO_CHECK_ARG(p && out, "null argument");
    *out = nullptr;
    HVO_CHECK_ARG(width >= 20 && height >= 20 && max_batch >= 1, "size out of range");
    HVO_CHECK_ARG(p->fx != 0.f && p->fy != 0.f, "focal length is zero");
    int ndev = 0;
    HVO_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev < 1) { set_error("no CUDA device: libhvofront has no CPU fallback"); return HVO_ERR_CUDA; }
    HVO_CHECK_ARG(device >= 0 && device < ndev, "device index out of range");
    hvo_plane* h = new (std::nothrow) hvo_plane();
    if (!h) { set_error("out of host memory"); return HVO_ERR_ARG; }
    h->device = device; h->width = width; h->height = height; h->max_batch = max_batch;
    h->Nw = width / 10; h->Nh = height / 10;
    h->cam.factor = (double)p->depth_factor; h->cam.fx = (double)p->fx; h->cam.fy = (double)p->fy;
    h->cam.cx = (double)p->cx; h->cam.cy = (double)p->cy;
    h->host_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&h->tev[0]);
    if (e == cudaSuccess) e = cudaEventCreate(&h->tev[1]);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_depth, (size_t)max_batch * width * height * 2);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_blocks, (size_t)max_batch * h->Nw * h->Nh * sizeof(BlockOut));
    if (e == cudaSuccess) e = cudaMallocHost(&h->h_blocks, (size_t)max_batch * h->Nw * h->Nh * sizeof(BlockOut));
    if (e != cudaSuccess) { set_error("hvo_plane_create: %s", cudaGetErrorString(e)); hvo_plane_destroy(h); return HVO_ERR_CUDA; }
    *out = h;
    return HVO_OK;
}

void hvo_plane_destroy(hvo_plane* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->d_depth) cudaFree(h->d_depth);
    if (h->d_blocks) cudaFree(h->d_blocks);
    if (h->h_blocks) cudaFreeHost(h->h_blocks);
    for (auto& e : h->tev) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

static int plane_blocks_launch(hvo_plane* h, const uint16_t* d_depth, int nframes) {
    const int nb = h->Nw * h->Nh;
    k_plane_blocks<<<dim3(div_up(nb, 128), nframes), 128, 0, h->stream>>>(d_depth, h->width, h->height, h->cam, h->Nw, h->Nh, h->d_blocks);
    HVO_CUDA(cudaGetLastError());
    return HVO_OK;
}

/* device-only leg (bench / roofline): block statistics of nframes device-resident depth images */
int hvo_plane_blocks_device(hvo_plane* h, const uint16_t* d_depth, int nframes) {
    HVO_CHECK_ARG(h && d_depth, "null argument");
    HVO_CHECK_ARG(nframes >= 1 && nframes <= h->max_batch, "nframes out of range for this handle");
    HVO_CUDA(cudaSetDevice(h->device));
    return plane_blocks_launch(h, d_depth, nframes);
}

int hvo_plane_get_blocks(hvo_plane* h, int frame, double* out9) {
    HVO_CHECK_ARG(h && out9, "null argument");
    HVO_CHECK_ARG(frame >= 0 && frame < h->max_batch, "frame out of range");
    HVO_CUDA(cudaSetDevice(h->device));
    const int nb = h->Nw * h->Nh;
    HVO_CUDA(cudaMemcpyAsync(h->h_blocks, h->d_blocks + (size_t)frame * nb, (size_t)nb * sizeof(BlockOut), cudaMemcpyDeviceToHost, h->stream));
    HVO_CUDA(cudaStreamSynchronize(h->stream));
    for (int b = 0; b < nb; ++b) {
        const BlockOut& o = h->h_blocks[b];
        double* r = out9 + 9 * (size_t)b;
        r[0] = o.queued; r[1] = o.N;
        if (o.N >= 4) {
            double c[3], n[3], mse, curv;
            stats_compute(o.s, o.N, c, n, mse, curv);
            for (int k = 0; k < 3; ++k) { r[2 + k] = c[k]; r[5 + k] = n[k]; }
            r[8] = o.mse;
        } else {
            for (int k = 2; k < 8; ++k) r[k] = 0;
            r[8] = std::numeric_limits<double>::quiet_NaN();
        }
    }
    return HVO_OK;
}

int hvo_plane_detect_batch(hvo_plane* h, const uint16_t* depth16, int nframes, int32_t* n_planes, double* planes7, int max_planes,
                           int32_t* membership) {
    HVO_CHECK_ARG(h && depth16 && n_planes && planes7 && membership, "null argument");
    HVO_CHECK_ARG(nframes >= 1 && nframes <= h->max_batch && max_planes >= 1, "nframes / max_planes out of range");
    HVO_CUDA(cudaSetDevice(h->device));
    const size_t px = (size_t)h->width * h->height, nb = (size_t)h->Nw * h->Nh;
    HVO_CUDA(cudaMemcpyAsync(h->d_depth, depth16, (size_t)nframes * px * 2, cudaMemcpyHostToDevice, h->stream));
    int st = plane_blocks_launch(h, h->d_depth, nframes);
    if (st != HVO_OK) return st;
    HVO_CUDA(cudaMemcpyAsync(h->h_blocks, h->d_blocks, (size_t)nframes * nb * sizeof(BlockOut), cudaMemcpyDeviceToHost, h->stream));
    HVO_CUDA(cudaStreamSynchronize(h->stream));
    // host graph stage, one frame per worker
    const int nt = std::min(nframes, h->host_threads);
    auto work = [&](int t) {
        HostAhc a;
        a.width = h->width; a.height = h->height; a.Nw = h->Nw; a.Nh = h->Nh; a.cam = h->cam;
        for (int f = t; f < nframes; f += nt) {
            a.depth = depth16 + (size_t)f * px;
            n_planes[f] = a.run(h->h_blocks + (size_t)f * nb, membership + (size_t)f * px, planes7 + (size_t)f * max_planes * 7, max_planes);
        }
    };
    if (nt <= 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    return HVO_OK;
}

int hvo_plane_detect(hvo_plane* h, const uint16_t* depth16, int32_t* n_planes, double* planes7, int max_planes, int32_t* membership) {
    return hvo_plane_detect_batch(h, depth16, 1, n_planes, planes7, max_planes, membership);
}

int hvo_plane_sync(hvo_plane* h) {
    HVO_CHECK_ARG(h, "null handle");
    HVO_CUDA(cudaSetDevice(h->device));
    HVO_CUDA(cudaStreamSynchronize(h->stream));
    return HVO_OK;
}
int hvo_plane_timer_start(hvo_plane* h) {
    HVO_CHECK_ARG(h, "null handle");
    HVO_CUDA(cudaSetDevice(h->device));
    HVO_CUDA(cudaEventRecord(h->tev[0], h->stream));
    return HVO_OK;
}
int hvo_plane_timer_stop(hvo_plane* h, float* ms_out) {
    HVO_CHECK_ARG(h && ms_out, "null argument");
    HVO_CUDA(cudaSetDevice(h->device));
    HVO_CUDA(cudaEventRecord(h->tev[1], h->stream));
    HVO_CUDA(cudaEventSynchronize(h->tev[1]));
    HVO_CUDA(cudaEventElapsedTime(ms_out, h->tev[0], h->tev[1]));
    return HVO_OK;
}

}  // extern "C"
