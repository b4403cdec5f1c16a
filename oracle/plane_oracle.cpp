// TEST INFRASTRUCTURE — CPU oracle for plane extraction (PEAC / AHC), never on the product path.
//
// Restates (paths relative to /root/reference):
//   PlaneDetection::readDepthImage / ImagePointCloud::get   src/PlaneExtractor.cpp:26-58, include/PlaneExtractor.h:18-34
//   ahc::PlaneSeg (block ctor, Stats::compute, merge ctor)  include/peac/AHCPlaneSeg.hpp:41-44, 60-163, 211-310
//   ahc::ParamSet thresholds                                 include/peac/AHCParamSet.hpp:68-146
//   ahc::PlaneFitter::run / initGraph / ahCluster            include/peac/AHCPlaneFitter.hpp:211-260, 786-954, 983-1189
//   refineDetails / findBlockMembership / floodFill          include/peac/AHCPlaneFitter.hpp:299-379, 428-476, 485-587
//   DisjointSet                                              include/peac/DisjointSet.hpp
// Eigen's SelfAdjointEigenSolver (un-vendored) is replaced by eig33_smallest below (Newton on the characteristic
// polynomial + cross-product eigenvector: only + - * / sqrt, so the CUDA side can be bit-identical); it agrees with
// numpy.linalg.eigh to < 1e-7 rad, far inside the 1e-3 rad tolerance of the plane-normal parity bar
// (tests/test_planes.py).  A cyclic Jacobi solver (eig33sym) is kept as an independent cross-check.  Deterministic choices: neighbour sets iterate in node-creation
// order (the reference iterates std::set<PlaneSeg*> in heap-address order; only exact MSE ties can differ).
// PARITY of the sequential graph logic is pinned by the source text only (PEAC needs OpenCV + Eigen to build).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <queue>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

namespace peaco {

// ---- 3x3 symmetric eigen decomposition: s ascending, V[:, i] <-> s[i] -------------------------------------
static void eig33sym(const double K[3][3], double s[3], double V[3][3]) {
    double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = K[i][j];
    for (int sweep = 0; sweep < 60; ++sweep) {
        const double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
        if (off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const double g = 100.0 * std::fabs(a[p][q]);
                // off-diagonal element negligible against both diagonal entries: annihilate it without a rotation
                if (sweep > 3 && std::fabs(a[p][p]) + g == std::fabs(a[p][p]) && std::fabs(a[q][q]) + g == std::fabs(a[q][q])) {
                    a[p][q] = a[q][p] = 0.0;
                    continue;
                }
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - sn * akq;
                    a[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - sn * aqk;
                    a[q][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - sn * vkq;
                    v[k][q] = sn * vkp + c * vkq;
                }
            }
    }
    int o[3] = {0, 1, 2};
    double d[3] = {a[0][0], a[1][1], a[2][2]};
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (d[o[j]] < d[o[i]]) std::swap(o[i], o[j]);
    for (int i = 0; i < 3; ++i) {
        s[i] = d[o[i]];
        for (int k = 0; k < 3; ++k) V[k][i] = v[k][o[i]];
    }
}

// Smallest eigenpair of a symmetric positive semi-definite 3x3 matrix (a covariance) with + - * / sqrt only, so the CPU
// oracle and the CUDA kernels produce bit-identical results: Newton's iteration on the characteristic polynomial
// p(x) = det(K - xI) = -x^3 + c2 x^2 - c1 x + c0 from x = 0 (p is convex and decreasing on (-inf, lambda_min], so the
// iterates approach lambda_min monotonically), then the eigenvector as the largest of the three row cross products
// of K - lambda I.  Accuracy on plane covariances: |d lambda| <= 2e-12 lambda_max, direction error < 1e-7 rad.
static void eig33_smallest(const double K[3][3], double& lam, double v[3]) {
    const double a = K[0][0], b = K[1][1], c = K[2][2], d = K[0][1], e = K[0][2], f = K[1][2];
    const double c2 = a + b + c;
    const double c1 = (a * b - d * d) + (a * c - e * e) + (b * c - f * f);
    const double c0 = a * (b * c - f * f) - d * (d * c - f * e) + e * (d * f - b * e);
    double x = 0, prev = INFINITY;
    for (int k = 0; k < 40; ++k) {
        const double p = ((-x + c2) * x - c1) * x + c0;
        const double dp = (-3 * x + 2 * c2) * x - c1;
        if (dp == 0) break;
        const double dx = p / dp, adx = std::fabs(dx);
        if (k >= 2 && adx >= prev) break;  // rounding floor reached
        x -= dx;
        prev = adx;
        if (adx <= 1e-16 * c2) break;
    }
    lam = x;
    const double r0[3] = {a - x, d, e}, r1[3] = {d, b - x, f}, r2[3] = {e, f, c - x};
    const double u0[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
    const double u1[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]};
    const double u2[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
    const double n0 = u0[0] * u0[0] + u0[1] * u0[1] + u0[2] * u0[2];
    const double n1 = u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2];
    const double n2 = u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2];
    const double* u = u0;
    double n = n0;
    if (n1 > n) { u = u1; n = n1; }
    if (n2 > n) { u = u2; n = n2; }
    if (n > 0) {
        const double s = std::sqrt(n);
        v[0] = u[0] / s; v[1] = u[1] / s; v[2] = u[2] / s;
    } else {
        v[0] = 0; v[1] = 0; v[2] = 1;
    }
}

struct Stats {
    double sx = 0, sy = 0, sz = 0, sxx = 0, syy = 0, szz = 0, sxy = 0, syz = 0, sxz = 0;
    int N = 0;
    void push(double x, double y, double z) {
        sx += x; sy += y; sz += z;
        sxx += x * x; syy += y * y; szz += z * z;
        sxy += x * y; syz += y * z; sxz += x * z;
        ++N;
    }
    static Stats merged(const Stats& a, const Stats& b) {
        Stats r;
        r.sx = a.sx + b.sx; r.sy = a.sy + b.sy; r.sz = a.sz + b.sz;
        r.sxx = a.sxx + b.sxx; r.syy = a.syy + b.syy; r.szz = a.szz + b.szz;
        r.sxy = a.sxy + b.sxy; r.syz = a.syz + b.syz; r.sxz = a.sxz + b.sxz;
        r.N = a.N + b.N;
        return r;
    }
    void compute(double center[3], double normal[3], double& mse, double& curvature) const {
        const double sc = ((double)1.0) / N;
        center[0] = sx * sc; center[1] = sy * sc; center[2] = sz * sc;
        double K[3][3] = {{sxx - sx * sx * sc, sxy - sx * sy * sc, sxz - sx * sz * sc},
                          {0, syy - sy * sy * sc, syz - sy * sz * sc},
                          {0, 0, szz - sz * sz * sc}};
        K[1][0] = K[0][1]; K[2][0] = K[0][2]; K[2][1] = K[1][2];
        double lam, v[3];
        eig33_smallest(K, lam, v);
        const double sgn = (v[0] * center[0] + v[1] * center[1] + v[2] * center[2] <= 0) ? 1.0 : -1.0;
        normal[0] = sgn * v[0]; normal[1] = sgn * v[1]; normal[2] = sgn * v[2];
        mse = lam * sc;
        curvature = lam / (K[0][0] + K[1][1] + K[2][2]);
    }
};

struct Params {
    double depthSigma = 1.6e-6, stdTol_init = 5, stdTol_merge = 8, z_near = 500, z_far = 4000;
    double angle_near = M_PI / 180.0 * 15.0, angle_far = M_PI / 180.0 * 90.0;
    double similarityTh_merge = std::cos(M_PI / 180.0 * 60.0), similarityTh_refine = std::cos(M_PI / 180.0 * 30.0);
    double depthAlpha = 0.04, depthChangeTol = 0.02;
    double T_mse_init(double z) const { return std::pow(depthSigma * z * z + stdTol_init, 2); }
    double T_mse_merge(double z) const { return std::pow(depthSigma * z * z + stdTol_merge, 2); }
    double T_ang_init(double z) const {
        double cz = std::max(z, z_near);
        cz = std::min(cz, z_far);
        const double factor = (angle_far - angle_near) / (z_far - z_near);
        return std::cos(factor * cz + angle_near - factor * z_near);
    }
    double T_dz(double z) const { return depthAlpha * std::fabs(z) + depthChangeTol; }
};

struct Cloud {
    const uint16_t* depth;
    int w, h;
    double factor, fx, fy, cx, cy;  // already promoted from float
    bool get(int row, int col, double& x, double& y, double& z) const {
        z = (double)depth[(size_t)row * w + col] * factor;
        if (z == 0 || std::isnan(z)) return false;
        x = ((double)col - cx) * z / fx;
        y = ((double)row - cy) * z / fy;
        return true;
    }
};

struct Seg {
    Stats st;
    int rid = 0, N = 0;
    double mse = 0, center[3] = {0, 0, 0}, normal[3] = {0, 0, 0}, curvature = 0;
    bool nouse = false;
    std::set<int> nbs;  // node ids; id order == creation order
};

struct DisjointSet {
    std::vector<int> parent, size;
    explicit DisjointSet(int n) : parent(n), size(n, 1) { for (int i = 0; i < n; ++i) parent[i] = i; }
    int Find(int x) { if (parent[x] != x) parent[x] = Find(parent[x]); return parent[x]; }
    int getSetSize(int x) { return size[Find(x)]; }
    int Union(int x, int y) {
        const int xr = Find(x), yr = Find(y);
        if (xr == yr) return xr;
        if (size[xr] < size[yr]) { parent[xr] = yr; size[yr] += size[xr]; return yr; }
        parent[yr] = xr; size[xr] += size[yr]; return xr;
    }
};

struct Fitter {
    Cloud cloud;
    Params params;
    int width, height, winW = 10, winH = 10, minSupport = 3000, maxStep = 100000;
    std::vector<Seg> nodes;
    std::vector<int> extracted;  // node ids
    std::vector<int> membershipImg, blkMap;
    std::vector<std::pair<int, int>> rfQueue;
    DisjointSet* ds = nullptr;

    struct QCmp {
        const std::vector<Seg>* n;
        bool operator()(int a, int b) const { return (*n)[b].mse < (*n)[a].mse; }
    };
    typedef std::priority_queue<int, std::vector<int>, QCmp> MinQ;

    double similarity(const Seg& a, const Seg& b) const {
        return std::abs(a.normal[0] * b.normal[0] + a.normal[1] * b.normal[1] + a.normal[2] * b.normal[2]);
    }
    void connect(int a, int b) { nodes[a].nbs.insert(b); nodes[b].nbs.insert(a); }
    void disconnectAll(int a) {
        for (int nb : nodes[a].nbs) nodes[nb].nbs.erase(a);
        nodes[a].nbs.clear();
    }

    int makeBlock(int rid, int seed_row, int seed_col) {
        Seg s;
        s.rid = rid;
        bool valid = true;
        for (int i = seed_row, ic = 0; ic < winH && i < height; ++i, ++ic) {
            for (int j = seed_col, jc = 0; jc < winW && j < width; ++j, ++jc) {
                double x = 0, y = 0, z = 10000;
                if (!cloud.get(i, j, x, y, z)) { valid = false; break; }  // INIT_STRICT
                double xn = 0, yn = 0, zn = 10000;
                if (j + 1 < width && (cloud.get(i, j + 1, xn, yn, zn) && std::fabs(z - zn) > params.T_dz(z))) { valid = false; break; }
                if (i + 1 < height && (cloud.get(i + 1, j, xn, yn, zn) && std::fabs(z - zn) > params.T_dz(z))) { valid = false; break; }
                s.st.push(x, y, z);
            }
            if (!valid) break;
        }
        if (valid) { s.nouse = false; s.N = s.st.N; }
        else { s.N = 0; s.st = Stats(); s.nouse = true; }
        if (s.N < 4) s.mse = s.curvature = std::numeric_limits<double>::quiet_NaN();
        else s.st.compute(s.center, s.normal, s.mse, s.curvature);
        nodes.push_back(s);
        return (int)nodes.size() - 1;
    }

    void initGraph(MinQ& minQ, std::vector<int>& G) {
        const int Nh = height / winH, Nw = width / winW;
        G.assign(Nh * Nw, -1);
        for (int i = 0; i < Nh; ++i)
            for (int j = 0; j < Nw; ++j) {
                const int id = makeBlock(i * Nw + j, i * winH, j * winW);
                const Seg& p = nodes[id];
                if (p.mse < params.T_mse_init(p.center[2]) && !p.nouse) { G[i * Nw + j] = id; minQ.push(id); }
            }
        for (int i = 0; i < Nh; ++i)
            for (int j = 1; j < Nw; j += 2) {
                const int c = i * Nw + j;
                if (G[c - 1] < 0) { --j; continue; }
                if (G[c] < 0) continue;
                if (j < Nw - 1 && G[c + 1] < 0) { ++j; continue; }
                const double th = params.T_ang_init(nodes[G[c]].center[2]);
                if ((j < Nw - 1 && similarity(nodes[G[c - 1]], nodes[G[c + 1]]) >= th) ||
                    (j == Nw - 1 && similarity(nodes[G[c]], nodes[G[c - 1]]) >= th)) {
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
                const double th = params.T_ang_init(nodes[G[c]].center[2]);
                if ((i < Nh - 1 && similarity(nodes[G[c - Nw]], nodes[G[c + Nw]]) >= th) ||
                    (i == Nh - 1 && similarity(nodes[G[c]], nodes[G[c - Nw]]) >= th)) {
                    connect(G[c], G[c - Nw]);
                    if (i < Nh - 1) connect(G[c], G[c + Nw]);
                } else {
                    --i;
                }
            }
    }

    long st_pops = 0, st_merges = 0, st_nbs = 0, st_batches = 0;  // diagnostics (HVO_ORACLE_STATS=1)
    int ahCluster(MinQ& minQ) {
        int step = 0;
        while (!minQ.empty() && step <= maxStep) {
            const int p = minQ.top();
            minQ.pop();
            if (nodes[p].nouse) continue;
            int cand = -1, cand_nb = -1;
            const std::vector<int> nbs(nodes[p].nbs.begin(), nodes[p].nbs.end());
            ++st_pops; st_nbs += (long)nbs.size(); st_batches += ((long)nbs.size() + 31) / 32;
            for (int nb : nbs) {
                if (similarity(nodes[p], nodes[nb]) < params.similarityTh_merge) continue;
                Seg m;
                m.st = Stats::merged(nodes[p].st, nodes[nb].st);
                m.nouse = false;
                m.rid = nodes[p].N >= nodes[nb].N ? nodes[p].rid : nodes[nb].rid;
                m.N = m.st.N;
                m.st.compute(m.center, m.normal, m.mse, m.curvature);
                nodes.push_back(m);
                const int mid = (int)nodes.size() - 1;
                if (cand < 0 || nodes[cand].mse > nodes[mid].mse ||
                    (nodes[cand].mse == nodes[mid].mse && nodes[cand].N < nodes[mid].mse)) {  // sic: N vs mse (AHCPlaneFitter.hpp:1045)
                    cand = mid;
                    cand_nb = nb;
                }
            }
            if (cand >= 0 && nodes[cand].mse < params.T_mse_merge(nodes[cand].center[2])) {
                minQ.push(cand); ++st_merges;
                ds->Union(nodes[p].rid, nodes[cand_nb].rid);
                std::set<int>& n = nodes[cand].nbs;
                n.insert(nodes[p].nbs.begin(), nodes[p].nbs.end());
                n.insert(nodes[cand_nb].nbs.begin(), nodes[cand_nb].nbs.end());
                n.erase(p);
                n.erase(cand_nb);
                disconnectAll(p);
                disconnectAll(cand_nb);
                for (int nb : nodes[cand].nbs) nodes[nb].nbs.insert(cand);
                nodes[p].nouse = nodes[cand_nb].nouse = true;
            } else {
                if (nodes[p].N >= minSupport) extracted.push_back(p);
                disconnectAll(p);
            }
            ++step;
        }
        while (!minQ.empty()) {
            const int p = minQ.top();
            minQ.pop();
            if (nodes[p].N >= minSupport) extracted.push_back(p);
            disconnectAll(p);
        }
        std::sort(extracted.begin(), extracted.end(), [this](int a, int b) { return nodes[b].N < nodes[a].N; });
        return step;
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
    int blockIdx(int px, int py) const {
        const int Nw = width / winW, Nh = height / winH, by = py / winH, bx = px / winW;
        return (by < Nh && bx < Nw) ? (by * Nw + bx) : -1;
    }

    void findBlockMembership(std::vector<bool>& isValid) {
        std::map<int, int> rid2plid;
        for (int plid = 0; plid < (int)extracted.size(); ++plid) rid2plid.insert(std::make_pair(nodes[extracted[plid]].rid, plid));
        const int Nh = height / winH, Nw = width / winW, npb = winH * winW;
        membershipImg.assign((size_t)width * height, -1);
        blkMap.assign(Nh * Nw, -1);
        isValid.assign(extracted.size(), false);
        for (int i = 0, blkid = 0; i < Nh; ++i)
            for (int j = 0; j < Nw; ++j, ++blkid) {
                const int setid = ds->Find(blkid);
                const int setSize = ds->getSetSize(setid) * npb;
                if (setSize >= minSupport) {
                    int nbs[4];
                    const int nn = valid4(i, j, Nh, Nw, nbs);
                    bool same = true;
                    for (int k = 0; k < nn; ++k)
                        if (ds->Find(nbs[k]) != setid) { same = false; break; }  // ERODE_ALL_BORDER
                    const int plid = rid2plid[setid];  // sic: operator[] inserts 0 for unknown roots
                    if (same) {
                        blkMap[blkid] = plid;
                        for (int y = i * winH; y < (i + 1) * winH; ++y)
                            for (int x = j * winW; x < (j + 1) * winW; ++x) membershipImg[(size_t)y * width + x] = plid;
                        isValid[plid] = true;
                    } else {
                        blkMap[blkid] = -1;
                    }
                } else {
                    blkMap[blkid] = -1;
                }
                if (blkMap[blkid] < 0) {
                    if (i > 0 && blkMap[blkid - Nw] >= 0) {
                        const int u = blkMap[blkid - Nw], sp = (i * winH - 1) * width + j * winW;
                        for (int k = 1; k < winW; ++k) rfQueue.push_back(std::make_pair(sp + k, u));
                    }
                    if (j > 0 && blkMap[blkid - 1] >= 0) {
                        const int l = blkMap[blkid - 1], sp = (i * winH) * width + j * winW - 1;
                        for (int k = 0; k < winH - 1; ++k) rfQueue.push_back(std::make_pair(sp + k * width, l));
                    }
                } else {
                    const int plid = blkMap[blkid];
                    if (i > 0 && blkMap[blkid - Nw] != plid) {
                        const int sp = (i * winH) * width + j * winW;
                        for (int k = 0; k < winW - 1; ++k) rfQueue.push_back(std::make_pair(sp + k, plid));
                    }
                    if (j > 0 && blkMap[blkid - 1] != plid) {
                        const int sp = (i * winH) * width + j * winW;
                        for (int k = 1; k < winH; ++k) rfQueue.push_back(std::make_pair(sp + k * width, plid));
                    }
                }
            }
    }

    void floodFill() {
        std::vector<float> distMap((size_t)height * width, std::numeric_limits<float>::max());
        for (int k = 0; k < (int)rfQueue.size(); ++k) {
            const int sIdx = rfQueue[k].first, seedy = sIdx / width, seedx = sIdx - seedy * width, plid = rfQueue[k].second;
            const Seg& pl = nodes[extracted[plid]];
            int nbs[4];
            const int nn = valid4(seedy, seedx, height, width, nbs);
            for (int it = 0; it < nn; ++it) {
                const int cIdx = nbs[it];
                int& trail = membershipImg[cIdx];
                if (trail <= -6) continue;
                if (trail >= 0 && trail == plid) continue;
                const int cy = cIdx / width, cx = cIdx - cy * width;
                const int blkid = blockIdx(cx, cy);
                if (blkid >= 0 && blkMap[blkid] >= 0) continue;
                double pt[3] = {0, 0, 0};
                float cdist = -1;
                bool ok = cloud.get(cy, cx, pt[0], pt[1], pt[2]);
                if (ok) {
                    cdist = (float)std::abs(pl.normal[0] * (pt[0] - pl.center[0]) + pl.normal[1] * (pt[1] - pl.center[1]) +
                                            pl.normal[2] * (pt[2] - pl.center[2]));
                    ok = std::pow(cdist, 2) < 9 * pl.mse + 1e-5;
                }
                if (ok) {
                    if (trail >= 0) {
                        Seg& n_pl = nodes[extracted[trail]];
                        if (similarity(pl, n_pl) >= params.similarityTh_refine) connect(extracted[trail], extracted[plid]);
                    }
                    float& old_dist = distMap[cIdx];
                    if (cdist < old_dist) {
                        trail = plid;
                        old_dist = cdist;
                        rfQueue.push_back(std::make_pair(cIdx, plid));
                    } else if (trail < 0) {
                        trail -= 1;
                    }
                } else {
                    if (trail < 0) trail -= 1;
                }
            }
        }
    }

    // returns the number of final planes; membership: final plane id per pixel or -1
    int run(std::vector<int>& membership) {
        nodes.clear(); extracted.clear(); rfQueue.clear();
        nodes.reserve(20000);
        const int Nh = height / winH, Nw = width / winW;
        DisjointSet dset(Nh * Nw);
        ds = &dset;
        QCmp cmp{&nodes};
        MinQ minQ(cmp);
        std::vector<int> G;
        initGraph(minQ, G);
        ahCluster(minQ);
        // refineDetails
        std::vector<bool> isValid;
        findBlockMembership(isValid);
        floodFill();
        if (std::getenv("HVO_ORACLE_STATS"))
            std::fprintf(stderr, "[plane_oracle] pops %ld merges %ld nbs %ld batches %ld flood_queue %zu\n", st_pops, st_merges, st_nbs, st_batches, rfQueue.size());
        std::vector<int> old;
        extracted.swap(old);
        MinQ minQ2(cmp);
        for (int i = 0; i < (int)old.size(); ++i)
            if (isValid[i]) minQ2.push(old[i]);
        ahCluster(minQ2);
        std::vector<int> plidmap(old.size(), -1);
        for (int i = 0; i < (int)old.size(); ++i) {
            if (!isValid[i]) continue;
            const int np_rid = ds->Find(nodes[old[i]].rid);
            for (size_t j = 0; j < extracted.size(); ++j)
                if (np_rid == nodes[extracted[j]].rid) { plidmap[i] = (int)j; break; }
        }
        membership.assign((size_t)width * height, -1);
        for (size_t i = 0; i < membership.size(); ++i) {
            const int plid = membershipImg[i];
            if (plid >= 0 && plidmap[plid] >= 0) membership[i] = plidmap[plid];
        }
        ds = nullptr;
        return (int)extracted.size();
    }
};

}  // namespace peaco

extern "C" {

// Per 10x10 block statistics of the initial graph (AHCPlaneFitter.hpp:786-826): out9 per block =
// {valid(0/1), N, cx, cy, cz, nx, ny, nz, mse}; valid means "pushed to the queue" (mse < T_mse(INIT) && !nouse).
void orc_plane_blocks(const uint16_t* depth, int w, int h, float factor, float fx, float fy, float cx, float cy, double* out9) {
    peaco::Fitter f;
    f.cloud = {depth, w, h, (double)factor, (double)fx, (double)fy, (double)cx, (double)cy};
    f.width = w; f.height = h;
    const int Nh = h / 10, Nw = w / 10;
    for (int i = 0; i < Nh; ++i)
        for (int j = 0; j < Nw; ++j) {
            const int id = f.makeBlock(i * Nw + j, i * 10, j * 10);
            const peaco::Seg& p = f.nodes[id];
            double* o = out9 + 9 * (size_t)(i * Nw + j);
            o[0] = (p.mse < f.params.T_mse_init(p.center[2]) && !p.nouse) ? 1 : 0;
            o[1] = p.N;
            for (int k = 0; k < 3; ++k) { o[2 + k] = p.center[k]; o[5 + k] = p.normal[k]; }
            o[8] = p.mse;
        }
}

// Full PlaneDetection::runPlaneDetection.  planes7: per plane {nx,ny,nz,cx,cy,cz,N}; membership: w*h int32.
int orc_plane_detect(const uint16_t* depth, int w, int h, float factor, float fx, float fy, float cx, float cy, double* planes7,
                     int max_planes, int32_t* membership) {
    peaco::Fitter f;
    f.cloud = {depth, w, h, (double)factor, (double)fx, (double)fy, (double)cx, (double)cy};
    f.width = w; f.height = h;
    std::vector<int> mem;
    const int n = f.run(mem);
    for (int i = 0; i < n && i < max_planes; ++i) {
        const peaco::Seg& p = f.nodes[f.extracted[i]];
        double* o = planes7 + 7 * (size_t)i;
        for (int k = 0; k < 3; ++k) { o[k] = p.normal[k]; o[3 + k] = p.center[k]; }
        o[6] = p.N;
    }
    std::memcpy(membership, mem.data(), mem.size() * sizeof(int32_t));
    return n;
}

void orc_eig33_smallest(const double* K9, double* lam, double* v3) {
    double K[3][3];
    std::memcpy(K, K9, sizeof(K));
    peaco::eig33_smallest(K, *lam, v3);
}

void orc_eig33sym(const double* K9, double* s3, double* V9) {
    double K[3][3], s[3], V[3][3];
    std::memcpy(K, K9, sizeof(K));
    peaco::eig33sym(K, s, V);
    std::memcpy(s3, s, sizeof(s));
    std::memcpy(V9, V, sizeof(V));
}

}  // extern "C"
