// TEST INFRASTRUCTURE — CPU oracle for the ORB extractor (see orb_oracle.hpp for the citation map).
#include "orb_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <utility>

#include "cvprims.hpp"

namespace orbo {

static const int8_t kPattern[1024] = {
#include "../include/hvo_orb_pattern.inc"
};

static const int kPatch = 31, kHalfPatch = 15, kEdge = 19;

// ---- ctor tables: ORBextractor.cc:408-468 ------------------------------------------------------------
Extractor::Extractor(const Params& p) : p_(p) {
    const int n = p.nlevels;
    const double sfd = (double)p.scale_factor;  // the member is a double initialised from a float (ORBextractor.h:98)
    sf_.assign(n, 1.f);
    isf_.assign(n, 1.f);
    for (int i = 1; i < n; ++i) sf_[i] = (float)((double)sf_[i - 1] * sfd);
    for (int i = 0; i < n; ++i) isf_[i] = 1.0f / sf_[i];
    nfeat_.assign(n, 0);
    const float factor = (float)(1.0 / sfd);
    float want = (float)p.nfeatures * (1.f - factor) / (1.f - (float)std::pow((double)factor, (double)n));
    int sum = 0;
    for (int l = 0; l < n - 1; ++l) {
        nfeat_[l] = cvp::cv_round(want);
        sum += nfeat_[l];
        want *= factor;
    }
    nfeat_[n - 1] = std::max(p.nfeatures - sum, 0);

    umax_.assign(kHalfPatch + 1, 0);
    const int vmax = cvp::cv_floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    const int vmin = cvp::cv_ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) umax_[v] = cvp::cv_round(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (umax_[v0] == umax_[v0 + 1]) ++v0;
        umax_[v] = v0;
        ++v0;
    }
    lv_.resize(n);
}

// ---- pyramid: ORBextractor.cc:1105-1130 (border skipped: never read on the RGB-D path) ---------------
void Extractor::pyramid(const uint8_t* gray, int w, int h, size_t stride) {
    for (int l = 0; l < p_.nlevels; ++l) {
        Level& L = lv_[l];
        L.w = cvp::cv_round((float)w * isf_[l]);
        L.h = cvp::cv_round((float)h * isf_[l]);
        L.img.assign((size_t)L.w * L.h, 0);
        L.blurred.clear();
        L.cand.clear();
        L.kps.clear();
        if (l == 0) {
            for (int y = 0; y < h; ++y) std::copy(gray + y * stride, gray + y * stride + w, &L.img[(size_t)y * w]);
        } else {
            const Level& P = lv_[l - 1];
            cvp::resize_linear_u8(P.img.data(), P.w, P.h, P.w, L.img.data(), L.w, L.h, L.w);
        }
    }
}

// ---- quadtree distribution: ORBextractor.cc:479-761 --------------------------------------------------
namespace {
struct QNode {
    int x0, x1, y0, y1;
    std::vector<int> keys;  // indices into the input list, input order preserved
    bool leaf = false;
    int prev = -1, next = -1;
};
struct QList {  // std::list<ExtractorNode> emulation with creation-ordered ids
    std::vector<QNode> n;
    int head = -1, tail = -1, count = 0;
    int push_front(QNode&& q) {
        int id = (int)n.size();
        n.push_back(std::move(q));
        n[id].prev = -1;
        n[id].next = head;
        if (head >= 0) n[head].prev = id; else tail = id;
        head = id;
        ++count;
        return id;
    }
    int push_back(QNode&& q) {
        int id = (int)n.size();
        n.push_back(std::move(q));
        n[id].next = -1;
        n[id].prev = tail;
        if (tail >= 0) n[tail].next = id; else head = id;
        tail = id;
        ++count;
        return id;
    }
    void erase(int id) {
        int p = n[id].prev, q = n[id].next;
        if (p >= 0) n[p].next = q; else head = q;
        if (q >= 0) n[q].prev = p; else tail = p;
        --count;
        std::vector<int>().swap(n[id].keys);
    }
};

// Split node `id` at (ceil(w/2), ceil(h/2)); push non-empty children to the list front in the order
// TL, TR, BL, BR; report children with >1 key as expandable.  (DivideNode :479-535 + callers)
void split_front(QList& L, int id, const std::vector<Candidate>& in, std::vector<std::pair<int, int>>& expandable,
                 int* n_expand) {
    const int x0 = L.n[id].x0, x1 = L.n[id].x1, y0 = L.n[id].y0, y1 = L.n[id].y1;
    const int hx = (int)std::ceil((float)(x1 - x0) / 2), hy = (int)std::ceil((float)(y1 - y0) / 2);
    QNode c[4];
    c[0].x0 = x0;      c[0].x1 = x0 + hx; c[0].y0 = y0;      c[0].y1 = y0 + hy;
    c[1].x0 = x0 + hx; c[1].x1 = x1;      c[1].y0 = y0;      c[1].y1 = y0 + hy;
    c[2].x0 = x0;      c[2].x1 = x0 + hx; c[2].y0 = y0 + hy; c[2].y1 = y1;
    c[3].x0 = x0 + hx; c[3].x1 = x1;      c[3].y0 = y0 + hy; c[3].y1 = y1;
    const float xs = (float)(x0 + hx), ys = (float)(y0 + hy);
    for (int k : L.n[id].keys) {
        const bool left = in[k].x < xs, top = in[k].y < ys;
        c[(left ? 0 : 1) + (top ? 0 : 2)].keys.push_back(k);
    }
    for (int q = 0; q < 4; ++q) {
        const int sz = (int)c[q].keys.size();
        if (sz == 0) continue;
        c[q].leaf = (sz == 1);
        int cid = L.push_front(std::move(c[q]));
        if (sz > 1) {
            if (n_expand) ++*n_expand;
            expandable.emplace_back(sz, cid);
        }
    }
}
}  // namespace

std::vector<Candidate> Extractor::distribute(const std::vector<Candidate>& in, int minX, int maxX, int minY,
                                             int maxY, int N) {
    std::vector<Candidate> out;
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1) return out;  // the reference divides by zero here; not reachable for landscape frames
    const float hX = (float)(maxX - minX) / nIni;
    QList L;
    L.n.reserve(in.size() * 4 + 16);
    std::vector<int> roots(nIni);
    for (int i = 0; i < nIni; ++i) {
        QNode r;
        r.x0 = (int)(hX * (float)i);
        r.x1 = (int)(hX * (float)(i + 1));
        r.y0 = 0;
        r.y1 = maxY - minY;
        roots[i] = L.push_back(std::move(r));
    }
    for (int k = 0; k < (int)in.size(); ++k) L.n[roots[(size_t)(in[k].x / hX)]].keys.push_back(k);
    for (int it = L.head; it != -1;) {
        int nx = L.n[it].next;
        if (L.n[it].keys.size() == 1) L.n[it].leaf = true;
        else if (L.n[it].keys.empty()) L.erase(it);
        it = nx;
    }
    bool finish = false;
    std::vector<std::pair<int, int>> expandable;  // (key count, node id); id grows with creation time
    while (!finish) {
        int prev_size = L.count, n_expand = 0;
        expandable.clear();
        for (int it = L.head; it != -1;) {
            int nx = L.n[it].next;
            if (!L.n[it].leaf) {
                split_front(L, it, in, expandable, &n_expand);
                L.erase(it);
            }
            it = nx;
        }
        if (L.count >= N || L.count == prev_size) {
            finish = true;
        } else if (L.count + n_expand * 3 > N) {
            while (!finish) {
                prev_size = L.count;
                std::vector<std::pair<int, int>> order = expandable;
                expandable.clear();
                std::sort(order.begin(), order.end());  // ascending (size, creation id)
                for (int j = (int)order.size() - 1; j >= 0; --j) {
                    split_front(L, order[j].second, in, expandable, nullptr);
                    L.erase(order[j].second);
                    if (L.count >= N) break;
                }
                if (L.count >= N || L.count == prev_size) finish = true;
            }
        }
    }
    out.reserve(L.count);
    for (int it = L.head; it != -1; it = L.n[it].next) {
        const std::vector<int>& ks = L.n[it].keys;
        int best = ks[0];
        for (size_t k = 1; k < ks.size(); ++k)
            if (in[ks[k]].response > in[best].response) best = ks[k];
        out.push_back(in[best]);
    }
    return out;
}

// ---- per-cell FAST + quadtree + orientation: ORBextractor.cc:763-851, :75-102 -----------------------
void Extractor::detect() {
    const float W = 30;
    std::vector<cvp::FastKp> cell;
    for (int l = 0; l < p_.nlevels; ++l) {
        Level& L = lv_[l];
        const int minBX = kEdge - 3, minBY = minBX, maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
        const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        if (nCols < 1 || nRows < 1) continue;  // reference: division by zero (levels narrower than 62 px)
        const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
        for (int i = 0; i < nRows; ++i) {
            const float iniY = (float)(minBY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBY - 3) continue;
            if (maxY > maxBY) maxY = (float)maxBY;
            for (int j = 0; j < nCols; ++j) {
                const float iniX = (float)(minBX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBX - 6) continue;
                if (maxX > maxBX) maxX = (float)maxBX;
                const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
                const uint8_t* roi = &L.img[(size_t)y0 * L.w + x0];
                cvp::fast9_nms(roi, cw, ch, L.w, p_.ini_th, cell);
                if (cell.empty()) cvp::fast9_nms(roi, cw, ch, L.w, p_.min_th, cell);
                for (const cvp::FastKp& k : cell)
                    L.cand.push_back({(float)(k.x + j * wCell), (float)(k.y + i * hCell), (float)k.score});
            }
        }
        std::vector<Candidate> kept = distribute(L.cand, minBX, maxBX, minBY, maxBY, nfeat_[l]);
        const int patch = (int)(kPatch * sf_[l]);
        for (const Candidate& c : kept) {
            KeyPoint kp;
            kp.x = c.x + minBX;
            kp.y = c.y + minBY;
            kp.size = (float)patch;
            kp.response = c.response;
            kp.octave = l;
            kp.class_id = -1;
            // IC_Angle on the unblurred level
            const int cx = cvp::cv_round(kp.x), cy = cvp::cv_round(kp.y);
            const uint8_t* ctr = &L.img[(size_t)cy * L.w + cx];
            int m01 = 0, m10 = 0;
            for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * ctr[u];
            for (int v = 1; v <= kHalfPatch; ++v) {
                int vsum = 0;
                const int d = umax_[v];
                for (int u = -d; u <= d; ++u) {
                    const int a = ctr[u + v * L.w], b = ctr[u - v * L.w];
                    vsum += a - b;
                    m10 += u * (a + b);
                }
                m01 += v * vsum;
            }
            kp.angle = cvp::fast_atan2_deg((float)m01, (float)m10);
            L.kps.push_back(kp);
        }
    }
}

// ---- operator(): ORBextractor.cc:1041-1103, descriptor :106-144 -------------------------------------
void Extractor::extract(const uint8_t* gray, int w, int h, size_t stride, std::vector<KeyPoint>& kps,
                        std::vector<uint8_t>& desc) {
    kps.clear();
    desc.clear();
    if (!gray || w <= 0 || h <= 0) return;
    pyramid(gray, w, h, stride);
    detect();
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    for (int l = 0; l < p_.nlevels; ++l) {
        Level& L = lv_[l];
        if (L.kps.empty()) continue;
        L.blurred.resize((size_t)L.w * L.h);
        cvp::gaussian_blur7_s2(L.img.data(), L.w, L.h, L.w, L.blurred.data(), L.w);
        for (const KeyPoint& k0 : L.kps) {
            const float ang = k0.angle * factorPI;
            const float a = (float)std::cos((double)ang), b = (float)std::sin((double)ang);
            const uint8_t* ctr = &L.blurred[(size_t)cvp::cv_round(k0.y) * L.w + cvp::cv_round(k0.x)];
            const int8_t* pt = kPattern;
            for (int i = 0; i < 32; ++i, pt += 32) {
                int val = 0;
                for (int t = 0; t < 8; ++t) {
                    const float ax = pt[4 * t], ay = pt[4 * t + 1], bx = pt[4 * t + 2], by = pt[4 * t + 3];
                    const float r0 = ax * b, r0b = ay * a, c0 = ax * a, c0b = ay * b;
                    const float r1 = bx * b, r1b = by * a, c1 = bx * a, c1b = by * b;
                    const int t0 = ctr[cvp::cv_round(r0 + r0b) * L.w + cvp::cv_round(c0 - c0b)];
                    const int t1 = ctr[cvp::cv_round(r1 + r1b) * L.w + cvp::cv_round(c1 - c1b)];
                    val |= (t0 < t1) << t;
                }
                desc.push_back((uint8_t)val);
            }
            KeyPoint k = k0;
            if (l != 0) { k.x *= sf_[l]; k.y *= sf_[l]; }
            kps.push_back(k);
        }
    }
}

}  // namespace orbo
