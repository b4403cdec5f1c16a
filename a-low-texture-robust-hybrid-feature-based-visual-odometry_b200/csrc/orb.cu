// ORB extractor kernels for sm_100a (B200).  From-scratch design, batched over frames:
//
//   k_resize      x(nlevels-1)  fixed-point bilinear pyramid, bit-exact with cv::resize INTER_LINEAR 8U
//                               (reference ORBextractor.cc:1105-1130)
//   k_fast_strips x1            one CTA per row of reference FAST cells: FAST-9/16 strength, cell-local NMS,
//                               per-cell threshold fallback (ORBextractor.cc:763-826) - no score map in HBM
//   k_octree      x1            one CTA per (frame, level): DistributeOctTree (ORBextractor.cc:537-761)
//                               with parallel key partitioning and the std::list order emulated exactly
//   k_describe    x1            one warp per keypoint: IC_Angle (.cc:75-102), on-the-fly 7x7 Q8 Gaussian
//                               of the 37x37 patch (.cc:1083-1084), steered rBRIEF-256 (.cc:106-144),
//                               KeyPoint assembly (.cc:835-846,1093-1099), RGB-D depth lookup
//                               (Frame.cc:1940-1961) - no blurred pyramid in HBM
//
// Float semantics are pinned with explicit round-to-nearest intrinsics (no FMA contraction) so that
// angles, sample coordinates and scaled keypoint positions are bit-identical to the CPU oracle.
#include "orb.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace hvo {

// ------------------------------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------------------------------
__constant__ int8_t c_pattern[1024] = {
#include "../../include/hvo_orb_pattern.inc"
};
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
static const int h_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

static const int kEdge = 19;
static const int kFastMaxZh = 64;    // hCell = ceil(height / floor(height/30)) <= 60
static const int kFastMaxZone = 672; // widest run of cells one strip CTA takes (px)

__device__ __forceinline__ const uint8_t* level_ptr(const OrbGeom& g, const ImgSrc& s, int l, int f, int& pitch) {
    if (l == 0) {
        pitch = s.l0_pitch;
        return s.l0 + (long long)f * s.l0_frame;
    }
    pitch = g.lv[l].pitch;
    return s.pyr + (long long)f * g.pyr_frame_bytes + g.lv[l].img_off;
}

// ------------------------------------------------------------------------------------------------------
// K1: bilinear resize.  One CTA = 128 x 32 destination pixels.  Pass H interpolates every source row the
// tile needs once (thread <-> destination column, coefficients in registers) into shared memory as
// (S0*w0 + S1*w1) >> 4 (15 bits); pass V blends two such rows per destination row and stores 4 px / thread.
// ------------------------------------------------------------------------------------------------------
static const int kRsW = 128, kRsH = 32, kRsMaxSrcRows = 72;

__global__ void __launch_bounds__(256) k_resize(const uint8_t* __restrict__ src, int spitch, long long sframe, int sw,
                                                uint8_t* __restrict__ dst, int dpitch, long long dframe, int dw, int dh,
                                                const int2* __restrict__ xtab, const int4* __restrict__ ytab) {
    __shared__ __align__(8) uint16_t hrow[kRsMaxSrcRows * kRsW];
    const int tid = threadIdx.x, f = blockIdx.z;
    const int tx = blockIdx.x * kRsW, ty = blockIdx.y * kRsH;
    const int ylast = min(ty + kRsH, dh) - 1;
    const int s_base = __ldg(&ytab[ty]).x;
    const int ns = min(__ldg(&ytab[ylast]).y - s_base + 1, kRsMaxSrcRows);
    const uint8_t* S = src + (long long)f * sframe;
    {   // pass H
        const int col = tid & (kRsW - 1), half = tid >> 7;
        const int x = tx + col;
        if (x < dw) {
            const int2 xt = __ldg(&xtab[x]);
            const int sx = xt.x, s1 = min(sx + 1, sw - 1);
            const int w0 = (int)(short)(xt.y & 0xffff), w1 = xt.y >> 16;
            for (int r = half; r < ns; r += 2) {
                const uint8_t* row = S + (long long)(s_base + r) * spitch;
                hrow[r * kRsW + col] = (uint16_t)((__ldg(row + sx) * w0 + __ldg(row + s1) * w1) >> 4);
            }
        }
    }
    __syncthreads();
    {   // pass V
        const int cg = tid & 31, rs = tid >> 5;
        const int x0 = tx + 4 * cg;
        if (x0 < dw) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int y = ty + rs * 4 + j;
                if (y < dh) {
                    const int4 yt = __ldg(&ytab[y]);
                    const uint2 a = *reinterpret_cast<const uint2*>(&hrow[(yt.x - s_base) * kRsW + 4 * cg]);
                    const uint2 b = *reinterpret_cast<const uint2*>(&hrow[(yt.y - s_base) * kRsW + 4 * cg]);
                    const int h0[4] = {(int)(a.x & 0xffff), (int)(a.x >> 16), (int)(a.y & 0xffff), (int)(a.y >> 16)};
                    const int h1[4] = {(int)(b.x & 0xffff), (int)(b.x >> 16), (int)(b.y & 0xffff), (int)(b.y >> 16)};
                    uint32_t packed = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        int v = (((yt.z * h0[i]) >> 16) + ((yt.w * h1[i]) >> 16) + 2) >> 2;
                        v = min(255, max(0, v));
                        packed |= (uint32_t)v << (8 * i);
                    }
                    uint8_t* D = dst + (long long)f * dframe + (long long)y * dpitch + x0;
                    if (x0 + 3 < dw) *reinterpret_cast<uint32_t*>(D) = packed;  // dpitch and x0 are multiples of 4
                    else for (int i = 0; x0 + i < dw; ++i) D[i] = (uint8_t)(packed >> (8 * i));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K2: FAST per cell
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool has_run9(uint32_t m16) {
    uint32_t m = m16 | (m16 << 16);
    uint32_t a = m & (m >> 1);
    a &= a >> 2;
    a &= a >> 4;
    a &= m >> 8;
    return (a & 0xffffu) != 0;
}

// Threshold-independent FAST-9/16 strength S = max(A, B) - 1 (see oracle/cvprims.hpp fast_strength):
//   A = max over the 16 arcs of min(v - ring), B = max over arcs of min(ring - v).
__device__ __forceinline__ int fast_strength(const int (&d)[16]) {
    // min / max over every 9-arc as two levels of 3-input min/max (VIMNMX3 on sm_100a)
    int lo3[16], hi3[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        lo3[i] = min(min(d[i], d[(i + 1) & 15]), d[(i + 2) & 15]);
        hi3[i] = max(max(d[i], d[(i + 1) & 15]), d[(i + 2) & 15]);
    }
    int A = -256, B = 256;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        A = max(A, min(min(lo3[i], lo3[(i + 3) & 15]), lo3[(i + 6) & 15]));
        B = min(B, max(max(hi3[i], hi3[(i + 3) & 15]), hi3[(i + 6) & 15]));
    }
    return max(A, -B) - 1;
}

#define HVO_RING(p, st, k)                                                                                   \
    ((k) == 0 ? (p)[3 * (st)] : (k) == 1 ? (p)[3 * (st) + 1] : (k) == 2 ? (p)[2 * (st) + 2] : (k) == 3 ? (p)[(st) + 3] \
     : (k) == 4 ? (p)[3] : (k) == 5 ? (p)[-(st) + 3] : (k) == 6 ? (p)[-2 * (st) + 2] : (k) == 7 ? (p)[-3 * (st) + 1]   \
     : (k) == 8 ? (p)[-3 * (st)] : (k) == 9 ? (p)[-3 * (st) - 1] : (k) == 10 ? (p)[-2 * (st) - 2]                      \
     : (k) == 11 ? (p)[-(st) - 3] : (k) == 12 ? (p)[-3] : (k) == 13 ? (p)[(st) - 3] : (k) == 14 ? (p)[2 * (st) - 2]    \
                                                                                                : (p)[3 * (st) - 1])

// One CTA per STRIP = one row of reference FAST cells of one level of one frame (the zones of a cell row tile the level
// without gaps: cell ROI = zone + 3-px ring, ORBextractor.cc:787-803).  The FAST strength of a pixel depends on the image
// only; what is cell-local is (a) non-max suppression (neighbours outside the cell's zone count as 0) and (b) the
// threshold fallback (ORBextractor.cc:807-814).  So the strip is processed as one image band:
//   fill     the band (zone rows + 3-px ring) goes to shared memory as one bulk copy per row (cp.async.bulk + mbarrier:
//            the copy engine moves the bytes, no LDG/STS issue slots), while the threads clear the score map
//   pass 1   quick reject, 4 pixels per thread on packed bytes; survivors go to a warp-private queue
//   pass 2   exact threshold-independent strength for queued survivors, 32 at a time (dense warps), no CTA barrier
//   pass 3   NMS over the score map (word-wise skip of empty groups), cell-aware; maxima -> list (reuses the band's memory)
//   emit     per-cell ini/min decision, one global atomic per strip
static const int kFastThreads = 256, kFastWarps = kFastThreads / 32;
static const int kFastQueue = 256;   // per-warp survivor queue (entries; >= 31 + 128)
static const int kFastMaxCellsPerStrip = 64;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Packed 4-pixel FAST ring compare: the 4 ring bytes seen by the 4 pixels of word g at horizontal offset dx (-3..3).
__device__ __forceinline__ uint32_t ring4(const uint32_t* row, int g, int dx) {
    switch (dx) {
        case 0: return row[g];
        case 1: return __byte_perm(row[g], row[g + 1], 0x4321);
        case 2: return __byte_perm(row[g], row[g + 1], 0x5432);
        case 3: return __byte_perm(row[g], row[g + 1], 0x6543);
        case -1: return __byte_perm(row[g - 1], row[g], 0x6543);
        case -2: return __byte_perm(row[g - 1], row[g], 0x5432);
        default: return __byte_perm(row[g - 1], row[g], 0x4321);
    }
}

__global__ void __launch_bounds__(kFastThreads) k_fast_strips(const __grid_constant__ OrbGeom g, ImgSrc src,
                                                              const StripDesc* __restrict__ strips, uint32_t* __restrict__ cand,
                                                              int* __restrict__ ncand, int ini_th, int min_th) {
    extern __shared__ __align__(16) unsigned char fs_smem[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_queue[kFastWarps][kFastQueue];
    __shared__ int s_nini[kFastMaxCellsPerStrip], s_nmin[kFastMaxCellsPerStrip];
    __shared__ int s_nlist, s_base, s_slot, s_total;

    const StripDesc c = strips[blockIdx.x];
    const int f = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const LevelGeom& L = g.lv[c.level];
    int pitch;
    const uint8_t* img = level_ptr(g, src, c.level, f, pitch);
    const int zh = c.zh, zw = c.zw, th = zh + 6;
    const int low_th = min(ini_th, min_th);

    // band geometry: tile column 0 = image column xal (16-byte aligned), zone starts at tile byte zb0
    const int xal = (c.x0 - 3) & ~15;
    const int zb0 = c.x0 - xal;                       // >= 3
    const int ts = c.tstride;                         // bytes per band row, multiple of 16
    const int tsw = ts >> 2;
    unsigned char* tile_b = fs_smem + 16;             // word -1 of row 0 must be addressable (content unused)
    uint32_t* tile = reinterpret_cast<uint32_t*>(tile_b);
    unsigned char* score = tile_b + c.score_off;      // (zh + 2) rows of ts bytes; row 0 and row zh+1 stay 0
    uint32_t* list = tile;                            // NMS maxima, written after the band is dead

    // ---- fill ----
    const bool bulk = ((pitch & 15) == 0) && ((reinterpret_cast<uintptr_t>(img) & 15) == 0);
    if (tid == 0) {
        s_nlist = 0; s_slot = 0;
        if (bulk) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&s_bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    if (tid < kFastMaxCellsPerStrip) { s_nini[tid] = 0; s_nmin[tid] = 0; }
    __syncthreads();
    if (bulk) {
        if (warp == 0) {
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&s_bar)), "r"(th * ts) : "memory");
            __syncwarp();
            for (int r = lane; r < th; r += 32) {
                const uint8_t* gp = img + (long long)(c.y0 - 3 + r) * pitch + xal;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_addr(tile_b + r * ts)),
                             "l"(gp), "r"(ts), "r"(smem_addr(&s_bar))
                             : "memory");
            }
        }
    } else {
        const int lim = L.w - 1 - xal;
        for (int i = tid; i < th * tsw; i += kFastThreads) {
            const int r = i / tsw, wi = i - r * tsw;
            const uint8_t* p = img + (long long)(c.y0 - 3 + r) * pitch + xal;
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) w |= (uint32_t)__ldg(p + min(4 * wi + j, lim)) << (8 * j);
            tile[r * tsw + wi] = w;
        }
    }
    {   // clear the score map while the copy engine works
        uint4* sc4 = reinterpret_cast<uint4*>(score);
        const int n16 = ((zh + 2) * ts) >> 4;
        for (int i = tid; i < n16; i += kFastThreads) sc4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (bulk) {
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_addr(&s_bar)) : "memory");
    }
    __syncthreads();

    // ---- pass 1 + 2 ----
    // A 9-arc always contains ring point 0 or 8 and ring point 4 or 12, so a corner needs |v - ring| > t on one point of each
    // pair.  VABSDIFF4 is native; "byte > t" is the carry trick ((x & 0x7f) + (127 - t)) | x.
    // Survivors (byte flags in bit 7 of `flags`) are compacted into a warp-private queue with four ballots (no shuffles, no loop
    // over the flags) and processed 32 at a time, so the expensive part runs on full warps without a CTA barrier in between.
    const int g0 = zb0 >> 2, g1 = (zb0 + zw - 1) >> 2;
    const bool use_quick = low_th <= 127;
    const uint32_t K = (uint32_t)(127 - min(low_th, 127)) * 0x01010101u;
    uint32_t* q = s_queue[warp];
    int qh = 0, qn = 0;                                  // warp-uniform queue head / tail (monotonic)
    const unsigned lt = (1u << lane) - 1u;
    auto enqueue = [&](uint32_t flags, uint32_t base_pos) -> bool {   // base_pos = zy << 16 | band column of byte 0
        const unsigned m0 = __ballot_sync(0xffffffffu, flags & 0x80u), m1 = __ballot_sync(0xffffffffu, flags & 0x8000u);
        const unsigned m2 = __ballot_sync(0xffffffffu, flags & 0x800000u), m3 = __ballot_sync(0xffffffffu, flags & 0x80000000u);
        const int c0 = __popc(m0), c1 = __popc(m1), c2 = __popc(m2), total = c0 + c1 + c2 + __popc(m3);
        if (total == 0) return false;
        if (flags & 0x80u) q[(qn + __popc(m0 & lt)) & (kFastQueue - 1)] = base_pos;
        if (flags & 0x8000u) q[(qn + c0 + __popc(m1 & lt)) & (kFastQueue - 1)] = base_pos + 1;
        if (flags & 0x800000u) q[(qn + c0 + c1 + __popc(m2 & lt)) & (kFastQueue - 1)] = base_pos + 2;
        if (flags & 0x80000000u) q[(qn + c0 + c1 + c2 + __popc(m3 & lt)) & (kFastQueue - 1)] = base_pos + 3;
        qn += total;
        __syncwarp();
        return true;
    };
    auto strength_at = [&](uint32_t pos) {
        const int py = pos >> 16, px = pos & 0xffff;
        const uint8_t* p = tile_b + (py + 3) * ts + px;
        const int v = *p;
        int d[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) d[k] = v - (int)HVO_RING(p, ts, k);
        const int sc = fast_strength(d);
        if (sc >= low_th) score[(py + 1) * ts + px] = (uint8_t)sc;   // in [low_th, 254]; non-corners stay 0
    };
    for (int zy = warp; zy < zh; zy += kFastWarps) {
        const uint32_t* r0 = tile + (zy + 3) * tsw;
        for (int gb = g0; gb <= g1; gb += 32) {
            const int gi = gb + lane;
            uint32_t maybe = 0;
            if (gi <= g1) {
                const uint32_t v4 = r0[gi];
                const int b0 = 4 * gi;
                maybe = 0x80808080u;
                if (b0 < zb0) maybe &= 0xffffffffu << (8 * (zb0 - b0));
                if (b0 + 3 > zb0 + zw - 1) maybe &= 0xffffffffu >> (8 * (b0 + 3 - (zb0 + zw - 1)));
                if (use_quick) {
                    const uint32_t a0 = __vabsdiffu4(v4, (r0 + 3 * tsw)[gi]), a8 = __vabsdiffu4(v4, (r0 - 3 * tsw)[gi]);
                    const uint32_t a4 = __vabsdiffu4(v4, ring4(r0, gi, 3)), a12 = __vabsdiffu4(v4, ring4(r0, gi, -3));
                    const uint32_t m08 = ((a0 & 0x7f7f7f7fu) + K) | ((a8 & 0x7f7f7f7fu) + K) | a0 | a8;
                    const uint32_t m412 = ((a4 & 0x7f7f7f7fu) + K) | ((a12 & 0x7f7f7f7fu) + K) | a4 | a12;
                    maybe &= m08 & m412;
                }
            }
            if (!enqueue(maybe, ((uint32_t)zy << 16) | (uint32_t)(4 * gi))) continue;
            while (qn - qh >= 32) {
                strength_at(q[(qh + lane) & (kFastQueue - 1)]);
                qh += 32;
            }
            __syncwarp();
        }
    }
    if (lane < qn - qh) strength_at(q[(qh + lane) & (kFastQueue - 1)]);
    __syncthreads();

    // ---- pass 3: cell-local non-max suppression (strict '>' on the 8 neighbours, outside the cell's zone counts as 0).  Same
    //      skeleton: pixels with a score are compacted per warp and tested 32 at a time, branch-free (the guard rows / columns of
    //      the score map are zero, so all 8 neighbours can always be read) ----
    const int wcell = c.wcell;
    const uint32_t cmagic = (1u << 20) / (uint32_t)wcell + 1;   // zx / wcell for zx < 2^12
    const uint32_t* score_w = reinterpret_cast<const uint32_t*>(score);
    qh = qn = 0;
    auto nms_at = [&](uint32_t pos) {
        const int zy = pos >> 16, px = pos & 0xffff, zx = px - zb0;
        int cj = (int)(((uint32_t)zx * cmagic) >> 20);
        if (cj * wcell > zx) --cj;
        const int cx = zx - cj * wcell;
        const uint8_t* sp = score + (zy + 1) * ts + px;
        const int s = sp[0];
        const int up = sp[-ts], dn = sp[ts];
        const int l = max(max((int)sp[-ts - 1], (int)sp[-1]), (int)sp[ts - 1]);
        const int r = max(max((int)sp[-ts + 1], (int)sp[1]), (int)sp[ts + 1]);
        int m = max(up, dn);
        m = max(m, cx > 0 ? l : 0);
        m = max(m, (cx < wcell - 1 && zx < zw - 1) ? r : 0);
        if (s > m) {
            if (s >= ini_th) atomicAdd(&s_nini[cj], 1);
            if (s >= min_th) atomicAdd(&s_nmin[cj], 1);
            list[atomicAdd(&s_nlist, 1)] = (uint32_t)(c.x0 + zx) | ((uint32_t)(c.y0 + zy) << 12) | ((uint32_t)s << 24);
        }
    };
    for (int zy = warp; zy < zh; zy += kFastWarps) {
        const uint32_t* srow = score_w + (zy + 1) * tsw;
        for (int gb = g0; gb <= g1; gb += 32) {
            const int gi = gb + lane;
            uint32_t w = gi <= g1 ? srow[gi] : 0u;
            // byte != 0 -> bit 7 of that byte
            w = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;
            if (!enqueue(w, ((uint32_t)zy << 16) | (uint32_t)(4 * gi))) continue;
            while (qn - qh >= 32) {
                nms_at(q[(qh + lane) & (kFastQueue - 1)]);
                qh += 32;
            }
            __syncwarp();
        }
    }
    if (lane < qn - qh) nms_at(q[(qh + lane) & (kFastQueue - 1)]);
    __syncthreads();

    // ---- emit: a cell that has a corner at ini_th keeps those, else the ones at min_th ----
    if (tid == 0) {
        int n = 0;
        for (int j = 0; j < c.ncells; ++j) n += s_nini[j] > 0 ? s_nini[j] : s_nmin[j];
        s_total = n;
        s_base = n > 0 ? atomicAdd(&ncand[f * g.nlevels + c.level], n) : 0;
    }
    __syncthreads();
    if (s_total == 0) return;
    uint32_t* out = cand + (long long)f * g.cand_total + L.cand_off + s_base;
    const int nlist = s_nlist;
    for (int i = tid; i < nlist; i += kFastThreads) {
        const uint32_t e = list[i];
        const int zx = (int)(e & 0xfff) - c.x0, s = (int)(e >> 24);
        int cj = (int)(((uint32_t)zx * cmagic) >> 20);
        if (cj * wcell > zx) --cj;
        const bool keep = s_nini[cj] > 0 ? s >= ini_th : s >= min_th;
        if (keep) {
            const int slot = atomicAdd(&s_slot, 1);
            if (s_base + slot < L.cand_cap) out[slot] = e;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K3: quadtree distribution, one CTA per (frame, level)
// ------------------------------------------------------------------------------------------------------
// The reference keeps a std::list of nodes; children are pushed to the FRONT in the order TL,TR,BL,BR and the
// parent is erased.  Here the list is an array in list order (front = index 0).  Per round:
//   * all threads histogram the keys of every splittable node into its 4 quadrants (shared atomics),
//   * thread 0 replays the list bookkeeping (<= quota+3 live nodes) and lays out the next list,
//   * all threads move their keys to the new node positions.
// Phase 2 ("expand the biggest first", ORBextractor.cc:671-735) sorts the splittable nodes by
// (key count, creation order) with a parallel rank sort; ties: later created node first (the oracle's
// documented stand-in for the reference's pointer comparison).
struct OctShared {
    int M, mode, nE, cur;
};
enum { OCT_PHASE1 = 0, OCT_PHASE2 = 1, OCT_DONE = 2 };
enum { NODE_LEAF = 1 };

__device__ __forceinline__ int oct_quadrant(int kx, int ky, const short* r) {
    const int hx = (r[2] - r[0] + 1) >> 1, hy = (r[3] - r[1] + 1) >> 1;  // ceil(w/2), ceil(h/2)
    return (kx < r[0] + hx ? 0 : 1) + (ky < r[1] + hy ? 0 : 2);
}
__device__ __forceinline__ void oct_child_rect(const short* r, int q, short* o) {
    const int hx = (r[2] - r[0] + 1) >> 1, hy = (r[3] - r[1] + 1) >> 1;
    o[0] = (q & 1) ? r[0] + hx : r[0];
    o[2] = (q & 1) ? r[2] : r[0] + hx;
    o[1] = (q & 2) ? r[1] + hy : r[1];
    o[3] = (q & 2) ? r[3] : r[1] + hy;
}

__global__ void __launch_bounds__(256) k_octree(const __grid_constant__ OrbGeom g, const uint32_t* __restrict__ cand_all,
                                                const int* __restrict__ ncand, uint16_t* __restrict__ knode_all,
                                                uint32_t* __restrict__ okp_all, int* __restrict__ on, int cap_nodes,
                                                int* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char oct_smem[];
    const int l = blockIdx.x, f = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
    const LevelGeom& L = g.lv[l];
    const int N = L.quota;
    int n = ncand[f * g.nlevels + l];
    if (n > L.cand_cap) n = L.cand_cap;
    if (n == 0 || L.nIni < 1 || L.nIni > cap_nodes) {
        if (tid == 0) on[f * g.nlevels + l] = 0;
        return;
    }
    const uint32_t* cand = cand_all + (long long)f * g.cand_total + L.cand_off;
    uint16_t* knode = knode_all + (long long)f * g.cand_total + L.cand_off;
    uint32_t* okp = okp_all + (long long)f * g.kp_total + L.kp_off;

    // shared layout (cap_nodes = quota + 8 entries each)
    const int CN = cap_nodes;
    unsigned long long* best = reinterpret_cast<unsigned long long*>(oct_smem);  // [CN]
    int* cnt4 = reinterpret_cast<int*>(best + CN);                                // [CN][4]
    int* size0 = cnt4 + 4 * CN;                                                   // [2][CN]
    int* e_size = size0 + 2 * CN;                                                 // [CN]
    int* e_pos = e_size + CN;                                                     // [CN]
    int* e_sorted = e_pos + CN;                                                   // [CN]
    short* rect0 = reinterpret_cast<short*>(e_sorted + CN);                       // [2][CN][4]
    short* childpos = rect0 + 2 * CN * 4;                                         // [CN][4]
    short* newpos = childpos + CN * 4;                                            // [CN]
    unsigned char* flags0 = reinterpret_cast<unsigned char*>(newpos + CN);        // [2][CN]
    unsigned char* divided = flags0 + 2 * CN;                                     // [CN]
    __shared__ OctShared S;

    // ---- roots (ORBextractor.cc:541-589) ----
    if (tid == 0) { S.cur = 0; S.M = L.nIni; S.mode = OCT_PHASE1; S.nE = 0; }
    for (int i = tid; i < L.nIni; i += nt) {
        short* r = rect0 + i * 4;
        r[0] = (short)(int)__fmul_rn(L.hX, (float)i);
        r[2] = (short)(int)__fmul_rn(L.hX, (float)(i + 1));
        r[1] = 0;
        r[3] = (short)(L.maxBY - L.minBY);
        size0[i] = 0;
    }
    __syncthreads();
    for (int k = tid; k < n; k += nt) {
        const uint32_t c = cand[k];
        const int kx = (int)(c & 0xfff) - L.minBX;
        int r = (int)__fdiv_rn((float)kx, L.hX);
        r = min(r, L.nIni - 1);
        knode[k] = (uint16_t)r;
        atomicAdd(&size0[r], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int pos = 0;
        for (int r = 0; r < L.nIni; ++r) {
            if (size0[r] == 0) { newpos[r] = -1; continue; }
            newpos[r] = (short)pos;
            for (int j = 0; j < 4; ++j) rect0[(CN + pos) * 4 + j] = rect0[r * 4 + j];
            size0[CN + pos] = size0[r];
            flags0[CN + pos] = size0[r] == 1 ? NODE_LEAF : 0;
            ++pos;
        }
        S.M = pos;
        S.cur = 1;
    }
    __syncthreads();
    for (int k = tid; k < n; k += nt) knode[k] = (uint16_t)newpos[knode[k]];
    __syncthreads();

    // ---- rounds ----
    while (S.mode != OCT_DONE) {
        const int cur = S.cur, M = S.M, mode = S.mode, nE = S.nE;
        short* rect = rect0 + cur * CN * 4;
        short* nrect = rect0 + (cur ^ 1) * CN * 4;
        int* size = size0 + cur * CN;
        int* nsize = size0 + (cur ^ 1) * CN;
        unsigned char* flags = flags0 + cur * CN;
        unsigned char* nflags = flags0 + (cur ^ 1) * CN;

        for (int i = tid; i < 4 * M; i += nt) cnt4[i] = 0;
        for (int i = tid; i < M; i += nt) divided[i] = 0;
        __syncthreads();
        for (int k = tid; k < n; k += nt) {
            const int p = knode[k] & 0x3fff;
            if (!(flags[p] & NODE_LEAF)) {
                const uint32_t c = cand[k];
                const int q = oct_quadrant((int)(c & 0xfff) - L.minBX, (int)((c >> 12) & 0xfff) - L.minBY, rect + p * 4);
                atomicAdd(&cnt4[p * 4 + q], 1);
                knode[k] = (uint16_t)(p | (q << 14));
            }
        }
        if (mode == OCT_PHASE2) {
            // rank sort of the splittable nodes by (size, creation order) ascending
            for (int e = tid; e < nE; e += nt) {
                const int se = e_size[e];
                int rank = 0;
                for (int j = 0; j < nE; ++j) {
                    const int sj = e_size[j];
                    rank += (sj < se || (sj == se && j < e)) ? 1 : 0;
                }
                e_sorted[rank] = e;
            }
        }
        __syncthreads();

        if (tid == 0) {
            int C = 0, newM, nExp = 0, newE = 0;
            if (mode == OCT_PHASE1) {
                // every non-leaf node splits, in list order
                int nLeaf = 0;
                for (int p = 0; p < M; ++p) {
                    if (flags[p] & NODE_LEAF) { newpos[p] = (short)nLeaf++; continue; }
                    divided[p] = 1;
                    for (int q = 0; q < 4; ++q)
                        if (cnt4[p * 4 + q] > 0) childpos[p * 4 + q] = (short)C++;
                }
                newM = C + nLeaf;
                for (int p = 0; p < M; ++p) {
                    if (flags[p] & NODE_LEAF) {
                        const int dst = C + newpos[p];
                        newpos[p] = (short)dst;
                        for (int j = 0; j < 4; ++j) nrect[dst * 4 + j] = rect[p * 4 + j];
                        nsize[dst] = size[p];
                        nflags[dst] = NODE_LEAF;
                    } else {
                        for (int q = 0; q < 4; ++q) {
                            const int cnt = cnt4[p * 4 + q];
                            if (cnt == 0) continue;
                            const int dst = C - 1 - childpos[p * 4 + q];
                            childpos[p * 4 + q] = (short)dst;
                            oct_child_rect(rect + p * 4, q, nrect + dst * 4);
                            nsize[dst] = cnt;
                            nflags[dst] = cnt == 1 ? NODE_LEAF : 0;
                            if (cnt > 1) { ++nExp; e_size[newE] = cnt; e_pos[newE] = dst; ++newE; }
                        }
                    }
                }
                if (newM >= N || newM == M) S.mode = OCT_DONE;
                else if (newM + 3 * nExp > N) S.mode = OCT_PHASE2;
            } else {
                // split the biggest splittable nodes first until the list holds >= N nodes
                int count = M, jstop = nE;  // nodes e_sorted[jstop..nE) get split
                for (int j = nE - 1; j >= 0; --j) {
                    const int p = e_pos[e_sorted[j]];
                    int nchild = 0;
                    for (int q = 0; q < 4; ++q)
                        if (cnt4[p * 4 + q] > 0) { childpos[p * 4 + q] = (short)C++; ++nchild; }
                    divided[p] = 1;
                    count += nchild - 1;
                    jstop = j;
                    if (count >= N) break;
                }
                newM = count;
                int pos = C;
                for (int p = 0; p < M; ++p) {
                    if (divided[p]) continue;
                    newpos[p] = (short)pos;
                    for (int j = 0; j < 4; ++j) nrect[pos * 4 + j] = rect[p * 4 + j];
                    nsize[pos] = size[p];
                    nflags[pos] = flags[p];
                    ++pos;
                }
                for (int j = nE - 1; j >= jstop; --j) {
                    const int p = e_pos[e_sorted[j]];
                    for (int q = 0; q < 4; ++q) {
                        const int cnt = cnt4[p * 4 + q];
                        if (cnt == 0) continue;
                        const int dst = C - 1 - childpos[p * 4 + q];
                        childpos[p * 4 + q] = (short)dst;
                        oct_child_rect(rect + p * 4, q, nrect + dst * 4);
                        nsize[dst] = cnt;
                        nflags[dst] = cnt == 1 ? NODE_LEAF : 0;
                    }
                }
                // new splittable list in creation order (= dst descending from C-1 to 0); e_* are free to overwrite now
                for (int dst = C - 1; dst >= 0; --dst)
                    if (!(nflags[dst] & NODE_LEAF)) { e_size[newE] = nsize[dst]; e_pos[newE] = dst; ++newE; }
                if (newM >= N || newM == M) S.mode = OCT_DONE;
            }
            S.M = newM;
            S.nE = newE;
            S.cur = cur ^ 1;
        }
        __syncthreads();
        for (int k = tid; k < n; k += nt) {
            const int kn = knode[k], p = kn & 0x3fff, q = kn >> 14;
            knode[k] = (uint16_t)(divided[p] ? childpos[p * 4 + q] : newpos[p]);
        }
        __syncthreads();
    }

    // ---- best key per node: max response, ties -> earliest in the reference's candidate order ----
    const int M = S.M;
    for (int i = tid; i < M; i += nt) best[i] = 0ull;
    __syncthreads();
    for (int k = tid; k < n; k += nt) {
        const uint32_t c = cand[k];
        const int x = c & 0xfff, y = (c >> 12) & 0xfff, s = c >> 24;
        const int ci = (y - L.minBY - 3) / L.hCell, cj = (x - L.minBX - 3) / L.wCell;
        const unsigned long long order = ((unsigned long long)(ci * L.nCols + cj) << 24) | ((unsigned long long)y << 12) | (unsigned long long)x;
        const unsigned long long key = ((unsigned long long)s << 48) | (0xffffffffffffull - order);
        atomicMax(&best[knode[k] & 0x3fff], key);
    }
    __syncthreads();
    for (int k = tid; k < n; k += nt) {
        const uint32_t c = cand[k];
        const int x = c & 0xfff, y = (c >> 12) & 0xfff, s = c >> 24;
        const int ci = (y - L.minBY - 3) / L.hCell, cj = (x - L.minBX - 3) / L.wCell;
        const unsigned long long order = ((unsigned long long)(ci * L.nCols + cj) << 24) | ((unsigned long long)y << 12) | (unsigned long long)x;
        const unsigned long long key = ((unsigned long long)s << 48) | (0xffffffffffffull - order);
        const int p = knode[k] & 0x3fff;
        if (best[p] == key && p < L.kp_cap) okp[p] = c;
    }
    if (tid == 0) {
        on[f * g.nlevels + l] = min(M, L.kp_cap);
        if (M > L.kp_cap) atomicExch(err, HVO_ERR_OVERFLOW);
    }
}

// ------------------------------------------------------------------------------------------------------
// K4: orientation + on-the-fly blur + rBRIEF + output assembly, one warp per keypoint
// ------------------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------------------
// K4: 7x7 sigma-2 Gaussian of every level (cv::GaussianBlur fixed point: Q8 kernel 18,34,48,56,48,34,18,
// horizontal pass exact in 16 bits, vertical pass rounded (v + 2^15) >> 16, BORDER_REFLECT_101).
// One CTA = 128 x 32 output pixels; horizontal pass with __dp4a on aligned words, vertical pass 4x4 px / thread.
// ------------------------------------------------------------------------------------------------------
static const int kBlW = 128, kBlH = 32, kBlRawStride = 35, kBlHbStride = 66;

__global__ void __launch_bounds__(256) k_blur(const __grid_constant__ OrbGeom g, ImgSrc src,
                                              const TileDesc* __restrict__ tiles, uint8_t* __restrict__ blur) {
    __shared__ uint32_t raw[(kBlH + 6) * kBlRawStride];
    __shared__ __align__(8) uint32_t hb[(kBlH + 6) * kBlHbStride];
    const TileDesc t = tiles[blockIdx.x];
    const int f = blockIdx.y, tid = threadIdx.x;
    const LevelGeom& L = g.lv[t.level];
    int pitch;
    const uint8_t* img = level_ptr(g, src, t.level, f, pitch);
    const bool aligned = ((pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(img) & 3) == 0);
    const int tx = t.tx, ty = t.ty;
    for (int idx = tid; idx < (kBlH + 6) * 34; idx += 256) {
        const int r = idx / 34, wi = idx - r * 34;
        const int y = reflect101(ty - 3 + r, L.h);
        const int x = tx - 4 + 4 * wi;
        const uint8_t* row = img + (long long)y * pitch;
        uint32_t w = 0;
        if (aligned && x >= 0 && x + 3 < L.w) {
            w = __ldg(reinterpret_cast<const uint32_t*>(row + x));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int xx = min(max(reflect101(x + j, L.w), 0), L.w - 1);
                w |= (uint32_t)__ldg(row + xx) << (8 * j);
            }
        }
        raw[r * kBlRawStride + wi] = w;
    }
    __syncthreads();
    const uint32_t K0123 = 18u | (34u << 8) | (48u << 16) | (56u << 24), K456 = 48u | (34u << 8) | (18u << 16);
    for (int idx = tid; idx < (kBlH + 6) * 32; idx += 256) {
        const int r = idx >> 5, gq = idx & 31;
        const uint32_t a = raw[r * kBlRawStride + gq], b = raw[r * kBlRawStride + gq + 1], c = raw[r * kBlRawStride + gq + 2];
        // output x = 4gq + j reads tile bytes 4gq + j + 1 .. + 7
        const uint32_t o0 = __dp4a(__byte_perm(a, b, 0x4321), K0123, __dp4a(__byte_perm(b, c, 0x4321), K456, 0u));
        const uint32_t o1 = __dp4a(__byte_perm(a, b, 0x5432), K0123, __dp4a(__byte_perm(b, c, 0x5432), K456, 0u));
        const uint32_t o2 = __dp4a(__byte_perm(a, b, 0x6543), K0123, __dp4a(__byte_perm(b, c, 0x6543), K456, 0u));
        const uint32_t o3 = __dp4a(b, K0123, __dp4a(c, K456, 0u));
        *reinterpret_cast<uint2*>(&hb[r * kBlHbStride + 2 * gq]) = make_uint2(o0 | (o1 << 16), o2 | (o3 << 16));
    }
    __syncthreads();
    {
        const int q = tid & 31, sgm = tid >> 5;  // 4 columns x 4 rows per thread
        const int x0 = tx + 4 * q;
        if (x0 < L.w) {
            uint32_t col[10][2];
#pragma unroll
            for (int r = 0; r < 10; ++r) {
                const uint2 v = *reinterpret_cast<const uint2*>(&hb[(4 * sgm + r) * kBlHbStride + 2 * q]);
                col[r][0] = v.x; col[r][1] = v.y;
            }
            uint8_t* out = blur + (long long)f * g.blur_frame_bytes + L.blur_off;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int y = ty + 4 * sgm + j;
                uint32_t packed = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t h[7];
#pragma unroll
                    for (int k = 0; k < 7; ++k) h[k] = (i & 1) ? (col[j + k][i >> 1] >> 16) : (col[j + k][i >> 1] & 0xffffu);
                    const uint32_t acc = 18u * (h[0] + h[6]) + 34u * (h[1] + h[5]) + 48u * (h[2] + h[4]) + 56u * h[3];
                    packed |= ((acc + 32768u) >> 16) << (8 * i);
                }
                if (y < L.h) {
                    uint8_t* D = out + (long long)y * L.bpitch + x0;
                    if (x0 + 3 < L.w) *reinterpret_cast<uint32_t*>(D) = packed;
                    else for (int i = 0; x0 + i < L.w; ++i) D[i] = (uint8_t)(packed >> (8 * i));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K5: orientation + steered rBRIEF + output assembly, one warp per keypoint
// ------------------------------------------------------------------------------------------------------
static const int kDescWarps = 8;

__global__ void __launch_bounds__(kDescWarps * 32) k_describe(const __grid_constant__ OrbGeom g, ImgSrc src,
                                                              const uint8_t* __restrict__ blur,
                                                              const uint32_t* __restrict__ okp_all,
                                                              const int* __restrict__ on, hvo_keypoint* __restrict__ kps,
                                                              uint8_t* __restrict__ desc, int32_t* __restrict__ counts,
                                                              const uint16_t* __restrict__ depth16, float depth_factor,
                                                              float bf, int distorted, float* __restrict__ kp_depth,
                                                              float* __restrict__ kp_uright) {
    __shared__ int8_t s_pat[1024];
    const int f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) reinterpret_cast<int*>(s_pat)[i] = reinterpret_cast<const int*>(c_pattern)[i];
    __syncthreads();

    // locate keypoint gi (level-major) of frame f
    const int gi = blockIdx.x * kDescWarps + warp;
    int total = 0, lvl = -1, idx = 0;
    for (int l = 0; l < g.nlevels; ++l) {
        const int nl = on[f * g.nlevels + l];
        if (lvl < 0 && gi < total + nl) { lvl = l; idx = gi - total; }
        total += nl;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[f] = min(total, g.out_cap);
    if (lvl < 0 || gi >= g.out_cap) return;

    const LevelGeom& L = g.lv[lvl];
    const uint32_t c = okp_all[(long long)f * g.kp_total + L.kp_off + idx];
    const int cx = c & 0xfff, cy = (c >> 12) & 0xfff, resp = c >> 24;
    int pitch;
    const uint8_t* img = level_ptr(g, src, lvl, f, pitch);

    // IC_Angle on the unblurred level: lane <-> column u = lane - 15, rows are coalesced 31-byte reads
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - 15, au = abs(u);
        const uint8_t* p = img + (long long)cy * pitch + cx + u;
#pragma unroll
        for (int v = -15; v <= 15; ++v) {
            if (au <= c_umax[v < 0 ? -v : v]) {
                const int val = __ldg(p + (long long)v * pitch);
                m10 += u * val;
                m01 += v * val;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // steered rBRIEF on the blurred level: lane i -> descriptor byte i
    const float factorPI = 0.017453292519943295769236907684886f;  // (float)(CV_PI / 180.f)
    const float ang = __fmul_rn(angle, factorPI);
    const float a = (float)cos((double)ang), b = (float)sin((double)ang);
    const uint8_t* ctr = blur + (long long)f * g.blur_frame_bytes + L.blur_off + (long long)cy * L.bpitch + cx;
    const int bp = L.bpitch;
    const int8_t* pt = s_pat + lane * 32;
    int t0v[8], t1v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const float ax = pt[4 * t], ay = pt[4 * t + 1], bx = pt[4 * t + 2], by = pt[4 * t + 3];
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(ax, b), __fmul_rn(ay, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(ax, a), __fmul_rn(ay, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(bx, b), __fmul_rn(by, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(bx, a), __fmul_rn(by, b)));
        t0v[t] = __ldg(ctr + r0 * bp + c0);
        t1v[t] = __ldg(ctr + r1 * bp + c1);
    }
    int val = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) val |= (t0v[t] < t1v[t] ? 1 : 0) << t;
    const long long o = (long long)f * g.out_cap + gi;
    desc[o * 32 + lane] = (uint8_t)val;

    if (lane == 0) {
        float x = (float)cx, y = (float)cy;
        if (lvl != 0) { x = __fmul_rn(x, L.scale); y = __fmul_rn(y, L.scale); }
        hvo_keypoint k;
        k.x = x; k.y = y; k.size = L.kp_size; k.angle = angle; k.response = (float)resp; k.octave = lvl; k.class_id = -1;
        kps[o] = k;
        if (depth16 != nullptr) {
            const int u = min((int)x, g.width - 1), v = min((int)y, g.height - 1);
            const float d = __fmul_rn((float)depth16[((long long)f * g.height + v) * g.width + u], depth_factor);
            const bool ok = d > 0.f && d < 7.0f;
            kp_depth[o] = ok ? d : -1.f;
            // mvuRight = kpU.pt.x - bf / d needs the UNDISTORTED x (Frame.cc:1944, 1957): only formed here when mvKeysUn == mvKeys
            kp_uright[o] = (ok && !distorted) ? __fsub_rn(x, __fdiv_rn(bf, d)) : -1.f;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
static inline int cv_round_f(float v) { return (int)lrintf(v); }

static short sat_short_round(float v) {
    int i = (int)lrintf(v);
    return (short)std::min(32767, std::max(-32768, i));
}

}  // namespace hvo

using namespace hvo;

int hvo_orb::init() {
    const int n = p.nlevels;
    // ---- scale tables and per-level quotas (ORBextractor.cc:408-444) ----
    const double sfd = (double)p.scale_factor;
    sf.assign(n, 1.f); isf.assign(n, 1.f); sigma2.assign(n, 1.f); isigma2.assign(n, 1.f);
    for (int i = 1; i < n; ++i) {
        sf[i] = (float)((double)sf[i - 1] * sfd);
        sigma2[i] = sf[i] * sf[i];
    }
    for (int i = 0; i < n; ++i) { isf[i] = 1.0f / sf[i]; isigma2[i] = 1.0f / sigma2[i]; }
    nfeat.assign(n, 0);
    {
        const float factor = (float)(1.0 / sfd);
        float want = (float)p.nfeatures * (1.f - factor) / (1.f - (float)std::pow((double)factor, (double)n));
        int sum = 0;
        for (int l = 0; l < n - 1; ++l) {
            nfeat[l] = cv_round_f(want);
            sum += nfeat[l];
            want *= factor;
        }
        nfeat[n - 1] = std::max(p.nfeatures - sum, 0);
    }
    umax.assign(h_umax, h_umax + 16);

    // ---- level geometry ----
    std::memset(&g, 0, sizeof(g));
    g.nlevels = n; g.width = width; g.height = height;
    long long off = 0, boff = 0;
    int cand_off = 0, kp_off = 0;
    std::vector<TileDesc> btiles;
    std::vector<StripDesc> strips;
    fast_smem = 0;
    max_quota = 0;
    for (int l = 0; l < n; ++l) {
        LevelGeom& L = g.lv[l];
        L.w = cv_round_f((float)width * isf[l]);
        L.h = cv_round_f((float)height * isf[l]);
        if (L.w > 4095 || L.h > 4095) { set_error("image too large (max 4095 px per side)"); return HVO_ERR_ARG; }
        if (l == 0) { L.pitch = width; L.img_off = 0; }
        else { L.pitch = (int)align_up((size_t)L.w, 128); L.img_off = off; off += (long long)L.pitch * L.h; }
        L.minBX = kEdge - 3; L.minBY = kEdge - 3; L.maxBX = L.w - kEdge + 3; L.maxBY = L.h - kEdge + 3;
        L.quota = nfeat[l];
        L.scale = sf[l];
        L.kp_size = (float)(int)(31 * sf[l]);
        const float fw = (float)(L.maxBX - L.minBX), fh = (float)(L.maxBY - L.minBY);
        L.nCols = (int)(fw / 30.f); L.nRows = (int)(fh / 30.f);
        L.bpitch = (int)align_up((size_t)L.w, 128); L.blur_off = boff; boff += (long long)L.bpitch * L.h;
        for (int yy = 0; yy < L.h; yy += kBlH)
            for (int xx = 0; xx < L.w; xx += kBlW) { TileDesc t; t.level = (short)l; t.tx = (short)xx; t.ty = (short)yy; t.pad = 0; btiles.push_back(t); }
        L.cand_off = cand_off; L.cand_cap = 0;
        L.kp_off = kp_off; L.kp_cap = L.quota + 4;
        kp_off += L.kp_cap;
        max_quota = std::max(max_quota, L.quota);
        if (L.nCols < 1 || L.nRows < 1 || L.maxBX <= L.minBX || L.maxBY <= L.minBY) {
            // level too small for a single 30-px cell: the reference divides by zero here; produce no keypoints
            L.nCols = L.nRows = 0; L.wCell = L.hCell = 1; L.nIni = 0; L.hX = 1.f;
            continue;
        }
        L.wCell = (int)std::ceil(fw / L.nCols); L.hCell = (int)std::ceil(fh / L.nRows);
        L.nIni = (int)std::round((float)(L.maxBX - L.minBX) / (L.maxBY - L.minBY));
        L.hX = L.nIni > 0 ? (float)(L.maxBX - L.minBX) / L.nIni : 1.f;
        const int max_seg_cells = std::max(1, kFastMaxZone / L.wCell);   // a strip = a run of cells of one cell row
        for (int i = 0; i < L.nRows; ++i) {
            const float iniY = (float)(L.minBY + i * L.hCell);
            float maxY = iniY + L.hCell + 6;
            if (iniY >= L.maxBY - 3) continue;
            if (maxY > L.maxBY) maxY = (float)L.maxBY;
            const int zh = (int)maxY - (int)iniY - 6;
            if (zh <= 0) continue;
            if (zh > kFastMaxZh) { set_error("internal: FAST cell higher than %d", kFastMaxZh); return HVO_ERR_ARG; }
            StripDesc st;
            std::memset(&st, 0, sizeof(st));
            int listcap = 0;
            auto flush = [&]() {
                if (st.ncells == 0) return;
                const int xal = (st.x0 - 3) & ~15, zb0 = st.x0 - xal;
                st.tstride = (short)align_up((size_t)(zb0 + st.zw + 8), 16);
                st.score_off = (int)align_up(std::max<size_t>((size_t)(zh + 6) * st.tstride, (size_t)listcap * 4), 16);
                fast_smem = std::max(fast_smem, (size_t)16 + st.score_off + (size_t)(zh + 2) * st.tstride);
                strips.push_back(st);
                std::memset(&st, 0, sizeof(st));
                listcap = 0;
            };
            for (int j = 0; j < L.nCols; ++j) {
                const float iniX = (float)(L.minBX + j * L.wCell);
                float maxX = iniX + L.wCell + 6;
                if (iniX >= L.maxBX - 6) continue;
                if (maxX > L.maxBX) maxX = (float)L.maxBX;
                const int zw = (int)maxX - (int)iniX - 6;
                if (zw <= 0) continue;
                if (st.ncells == 0) {
                    st.level = (short)l; st.x0 = (short)((int)iniX + 3); st.y0 = (short)((int)iniY + 3); st.zh = (short)zh;
                    st.wcell = (short)L.wCell;
                }
                st.zw += (short)zw; st.ncells += 1;
                const int nms_cap = ((zw + 1) / 2) * ((zh + 1) / 2);  // NMS survivors are never 8-adjacent
                L.cand_cap += nms_cap; listcap += nms_cap;
                if (st.ncells == max_seg_cells) flush();
            }
            flush();
        }
        cand_off += L.cand_cap;
    }
    if (max_quota + 8 > 16383) { set_error("nfeatures too large"); return HVO_ERR_ARG; }
    g.pyr_frame_bytes = (long long)align_up((size_t)off, 256);
    g.blur_frame_bytes = (long long)align_up((size_t)boff, 256);
    nbtiles = (int)btiles.size();
    g.cand_total = std::max(cand_off, 1);
    g.kp_total = kp_off;
    g.out_cap = kp_off;
    nstrips = (int)strips.size();

    // ---- CUDA resources ----
    HVO_CUDA(cudaSetDevice(device));
    pin_carveout(k_resize); pin_carveout(k_fast_strips); pin_carveout(k_octree); pin_carveout(k_blur); pin_carveout(k_describe);
    HVO_CUDA(create_stream(&stream));
    for (auto& e : ev) HVO_CUDA(cudaEventCreate(&e));
    for (auto& e : tev) HVO_CUDA(cudaEventCreate(&e));
    const size_t B = (size_t)max_batch;
    HVO_CUDA(cudaMalloc(&d_pyr, std::max<size_t>(B * (size_t)g.pyr_frame_bytes, 256)));
    HVO_CUDA(cudaMalloc(&d_blur, B * (size_t)g.blur_frame_bytes));
    HVO_CUDA(cudaMalloc(&d_btiles, btiles.size() * sizeof(TileDesc)));
    HVO_CUDA(cudaMemcpy(d_btiles, btiles.data(), btiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice));
    HVO_CUDA(cudaMalloc(&d_cand, B * g.cand_total * sizeof(uint32_t)));
    HVO_CUDA(cudaMalloc(&d_knode, B * g.cand_total * sizeof(uint16_t)));
    HVO_CUDA(cudaMalloc(&d_ncand, B * n * sizeof(int)));
    HVO_CUDA(cudaMalloc(&d_okp, B * g.kp_total * sizeof(uint32_t)));
    HVO_CUDA(cudaMalloc(&d_on, B * n * sizeof(int)));
    HVO_CUDA(cudaMalloc(&d_err, sizeof(int)));
    HVO_CUDA(cudaMemset(d_err, 0, sizeof(int)));
    HVO_CUDA(cudaMalloc(&d_strips, std::max<size_t>(strips.size(), 1) * sizeof(StripDesc)));
    if (!strips.empty()) HVO_CUDA(cudaMemcpy(d_strips, strips.data(), strips.size() * sizeof(StripDesc), cudaMemcpyHostToDevice));
    if (fast_smem > 200 * 1024) { set_error("internal: FAST strip needs %zu bytes of shared memory", fast_smem); return HVO_ERR_ARG; }
    {   // the attribute is per function (shared by all handles on the device): only ever raise it
        static size_t s_fast_smem_max[64] = {0};
        if (device < 64 && fast_smem > s_fast_smem_max[device]) {
            HVO_CUDA(cudaFuncSetAttribute(k_fast_strips, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem));
            s_fast_smem_max[device] = fast_smem;
        }
    }

    // ---- resize tables (cv::resize INTER_LINEAR 8U coefficient generation, oracle/cvprims.hpp) ----
    std::vector<int2> xt;
    std::vector<int4> yt;
    xtab_off.assign(n, 0); ytab_off.assign(n, 0);
    for (int l = 1; l < n; ++l) {
        const LevelGeom& S = g.lv[l - 1];
        const LevelGeom& D = g.lv[l];
        xtab_off[l] = (int)xt.size(); ytab_off[l] = (int)yt.size();
        const double scale_x = 1.0 / ((double)D.w / S.w), scale_y = 1.0 / ((double)D.h / S.h);
        for (int dx = 0; dx < D.w; ++dx) {
            float fx = (float)((dx + 0.5) * scale_x - 0.5);
            int sx = (int)std::floor(fx);
            fx -= sx;
            if (sx < 0) { fx = 0; sx = 0; }
            if (sx >= S.w - 1) { fx = 0; sx = S.w - 1; }
            const short w0 = sat_short_round((1.f - fx) * 2048.f), w1 = sat_short_round(fx * 2048.f);
            xt.push_back(make_int2(sx, (int)(unsigned short)w0 | ((int)w1 << 16)));
        }
        for (int dy = 0; dy < D.h; ++dy) {
            float fy = (float)((dy + 0.5) * scale_y - 0.5);
            int sy = (int)std::floor(fy);
            fy -= sy;
            const short b0 = sat_short_round((1.f - fy) * 2048.f), b1 = sat_short_round(fy * 2048.f);
            const int sy0 = std::min(std::max(sy, 0), S.h - 1), sy1 = std::min(std::max(sy + 1, 0), S.h - 1);
            yt.push_back(make_int4(sy0, sy1, b0, b1));
        }
    }
    HVO_CUDA(cudaMalloc(&d_xtab, std::max<size_t>(xt.size(), 1) * sizeof(int2)));
    HVO_CUDA(cudaMalloc(&d_ytab, std::max<size_t>(yt.size(), 1) * sizeof(int4)));
    if (!xt.empty()) HVO_CUDA(cudaMemcpy(d_xtab, xt.data(), xt.size() * sizeof(int2), cudaMemcpyHostToDevice));
    if (!yt.empty()) HVO_CUDA(cudaMemcpy(d_ytab, yt.data(), yt.size() * sizeof(int4), cudaMemcpyHostToDevice));

    // quadtree shared memory: see the layout in k_octree
    const size_t CN = (size_t)max_quota + 8;
    oct_smem = CN * (8 + 16 + 8 + 4 + 4 + 4 + 16 + 8 + 2 + 2 + 1) + 64;
    if (oct_smem > 200 * 1024) { set_error("nfeatures too large for the quadtree kernel"); return HVO_ERR_ARG; }
    // the attribute is per function (shared by all handles on the device): only ever raise it
    static size_t s_oct_smem_max[64] = {0};
    const size_t want = std::max<size_t>(oct_smem, 48 * 1024);
    if (device < 64 && want > s_oct_smem_max[device]) {
        HVO_CUDA(cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
        s_oct_smem_max[device] = want;
    }
    return HVO_OK;
}

void hvo_orb::release() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    void* bufs[] = {d_blur, d_btiles, d_l0, d_depth, d_pyr, d_xtab, d_ytab, d_strips, d_cand, d_ncand, d_knode, d_okp, d_on, d_err,
                    d_kps, d_desc, d_counts, d_kpdepth, d_kpuright};
    for (void* b : bufs) if (b) cudaFree(b);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    for (auto& e : tev) if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
}

int hvo_orb::run(const uint8_t* d_gray, int nframes, hvo_keypoint* d_kps_out, uint8_t* d_desc_out, int32_t* d_counts_out,
                 const uint16_t* d_depth16, const hvo_rgbd_params* rgbd, float* d_kp_depth, float* d_kp_uright) {
    const int n = p.nlevels, B = nframes;
    int launches = 0;
    ImgSrc src;
    src.l0 = d_gray; src.l0_frame = (long long)width * height; src.l0_pitch = width; src.pyr = d_pyr;
    last_l0 = d_gray;
    last_nframes = nframes;
    if (profiling) HVO_CUDA(cudaEventRecord(ev[0], stream));
    HVO_CUDA(cudaMemsetAsync(d_ncand, 0, (size_t)B * n * sizeof(int), stream));
    // K1: pyramid, level l from level l-1
    for (int l = 1; l < n; ++l) {
        const LevelGeom& S = g.lv[l - 1];
        const LevelGeom& D = g.lv[l];
        const uint8_t* sp = l == 1 ? d_gray : d_pyr + S.img_off;
        const long long sframe = l == 1 ? src.l0_frame : g.pyr_frame_bytes;
        dim3 grd(div_up(D.w, kRsW), div_up(D.h, kRsH), B);
        timeline_mark(stream, "k_resize");
        k_resize<<<grd, 256, 0, stream>>>(sp, S.pitch, sframe, S.w, d_pyr + D.img_off, D.pitch, g.pyr_frame_bytes, D.w, D.h,
                                          d_xtab + xtab_off[l], d_ytab + ytab_off[l]);
        ++launches;
    }
    if (profiling) HVO_CUDA(cudaEventRecord(ev[1], stream));
    // K2: FAST cells
    if (nstrips > 0) {
        timeline_mark(stream, "k_fast_strips");
        k_fast_strips<<<dim3(nstrips, B), kFastThreads, fast_smem, stream>>>(g, src, d_strips, d_cand, d_ncand, p.ini_th_fast, p.min_th_fast);
        ++launches;
    }
    if (profiling) HVO_CUDA(cudaEventRecord(ev[2], stream));
    // K3: quadtree
    timeline_mark(stream, "k_octree");
    k_octree<<<dim3(n, B), 256, oct_smem, stream>>>(g, d_cand, d_ncand, d_knode, d_okp, d_on, max_quota + 8, d_err);
    ++launches;
    if (profiling) HVO_CUDA(cudaEventRecord(ev[3], stream));
    // K4: blur of every level, K5: describe
    timeline_mark(stream, "k_blur");
    k_blur<<<dim3(nbtiles, B), 256, 0, stream>>>(g, src, d_btiles, d_blur);
    ++launches;
    if (profiling) HVO_CUDA(cudaEventRecord(ev[4], stream));
    const bool rgbd_on = d_depth16 != nullptr && rgbd != nullptr && d_kp_depth != nullptr && d_kp_uright != nullptr;
    timeline_mark(stream, "k_describe");
    k_describe<<<dim3(div_up(g.out_cap, kDescWarps), B), kDescWarps * 32, 0, stream>>>(
        g, src, d_blur, d_okp, d_on, d_kps_out, d_desc_out, d_counts_out, rgbd_on ? d_depth16 : nullptr,
        rgbd_on ? rgbd->depth_factor : 0.f, rgbd_on ? rgbd->bf : 0.f, rgbd_on ? rgbd->distorted : 0, d_kp_depth, d_kp_uright);
    ++launches;
    if (profiling) { HVO_CUDA(cudaEventRecord(ev[5], stream)); have_stage_times = true; }
    HVO_CUDA(cudaGetLastError());
    last_launches = launches;
    return HVO_OK;
}
