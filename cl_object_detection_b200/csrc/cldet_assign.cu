// K1 anchors, K2 IoU-match/assign, standalone calc_iou.
// This translation unit is compiled with -fmad=false: every fp32 result here must equal the reference's
// un-fused ATen op sequence bit for bit (assignment labels and argmax are integer outputs of fp32 compares).
#include <math.h>
#include <stdio.h>

#include <algorithm>

#include <stdlib.h>

#include "cldet_common.cuh"

namespace cldet {

static thread_local cudaError_t g_last_cuda_error = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { g_last_cuda_error = e; }

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("CLDET_NO_PDL");
        return !(e && e[0] == '1');
    }();
    return on;
}

// ------------------------------------------------------------------------------------------------
// K1: anchors.  Reference: retinanet/anchors.py:21-40 (levels, ceil-div shapes), :42-73 (base boxes),
// :109-129 (shift, x fastest then y, 9 boxes per cell).  The base table is fp64 and is computed on the
// host with the reference's own expression order; the device adds the fp64 cell centre and rounds ONCE.
// ------------------------------------------------------------------------------------------------
struct AnchorPlan {
    double base[kNumLevels][kAnchorsPerCell][4];
    int64_t level_offset[kNumLevels + 1];  // in anchors
    int level_w[kNumLevels];
    int stride[kNumLevels];
};

static void build_anchor_plan(int height, int width, AnchorPlan* plan) {
    // scales = 2**0, 2**(1/3), 2**(2/3) as Python evaluates them (anchors.py:19); hex literals pin the bits.
    const double scales[3] = {1.0, 0x1.428a2f98d728bp+0, 0x1.965fea53d6e3cp+0};
    const double ratios[3] = {0.5, 1.0, 2.0};
    int64_t off = 0;
    for (int l = 0; l < kNumLevels; ++l) {
        const int level = 3 + l;
        const int stride = 1 << level;
        const double base_size = (double)(1 << (level + 2));
        int k = 0;
        for (int r = 0; r < 3; ++r) {
            for (int s = 0; s < 3; ++s, ++k) {
                const double side = base_size * scales[s];
                const double area = side * side;
                const double w = sqrt(area / ratios[r]);
                const double h = w * ratios[r];
                plan->base[l][k][0] = 0.0 - w * 0.5;
                plan->base[l][k][1] = 0.0 - h * 0.5;
                plan->base[l][k][2] = w - w * 0.5;
                plan->base[l][k][3] = h - h * 0.5;
            }
        }
        const int hl = (height + stride - 1) / stride;
        const int wl = (width + stride - 1) / stride;
        plan->level_offset[l] = off;
        plan->level_w[l] = wl;
        plan->stride[l] = stride;
        off += (int64_t)hl * wl * kAnchorsPerCell;
    }
    plan->level_offset[kNumLevels] = off;
}

__global__ void __launch_bounds__(256) anchors_kernel(const AnchorPlan plan, float4* __restrict__ out) {
    const int64_t total = plan.level_offset[kNumLevels];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int l = 0;
#pragma unroll
        for (int t = 1; t < kNumLevels; ++t) l += (i >= plan.level_offset[t]) ? 1 : 0;
        const int64_t rel = i - plan.level_offset[l];
        const int64_t cell = rel / kAnchorsPerCell;
        const int k = (int)(rel - cell * kAnchorsPerCell);
        const int wl = plan.level_w[l];
        const int y = (int)(cell / wl);
        const int x = (int)(cell - (int64_t)y * wl);
        const double sx = ((double)x + 0.5) * (double)plan.stride[l];
        const double sy = ((double)y + 0.5) * (double)plan.stride[l];
        float4 o;
        o.x = (float)(plan.base[l][k][0] + sx);
        o.y = (float)(plan.base[l][k][1] + sy);
        o.z = (float)(plan.base[l][k][2] + sx);
        o.w = (float)(plan.base[l][k][3] + sy);
        out[i] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// K2: IoU match + assignment.  Reference: calc_iou (retinanet/losses.py:4-21), GT filter (:287-288),
// torch.max(IoU, dim=1) (:310, first maximal index), thresholds lt 0.4 / ge 0.5 (:316, :330).
//
// Layout: one thread per (image, anchor); the image's GT rows are staged in shared memory tile by tile
// and read as warp-wide broadcasts.  Padding rows (label == -1) are skipped with a block-uniform branch,
// so the running count of valid rows IS the compacted index the reference's argmax refers to -- no
// compaction pass and no host sync.  The [A,G] IoU matrix is never written.
// The divide is skipped when the intersection is empty (IoU = +0 exactly, as 0/ua with ua >= 1e-8).
// ------------------------------------------------------------------------------------------------
constexpr int kGtTile = 256;
constexpr int kAssignThreads = 256;
constexpr int kAnchorsPerThread = 4;

__global__ void __launch_bounds__(kAssignThreads, 5)
iou_assign_kernel(const float4* __restrict__ anchors, int64_t A, const float* __restrict__ annotations, int G,
                  int num_classes, uint32_t* __restrict__ meta, int32_t* __restrict__ argmax_out,
                  float* __restrict__ iou_max_out, int32_t* __restrict__ npos, int32_t* __restrict__ nvalid) {
    pdl_launch_dependents();
    // valid rows of the current tile, compacted in order: box, area, (label, raw row)
    __shared__ float4 s_box[kGtTile];
    __shared__ float s_area[kGtTile];
    __shared__ int s_idx[kGtTile];        // compacted GT index of each staged (live) row
    __shared__ float4 s_bbox[kAssignThreads / 32];
    __shared__ int s_tile_live;
    __shared__ int s_warp[kAssignThreads / 32];
    __shared__ int s_tile_valid;
    __shared__ int2 s_first[kGtTile];     // (label, raw row) of the image's first kGtTile valid rows

    const int j = blockIdx.y;
    const int64_t a_base = (int64_t)blockIdx.x * (kAssignThreads * kAnchorsPerThread) + threadIdx.x;
    const float* ann = annotations + (int64_t)j * G * 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // kAnchorsPerThread anchors per thread (strided by the block size: coalesced), so one broadcast read of a GT row
    // feeds several independent IoU chains and the per-block tile set-up is amortised.
    // first GT tile: issue this thread's row load before anything depends on the anchors (both latencies overlap)
    float4 pre_b = make_float4(0.f, 0.f, 0.f, 0.f);
    float pre_lab = -1.0f;
    if (threadIdx.x < min(kGtTile, G)) {
        const float* r = ann + (int64_t)threadIdx.x * 5;
        pre_b = make_float4(r[0], r[1], r[2], r[3]);
        pre_lab = r[4];
    }

    float4 box[kAnchorsPerThread];
    float area_a[kAnchorsPerThread];
    float best[kAnchorsPerThread];
    int best_t[kAnchorsPerThread];     // compacted index of the winner
#pragma unroll
    for (int u = 0; u < kAnchorsPerThread; ++u) {
        const int64_t a = a_base + u * kAssignThreads;
        box[u] = (a < A) ? anchors[a] : make_float4(0.f, 0.f, 0.f, 0.f);
        area_a[u] = (box[u].z - box[u].x) * (box[u].w - box[u].y);
        // torch.max returns the FIRST maximal index and every IoU is >= 0, so the first valid row is the answer unless
        // a later row has a strictly larger (hence strictly positive) IoU: start from (IoU 0, compacted index 0).
        best[u] = 0.0f;
        best_t[u] = 0;
    }
    int k = 0;            // valid rows seen so far (block-uniform) = compacted index of the next valid row

    // Bounding box of this block's anchors: a GT row that does not overlap it has IoU 0 with every anchor here and can
    // never beat the default winner, so it is culled for the WHOLE block when the tile is staged (it still counts in k).
    {
        float x1 = INFINITY, y1 = INFINITY, x2 = -INFINITY, y2 = -INFINITY;
#pragma unroll
        for (int u = 0; u < kAnchorsPerThread; ++u) {
            if (a_base + u * kAssignThreads < A) {
                x1 = fminf(x1, box[u].x);
                y1 = fminf(y1, box[u].y);
                x2 = fmaxf(x2, box[u].z);
                y2 = fmaxf(y2, box[u].w);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            x1 = fminf(x1, __shfl_xor_sync(0xffffffffu, x1, o));
            y1 = fminf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
            x2 = fmaxf(x2, __shfl_xor_sync(0xffffffffu, x2, o));
            y2 = fmaxf(y2, __shfl_xor_sync(0xffffffffu, y2, o));
        }
        if (lane == 0) s_bbox[warp] = make_float4(x1, y1, x2, y2);
    }
    __syncthreads();
    float4 bb = s_bbox[0];
#pragma unroll
    for (int w = 1; w < kAssignThreads / 32; ++w) {
        const float4 o = s_bbox[w];
        bb.x = fminf(bb.x, o.x);
        bb.y = fminf(bb.y, o.y);
        bb.z = fmaxf(bb.z, o.z);
        bb.w = fmaxf(bb.w, o.w);
    }

    for (int g0 = 0; g0 < G; g0 += kGtTile) {
        const int n = min(kGtTile, G - g0);
        __syncthreads();
        // ordered compaction of this tile's valid rows (label != -1, losses.py:288); tile size == block size
        float4 b = pre_b;
        float lab = pre_lab;
        if (g0 > 0) {
            b = make_float4(0.f, 0.f, 0.f, 0.f);
            lab = -1.0f;
            if (threadIdx.x < n) {
                const float* r = ann + (int64_t)(g0 + threadIdx.x) * 5;
                b = make_float4(r[0], r[1], r[2], r[3]);
                lab = r[4];
            }
        }
        const bool valid = (threadIdx.x < n) && (lab != -1.0f);
        // live = valid and with a non-empty intersection with the block's bounding box (same strict tests as the IoU)
        const bool live = valid && (fminf(bb.z, b.z) - fmaxf(bb.x, b.x) > 0.0f) && (fminf(bb.w, b.w) - fmaxf(bb.y, b.y) > 0.0f);
        const unsigned ballot = __ballot_sync(0xffffffffu, valid);
        const unsigned ballot_live = __ballot_sync(0xffffffffu, live);
        if (lane == 0) s_warp[warp] = __popc(ballot) | (__popc(ballot_live) << 16);
        __syncthreads();
        int before = 0, before_live = 0;
#pragma unroll
        for (int w = 0; w < kAssignThreads / 32; ++w) {
            const int v = (w < warp) ? s_warp[w] : 0;
            before += v & 0xffff;
            before_live += v >> 16;
        }
        if (valid) {
            const int slot = before + __popc(ballot & ((1u << lane) - 1u));       // compacted index inside this tile
            // (label, raw row) of every valid row of the IMAGE, by compacted index, for the epilogue
            if (k + slot < kGtTile) s_first[k + slot] = make_int2((int)(long long)lab, g0 + threadIdx.x);   // .long(), losses.py:341
            if (live) {
                const int ls = before_live + __popc(ballot_live & ((1u << lane) - 1u));
                s_box[ls] = b;
                s_area[ls] = (b.z - b.x) * (b.w - b.y);
                s_idx[ls] = k + slot;
            }
        }
        if (threadIdx.x == kAssignThreads - 1) {
            s_tile_valid = before + __popc(ballot);
            s_tile_live = before_live + __popc(ballot_live);
        }
        __syncthreads();
        const int nv = s_tile_valid;
        const int nl = s_tile_live;
        for (int t = 0; t < nl; ++t) {
            const float4 gb = s_box[t];                        // warp-wide broadcast
            const float ga = s_area[t];
            const int gi = s_idx[t];
#pragma unroll
            for (int u = 0; u < kAnchorsPerThread; ++u) {
                // calc_iou (losses.py:4-21) with the clamps resolved by early exits: an empty intersection gives IoU
                // +0, which can never beat `best` under the strict compare.
                const float iw = fminf(box[u].z, gb.z) - fmaxf(box[u].x, gb.x);
                const float ih = fminf(box[u].w, gb.w) - fmaxf(box[u].y, gb.y);
                if (iw > 0.0f && ih > 0.0f) {
                    const float inter = iw * ih;
                    float ua = (area_a[u] + ga) - inter;
                    ua = fmaxf(ua, 1e-8f);
                    const float v = __fdiv_rn(inter, ua);
                    if (v > best[u]) {                         // strict: first maximal index (torch.max)
                        best[u] = v;
                        best_t[u] = gi;
                    }
                }
            }
        }
        k += nv;
    }
    __syncthreads();

    int npos_local = 0;
#pragma unroll
    for (int u = 0; u < kAnchorsPerThread; ++u) {
        const int64_t a = a_base + u * kAssignThreads;
        uint32_t state;
        if (k == 0) {
            state = CLDET_STATE_EMPTY;
        } else if (best[u] >= 0.5f) {       // torch.ge(IoU_max, 0.5)   losses.py:330
            state = CLDET_STATE_POS;
        } else if (best[u] < 0.4f) {        // torch.lt(IoU_max, 0.4)   losses.py:316
            state = CLDET_STATE_BG;
        } else {
            state = CLDET_STATE_IGNORE;
        }
        if (a < A) {
            // winner's (label, raw row): from the staged table, or (images with > kGtTile valid rows) re-derived from
            // the annotations by walking to the best_t-th valid row
            int2 lr = make_int2(0, 0);
            if (k > 0) {
                if (best_t[u] < kGtTile) {
                    lr = s_first[best_t[u]];
                } else {
                    int seen = 0;
                    for (int g = 0; g < G; ++g) {
                        const float lab = ann[(int64_t)g * 5 + 4];
                        if (lab != -1.0f) {
                            if (seen == best_t[u]) {
                                lr = make_int2((int)(long long)lab, g);
                                break;
                            }
                            ++seen;
                        }
                    }
                }
            }
            const uint32_t lab = (lr.x >= 0 && lr.x < num_classes) ? (uint32_t)lr.x : CLDET_BAD_LABEL;
            meta[(int64_t)j * A + a] = meta_pack(state, lab, (uint32_t)lr.y);
            if (argmax_out) argmax_out[(int64_t)j * A + a] = (k == 0) ? -1 : best_t[u];
            if (iou_max_out) iou_max_out[(int64_t)j * A + a] = best[u];
            npos_local += (state == CLDET_STATE_POS) ? 1 : 0;
        }
    }
    // positives per image: warp -> block -> one integer atomic (deterministic)
    const int c = warp_sum_int(npos_local);
    if (lane == 0) s_warp[warp] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < kAssignThreads / 32; ++w) tot += s_warp[w];
        if (tot) atomicAdd(&npos[j], tot);
        if (blockIdx.x == 0) nvalid[j] = k;
    }
}

__global__ void __launch_bounds__(256)
calc_iou_kernel(const float4* __restrict__ a, int64_t na, const float4* __restrict__ b, int nb, float* __restrict__ out) {
    const int64_t total = na * nb;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ia = i / nb;
        const int ib = (int)(i - ia * nb);
        const float4 p = a[ia];
        const float4 q = b[ib];
        const float area_a = (p.z - p.x) * (p.w - p.y);
        const float area_b = (q.z - q.x) * (q.w - q.y);
        float iw = fmaxf(fminf(p.z, q.z) - fmaxf(p.x, q.x), 0.0f);
        float ih = fmaxf(fminf(p.w, q.w) - fmaxf(p.y, q.y), 0.0f);
        float ua = fmaxf((area_a + area_b) - iw * ih, 1e-8f);
        out[i] = __fdiv_rn(iw * ih, ua);
    }
}

// ------------------------------------------------------------------------------------------------
// K2' GT-centric assignment (used by the fused loss call when the anchors are the standard grid of a known image size).
//
// IoU >= 0.4 forces  iw >= 0.4*max(w_a, w_g)  and  ih >= 0.4*max(h_a, h_g)  (I >= 0.4*U >= 0.4*max(area), and the
// intersection is no wider/taller than either box), while  iw <= (w_a + w_g)/2 - |cx_a - cx_g|.  So for a GT box only the
// anchor types (level, ratio, scale) with min(w)/max(w) >= 0.4 and min(h)/max(h) >= 0.4 matter, and of those only the cells
// whose centre is within  (w_a + w_g)/2 - 0.4*max(w_a, w_g)  of the GT centre (same in y).  That is a few hundred anchors
// per GT instead of all 200 k.  Everything visited gets the EXACT reference IoU (same fp32 op order) folded into
// best[anchor] with a 64-bit integer atomicMax on the key (IoU bits << 32 | ~GT row): IoU >= 0, so float bits order like
// unsigned integers, and among equal IoUs the SMALLEST row wins -- torch.max's first maximal index.  For every anchor whose
// true maximum is >= 0.4 the stored key is exactly (maximum, argmax); all others stay below 0.4 = background.  The bounds use 0.39
// plus slack and one extra cell on each side, so rounding can only ADD candidates.
// ------------------------------------------------------------------------------------------------
struct ScatterPlan {
    float type_w[kNumLevels][kAnchorsPerCell];
    float type_h[kNumLevels][kAnchorsPerCell];
    int64_t level_offset[kNumLevels + 1];
    int level_w[kNumLevels];
    int level_h[kNumLevels];
    int stride[kNumLevels];
};

__global__ void __launch_bounds__(128)
gt_scatter_kernel(const ScatterPlan plan, const float4* __restrict__ anchors, int64_t A, const float* __restrict__ annotations,
                  int G, unsigned long long* __restrict__ best, uint32_t* __restrict__ touched, int32_t* __restrict__ npos_acc,
                  int32_t* __restrict__ nvalid) {
    pdl_launch_dependents();      // the loss kernel may be scheduled while this grid drains (it waits before reading)
    // grid = (GT row, image, pyramid level): one block visits the candidate anchors of one GT box on one level
    const int j = blockIdx.y;
    const int l = blockIdx.z;
    const float* ann = annotations + (int64_t)j * G * 5;
    if (blockIdx.x == 0 && l == 0) {                         // valid rows of the image (label != -1, losses.py:288)
        __shared__ int s_cnt[4];
        int c = 0;
        for (int g = threadIdx.x; g < G; g += blockDim.x) c += (ann[(int64_t)g * 5 + 4] != -1.0f) ? 1 : 0;
        c = warp_sum_int(c);
        if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) nvalid[j] = s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
    }
    const int g = blockIdx.x;
    const float* r = ann + (int64_t)g * 5;
    if (r[4] == -1.0f) return;                               // padding row
    const float gx1 = r[0], gy1 = r[1], gx2 = r[2], gy2 = r[3];
    const float gw = gx2 - gx1, gh = gy2 - gy1;
    if (!(gw > 0.0f) || !(gh > 0.0f)) return;                // degenerate box: IoU is 0 (or NaN) with every anchor
    const float area_g = gw * gh;
    const float gcx = 0.5f * (gx1 + gx2), gcy = 0.5f * (gy1 + gy2);
    const float s = (float)plan.stride[l];
    const int wl = plan.level_w[l], hl = plan.level_h[l];

    // candidate cell window of each of the 9 anchor types of this level (block-uniform), flattened into one index space
    __shared__ int s_x0[kAnchorsPerCell], s_y0[kAnchorsPerCell], s_nx[kAnchorsPerCell], s_first[kAnchorsPerCell + 1];
    if (threadIdx.x < kAnchorsPerCell) {
        const int k = threadIdx.x;
        const float wa = plan.type_w[l][k], ha = plan.type_h[l][k];
        int cells = 0, x0 = 0, y0 = 0, nx = 0;
        // type filter (0.39 < 0.4: conservative)
        if (!(fminf(wa, gw) < 0.39f * fmaxf(wa, gw) || fminf(ha, gh) < 0.39f * fmaxf(ha, gh))) {
            const float dx = 0.5f * (wa + gw) - 0.39f * fmaxf(wa, gw) + 1e-3f * (wa + gw) + 0.05f;
            const float dy = 0.5f * (ha + gh) - 0.39f * fmaxf(ha, gh) + 1e-3f * (ha + gh) + 0.05f;
            // cell centres (x + 0.5) * s within [gc - d, gc + d], one extra cell on each side
            x0 = max((int)floorf((gcx - dx) / s - 0.5f) - 1, 0);
            y0 = max((int)floorf((gcy - dy) / s - 0.5f) - 1, 0);
            const int x1 = min((int)ceilf((gcx + dx) / s - 0.5f) + 1, wl - 1);
            const int y1 = min((int)ceilf((gcy + dy) / s - 0.5f) + 1, hl - 1);
            if (x1 >= x0 && y1 >= y0) {
                nx = x1 - x0 + 1;
                cells = nx * (y1 - y0 + 1);
            }
        }
        s_x0[k] = x0; s_y0[k] = y0; s_nx[k] = nx;
        s_first[k + 1] = cells;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        s_first[0] = 0;
        for (int k = 0; k < kAnchorsPerCell; ++k) s_first[k + 1] += s_first[k];
    }
    __syncthreads();
    const int total = s_first[kAnchorsPerCell];
    unsigned long long* best_j = best + (int64_t)j * A;
    // one bit per anchor that received a key: the loss kernel reads 1 bit instead of 8 bytes for the ~98 % that did not
    uint32_t* touched_j = touched ? touched + (int64_t)j * ((A + 31) / 32) : nullptr;
    const unsigned long long low = 0xFFFFFFFFull - (unsigned long long)g;      // ties -> smallest row wins the max
    int crossings = 0;
    for (int c = threadIdx.x; c < total; c += blockDim.x) {
        int k = 0;
#pragma unroll
        for (int t = 1; t < kAnchorsPerCell; ++t) k += (c >= s_first[t]) ? 1 : 0;
        const int rel = c - s_first[k];
        const int nx = s_nx[k];
        const int yy = s_y0[k] + rel / nx, xx = s_x0[k] + (rel - (rel / nx) * nx);
        const int64_t idx = plan.level_offset[l] + ((int64_t)yy * wl + xx) * kAnchorsPerCell + k;
        const float4 an = anchors[idx];
        // exact calc_iou (losses.py:4-21), this translation unit is built with -fmad=false
        const float iw = fminf(an.z, gx2) - fmaxf(an.x, gx1);
        const float ih = fminf(an.w, gy2) - fmaxf(an.y, gy1);
        if (iw > 0.0f && ih > 0.0f) {
            const float area_a = (an.z - an.x) * (an.w - an.y);
            const float inter = iw * ih;
            float ua = (area_a + area_g) - inter;
            ua = fmaxf(ua, 1e-8f);
            const float v = __fdiv_rn(inter, ua);
            if (v >= 0.39f) {                                   // lower values cannot change any anchor's state
                const unsigned long long key = ((unsigned long long)__float_as_uint(v) << 32) | low;
                const unsigned long long old = atomicMax(best_j + idx, key);
                if (touched_j && old == 0ull) atomicOr(touched_j + (idx >> 5), 1u << (idx & 31));
                // the anchor becomes positive exactly once: when its maximum first reaches 0.5 (losses.py:330)
                if ((uint32_t)(old >> 32) < 0x3f000000u && v >= 0.5f) ++crossings;
            }
        }
    }
    crossings = warp_sum_int(crossings);
    if ((threadIdx.x & 31) == 0 && crossings) atomicAdd(&npos_acc[j], crossings);
}

// fp64 IoU + max over the second set (IL_method/persuado_label.py:68-72: calc_iou on float64 annotations, .max(dim=1))
__global__ void __launch_bounds__(128)
iou_max_f64_kernel(const double* __restrict__ a, int64_t na, const double* __restrict__ b, int nb, double* __restrict__ out_max,
                   int32_t* __restrict__ out_arg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na) return;
    const double ax1 = a[4 * i], ay1 = a[4 * i + 1], ax2 = a[4 * i + 2], ay2 = a[4 * i + 3];
    const double area_a = (ax2 - ax1) * (ay2 - ay1);
    double best = -1.0;
    int arg = -1;
    for (int g = 0; g < nb; ++g) {
        const double bx1 = b[4 * g], by1 = b[4 * g + 1], bx2 = b[4 * g + 2], by2 = b[4 * g + 3];
        const double area_b = (bx2 - bx1) * (by2 - by1);
        double iw = fmax(fmin(ax2, bx2) - fmax(ax1, bx1), 0.0);
        double ih = fmax(fmin(ay2, by2) - fmax(ay1, by1), 0.0);
        const double ua = fmax((area_a + area_b) - iw * ih, 1e-8);
        const double v = __ddiv_rn(iw * ih, ua);
        if (v > best) {
            best = v;
            arg = g;
        }
    }
    out_max[i] = best;
    if (out_arg) out_arg[i] = arg;
}

}  // namespace cldet

namespace cldet {

int launch_gt_scatter(int height, int width, const float* d_anchors, int64_t num_anchors, const float* d_annotations,
                      int num_images, int gt_rows, unsigned long long* d_best, uint32_t* d_touched, int32_t* d_npos_acc,
                      int32_t* d_nvalid, cudaStream_t s) {
    AnchorPlan ap;
    build_anchor_plan(height, width, &ap);
    if (ap.level_offset[kNumLevels] != num_anchors) return CLDET_ERR_UNSUPPORTED;
    ScatterPlan sp;
    for (int l = 0; l < kNumLevels; ++l) {
        for (int k = 0; k < kAnchorsPerCell; ++k) {
            sp.type_w[l][k] = (float)(ap.base[l][k][2] - ap.base[l][k][0]);
            sp.type_h[l][k] = (float)(ap.base[l][k][3] - ap.base[l][k][1]);
        }
        sp.level_offset[l] = ap.level_offset[l];
        sp.level_w[l] = ap.level_w[l];
        sp.stride[l] = ap.stride[l];
        sp.level_h[l] = (height + ap.stride[l] - 1) / ap.stride[l];
    }
    sp.level_offset[kNumLevels] = ap.level_offset[kNumLevels];
    dim3 grid((unsigned)gt_rows, (unsigned)num_images, (unsigned)kNumLevels);
    gt_scatter_kernel<<<grid, 128, 0, s>>>(sp, reinterpret_cast<const float4*>(d_anchors), num_anchors, d_annotations, gt_rows,
                                           d_best, d_touched, d_npos_acc, d_nvalid);
    if (cudaPeekAtLastError() != cudaSuccess) {
        set_last_cuda_error(cudaGetLastError());
        return CLDET_ERR_CUDA;
    }
    return CLDET_OK;
}

}  // namespace cldet

using namespace cldet;

extern "C" {

int cldet_abi_version(void) { return CLDET_ABI_VERSION; }

const char* cldet_status_string(int status) {
    switch (status) {
        case CLDET_OK: return "ok";
        case CLDET_ERR_INVALID_ARGUMENT: return "invalid argument";
        case CLDET_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
        case CLDET_ERR_CUDA: return "CUDA runtime error";
        case CLDET_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown status";
    }
}

const char* cldet_last_cuda_error(void) { return cudaGetErrorString(g_last_cuda_error); }

int cldet_num_anchors(int height, int width, int64_t* out_num_anchors) {
    if (height <= 0 || width <= 0 || !out_num_anchors) return CLDET_ERR_INVALID_ARGUMENT;
    AnchorPlan plan;
    build_anchor_plan(height, width, &plan);
    *out_num_anchors = plan.level_offset[kNumLevels];
    return CLDET_OK;
}

int cldet_anchors(int height, int width, float* d_anchors, void* stream) {
    if (height <= 0 || width <= 0 || !d_anchors) return CLDET_ERR_INVALID_ARGUMENT;
    AnchorPlan plan;
    build_anchor_plan(height, width, &plan);
    const int64_t total = plan.level_offset[kNumLevels];
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 8);
    anchors_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(plan, reinterpret_cast<float4*>(d_anchors));
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

int cldet_iou_assign(const float* d_anchors, int64_t num_anchors, const float* d_annotations, int num_images,
                     int gt_rows, int num_classes, uint32_t* d_meta, int32_t* d_argmax, float* d_iou_max,
                     int32_t* d_npos, int32_t* d_nvalid, void* stream) {
    if (!d_anchors || !d_annotations || !d_meta || !d_npos || !d_nvalid) return CLDET_ERR_INVALID_ARGUMENT;
    if (num_anchors <= 0 || num_images <= 0 || gt_rows <= 0 || gt_rows > CLDET_MAX_GT_ROWS) return CLDET_ERR_INVALID_ARGUMENT;
    if (num_classes <= 0 || num_classes > CLDET_MAX_CLASSES || num_images > 65535) return CLDET_ERR_INVALID_ARGUMENT;
    const int64_t per_block = (int64_t)kAssignThreads * kAnchorsPerThread;
    dim3 grid((unsigned)((num_anchors + per_block - 1) / per_block), (unsigned)num_images);
    iou_assign_kernel<<<grid, kAssignThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(d_anchors), num_anchors, d_annotations, gt_rows, num_classes, d_meta, d_argmax,
        d_iou_max, d_npos, d_nvalid);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

int cldet_iou_max_f64(const double* d_a, int64_t num_a, const double* d_b, int num_b, double* d_max, int32_t* d_argmax,
                      void* stream) {
    if (!d_a || !d_b || !d_max || num_a < 0 || num_b <= 0) return CLDET_ERR_INVALID_ARGUMENT;
    if (num_a == 0) return CLDET_OK;
    iou_max_f64_kernel<<<(unsigned)((num_a + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_a, num_a, d_b, num_b, d_max, d_argmax);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

int cldet_calc_iou(const float* d_a, int64_t num_a, const float* d_b, int num_b, float* d_iou, void* stream) {
    if (!d_a || !d_b || !d_iou || num_a < 0 || num_b < 0) return CLDET_ERR_INVALID_ARGUMENT;
    if (num_a == 0 || num_b == 0) return CLDET_OK;
    const int64_t total = num_a * num_b;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
    calc_iou_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(d_a), num_a,
                                                              reinterpret_cast<const float4*>(d_b), num_b, d_iou);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

}  // extern "C"
