// K1 anchors, K2 IoU-match/assign, standalone calc_iou.
// This translation unit is compiled with -fmad=false: every fp32 result here must equal the reference's
// un-fused ATen op sequence bit for bit (assignment labels and argmax are integer outputs of fp32 compares).
#include <math.h>
#include <stdio.h>

#include <algorithm>

#include "cldet_common.cuh"

namespace cldet {

static thread_local cudaError_t g_last_cuda_error = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { g_last_cuda_error = e; }

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

struct GtRow {
    float x1, y1, x2, y2, area;
    int label;  // -1 (as int) marks a padding row
    int valid;
};

__device__ __forceinline__ float iou_exact(float ax1, float ay1, float ax2, float ay2, float area_a, float bx1, float by1,
                                           float bx2, float by2, float area_b) {
    float iw = fminf(ax2, bx2) - fmaxf(ax1, bx1);
    float ih = fminf(ay2, by2) - fmaxf(ay1, by1);
    iw = fmaxf(iw, 0.0f);
    ih = fmaxf(ih, 0.0f);
    const float inter = iw * ih;
    float ua = (area_a + area_b) - inter;
    ua = fmaxf(ua, 1e-8f);
    // inter == 0 -> the quotient is +0 for every finite ua >= 1e-8; skip the IEEE divide (most pairs).
    return (inter > 0.0f) ? __fdiv_rn(inter, ua) : 0.0f;
}

__global__ void __launch_bounds__(kAssignThreads)
iou_assign_kernel(const float4* __restrict__ anchors, int64_t A, const float* __restrict__ annotations, int G,
                  int num_classes, uint32_t* __restrict__ meta, int32_t* __restrict__ argmax_out,
                  float* __restrict__ iou_max_out, int32_t* __restrict__ npos, int32_t* __restrict__ nvalid) {
    __shared__ GtRow tile[kGtTile];
    __shared__ int warp_counts[kAssignThreads / 32];

    const int j = blockIdx.y;
    const int64_t a = (int64_t)blockIdx.x * kAssignThreads + threadIdx.x;
    const bool active = a < A;
    const float* ann = annotations + (int64_t)j * G * 5;

    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) box = anchors[a];
    const float area_a = (box.z - box.x) * (box.w - box.y);

    float best = -1.0f;   // IoU >= 0 always, so the first valid row always wins the first compare
    int best_k = -1;      // compacted index
    int best_row = 0;     // raw row
    int best_label = 0;
    int k = 0;            // valid rows seen so far (block-uniform)

    for (int g0 = 0; g0 < G; g0 += kGtTile) {
        const int n = min(kGtTile, G - g0);
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += kAssignThreads) {
            const float* r = ann + (int64_t)(g0 + t) * 5;
            GtRow row;
            row.x1 = r[0];
            row.y1 = r[1];
            row.x2 = r[2];
            row.y2 = r[3];
            const float lab = r[4];
            row.area = (row.x2 - row.x1) * (row.y2 - row.y1);
            row.valid = (lab != -1.0f) ? 1 : 0;          // losses.py:288  annotation[:, 4] != -1
            row.label = (int)(long long)lab;             // .long() truncation, losses.py:341
            tile[t] = row;
        }
        __syncthreads();
        for (int t = 0; t < n; ++t) {
            const GtRow row = tile[t];                   // broadcast read
            if (!row.valid) continue;                    // block-uniform
            const float v = iou_exact(box.x, box.y, box.z, box.w, area_a, row.x1, row.y1, row.x2, row.y2, row.area);
            if (v > best) {                              // strict: first maximal index (torch.max)
                best = v;
                best_k = k;
                best_row = g0 + t;
                best_label = row.label;
            }
            ++k;
        }
    }

    uint32_t state;
    int is_pos = 0;
    if (k == 0) {
        state = CLDET_STATE_EMPTY;
    } else if (best >= 0.5f) {       // torch.ge(IoU_max, 0.5)   losses.py:330
        state = CLDET_STATE_POS;
        is_pos = 1;
    } else if (best < 0.4f) {        // torch.lt(IoU_max, 0.4)   losses.py:316
        state = CLDET_STATE_BG;
    } else {
        state = CLDET_STATE_IGNORE;
    }
    if (active) {
        uint32_t lab = (best_label >= 0 && best_label < num_classes) ? (uint32_t)best_label : CLDET_BAD_LABEL;
        meta[(int64_t)j * A + a] = meta_pack(state, lab, (uint32_t)best_row);
        if (argmax_out) argmax_out[(int64_t)j * A + a] = best_k;
        if (iou_max_out) iou_max_out[(int64_t)j * A + a] = (k == 0) ? 0.0f : best;
    }
    // positives per image: warp -> block -> one integer atomic (deterministic)
    int c = warp_sum_int((active && is_pos) ? 1 : 0);
    if ((threadIdx.x & 31) == 0) warp_counts[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < kAssignThreads / 32; ++w) tot += warp_counts[w];
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
    dim3 grid((unsigned)((num_anchors + kAssignThreads - 1) / kAssignThreads), (unsigned)num_images);
    iou_assign_kernel<<<grid, kAssignThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(d_anchors), num_anchors, d_annotations, gt_rows, num_classes, d_meta, d_argmax,
        d_iou_max, d_npos, d_nvalid);
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
