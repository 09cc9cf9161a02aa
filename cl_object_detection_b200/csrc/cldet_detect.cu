// K4 decode+filter, K5 radix-select top-k + ordering, K6 bitmask NMS, for the eval-mode detection output.
// Reference: retinanet/utils.py:102-126 (BBoxTransform), :134-144 (ClipBoxes); retinanet/model.py:507-550
// (ResNet.predict); IL_method/persuado_label.py:99-127 (Labeler.predict); torchvision.ops.batched_nms
// (0.26.0: ops/boxes.py _batched_nms_coordinate_trick/_batched_nms_vanilla + torchvision::nms).
//
// Compiled with -fmad=false: decoded boxes, scores and IoU compares must round exactly like the reference's
// separate ATen ops, because keep indices / labels are integer outputs of those fp32 values.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include <cooperative_groups.h>

#include "cldet_common.cuh"

namespace cldet {

__host__ __device__ __forceinline__ int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }

// ------------------------------------------------------------------------------------------------
// Box decode + clip (utils.py:102-126, 134-144); std = (0.1, 0.1, 0.2, 0.2), mean = 0.
// ------------------------------------------------------------------------------------------------
template <bool CLIP = true>
__device__ __forceinline__ float4 decode_clip(const float4 an, const float4 d, float img_w, float img_h) {
    const float w = an.z - an.x;
    const float h = an.w - an.y;
    const float cx = an.x + 0.5f * w;
    const float cy = an.y + 0.5f * h;
    const float dx = d.x * 0.1f + 0.0f;
    const float dy = d.y * 0.1f + 0.0f;
    const float dw = d.z * 0.2f + 0.0f;
    const float dh = d.w * 0.2f + 0.0f;
    const float pcx = cx + dx * w;
    const float pcy = cy + dy * h;
    const float pw = expf(dw) * w;
    const float ph = expf(dh) * h;
    float4 o;
    o.x = pcx - 0.5f * pw;
    o.y = pcy - 0.5f * ph;
    o.z = pcx + 0.5f * pw;
    o.w = pcy + 0.5f * ph;
    if (CLIP) {
        o.x = fmaxf(o.x, 0.0f);      // clamp(min=0)
        o.y = fmaxf(o.y, 0.0f);
        o.z = fminf(o.z, img_w);     // clamp(max=width)
        o.w = fminf(o.w, img_h);
    }
    return o;
}

template <bool CLIP>
__global__ void __launch_bounds__(256) decode_boxes_kernel(const float4* __restrict__ anchors, const float4* __restrict__ reg,
                                                           int64_t A, int64_t total, float img_w, float img_h,
                                                           float4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = decode_clip<CLIP>(anchors[i % A], reg[i], img_w, img_h);
}

// ClipBoxes.forward (utils.py:134-144), in place
__global__ void __launch_bounds__(256) clip_boxes_kernel(float4* __restrict__ boxes, int64_t total, float img_w, float img_h) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float4 b = boxes[i];
        b.x = fmaxf(b.x, 0.0f);
        b.y = fmaxf(b.y, 0.0f);
        b.z = fminf(b.z, img_w);
        b.w = fminf(b.w, img_h);
        boxes[i] = b;
    }
}

// ATen's CUDA sigmoid for float: 1 / (1 + exp(-x)), IEEE divide.
__device__ __forceinline__ float sigmoid_exact(float x) { return __fdiv_rn(1.0f, 1.0f + expf(-x)); }

// order-preserving float -> uint32 (larger float <=> larger uint)
__device__ __forceinline__ uint32_t float_ordered(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
// sort key: higher score first; equal scores: lower anchor index first.  All keys of one image are distinct.
__device__ __forceinline__ uint64_t make_key(float score, int anchor) {
    return ((uint64_t)float_ordered(score) << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)anchor);
}

// ------------------------------------------------------------------------------------------------
// K4: fused class-max + threshold + decode.  Block = 256 threads, R anchor rows of one image.
//   phase 1: coalesced 128-bit sweep over the R*C logits; each vector's max goes to shared memory
//   phase 2: one thread per row reduces its row's partials; rows whose max can pass the threshold re-read only the
//            vectors that can hold the winning class, evaluate the exact sigmoid there (first-index tie rule of
//            torch.max), decode + clip the box and join a block-aggregated append (one atomic per block).
// Only survivors are decoded; the [N,A,C] sigmoid map and the [N,A,4] box tensor are never materialised.
// ------------------------------------------------------------------------------------------------
constexpr int kFilterThreads = 256;
constexpr int kFilterMaxPartials = 8192;   // 32 KB of shared memory per block

template <int VEC>
__global__ void __launch_bounds__(kFilterThreads)
decode_filter_kernel(const float* __restrict__ cls, int is_logits, const float4* __restrict__ reg,
                     const float4* __restrict__ anchors, int64_t A, int C, int rows_per_block, int stride, float img_w,
                     float img_h, float score_thresh, float prefilter, cldet_candidate* __restrict__ cand,
                     uint64_t* __restrict__ keys, int64_t capacity, int32_t* __restrict__ counts, int block_append) {
    pdl_launch_dependents();          // head of the chain: the select / sort kernel may be scheduled while this grid drains
    extern __shared__ float part[];          // [rows_per_block][stride] per-vector maxima; stride is odd: conflict-free
    // per vector (VEC == 4): bits 0-1 = index of the FIRST element attaining the vector's maximum, bit 2 = another element of
    // the vector comes within reach of it (could tie it in probability) -- lets phase 2 skip the re-read of the winning vector
    unsigned char* vinfo = reinterpret_cast<unsigned char*>(part + (size_t)rows_per_block * stride);

    const int j = blockIdx.y;
    const int chunk = blockIdx.x;
    const int64_t a0 = (int64_t)chunk * rows_per_block;
    const int nrows = (int)min64(rows_per_block, A - a0);
    const int ppr = (C + VEC - 1) / VEC;
    const int nvec = nrows * ppr;
    const float* base = cls + ((int64_t)j * A + a0) * C;

    // ---- phase 1: stream the [nrows, C] tile, keep each vector's maximum ----
    {
        // (row, k) of this thread's vector advance by constants per step of kFilterThreads vectors
        const int drow = kFilterThreads / ppr, dk = kFilterThreads - drow * ppr;
        int row = threadIdx.x / ppr, k = threadIdx.x - row * ppr;
        if (VEC == 4) {
            const float4* src = reinterpret_cast<const float4*>(base);
            for (int v0 = threadIdx.x; v0 < nvec; v0 += kFilterThreads * 4) {
                float4 x[4];
                int idx[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int v = v0 + u * kFilterThreads;
                    idx[u] = row * stride + k;
                    k += dk;
                    row += drow;
                    if (k >= ppr) {
                        k -= ppr;
                        row += 1;
                    }
                    if (v < nvec) x[u] = ld_stream_f4(src + v);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int v = v0 + u * kFilterThreads;
                    if (v < nvec) {
                        const float hi1 = fmaxf(x[u].x, x[u].y), hi2 = fmaxf(x[u].z, x[u].w);
                        const float m = fmaxf(hi1, hi2);
                        part[idx[u]] = m;
                        if (m > prefilter) {
                            // only a vector whose maximum can pass the threshold can hold a detection's class (a trained model:
                            // ~1 vector in 3000, so the usual cost is this one compare): record where its maximum sits
                            const float lo1 = fminf(x[u].x, x[u].y), lo2 = fminf(x[u].z, x[u].w);
                            const float s2 = fmaxf(fminf(hi1, hi2), fmaxf(lo1, lo2));      // second largest of the four
                            // first index attaining the maximum (ties go to the lower index, like torch.max)
                            const int first = (hi1 >= hi2) ? ((x[u].x >= x[u].y) ? 0 : 1) : ((x[u].z >= x[u].w) ? 2 : 3);
                            // logits: a runner-up within 0.05 (or both saturated) may tie the maximum's fp32 sigmoid;
                            // probabilities: equal values tie and the first index already wins
                            const bool reach = is_logits && (s2 >= ((m > 10.05f) ? 10.0f : m - 0.05f));
                            vinfo[idx[u]] = (unsigned char)(first | (reach ? 4 : 0));
                        }
                    }
                }
            }
        } else {
            for (int v = threadIdx.x; v < nvec; v += kFilterThreads) {
                part[row * stride + k] = base[v];
                k += dk;
                row += drow;
                if (k >= ppr) {
                    k -= ppr;
                    row += 1;
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 2: one thread per anchor row ----
    bool is_cand = false;
    float best = 0.0f;
    int best_c = 0;
    const int r = threadIdx.x;
    if (r < nrows) {
        const float* p = part + r * stride;
        float m = p[0], m2 = -INFINITY;      // largest and second largest per-vector maximum
        int kmax = 0;
        for (int k = 1; k < ppr; ++k) {
            const float v = p[k];
            if (v > m) {
                m2 = m;
                m = v;
                kmax = k;
            } else {
                m2 = fmaxf(m2, v);
            }
        }
        if (m > prefilter) {
            // Which raw values can attain the row's maximum PROBABILITY?  For probabilities: only values == m.
            // For logits: sigmoid is evaluated in fp32 and saturates, so every logit within 0.05 of the max (or above 10
            // when the max is) is evaluated exactly; anything lower is smaller by >= 2e-6 relative, far beyond rounding.
            const float t = is_logits ? ((m > 10.05f) ? 10.0f : m - 0.05f) : m;
            const float* row = base + (int64_t)r * C;
            const int info = (VEC == 4) ? (int)vinfo[r * stride + kmax] : 4;
            if (VEC == 4 && !(info & 4) && (is_logits ? (m2 < t) : true)) {
                // the common case: one element can attain the maximum probability (for probabilities: the first of the equal
                // maxima, and kmax is the first vector that holds it) -- known from phase 1, nothing is re-read
                best = is_logits ? sigmoid_exact(m) : m;
                best_c = kmax * VEC + (info & 3);
            } else {
                best = -1.0f;
                // walk every vector that can hold a winner, in class order
                const int k_lo = (m2 >= t) ? 0 : kmax;
                const int k_hi = (m2 >= t) ? ppr : kmax + 1;
                for (int k = k_lo; k < k_hi; ++k) {
                    if (p[k] >= t) {
                        if (VEC == 4 && p[k] > prefilter) {
                            // phase 1 left this vector's record: unless a second element of the vector is within reach of its
                            // maximum, the maximum (p[k], at a known index) is the vector's only contender -- nothing is re-read.
                            // (A vector at or below the prefilter has no record; it cannot hold a CANDIDATE's class either: a
                            // candidate's best logit lies >= 0.01 above it, far beyond a tie of the fp32 sigmoids.)
                            const int vi = (int)vinfo[r * stride + k];
                            if (!(vi & 4)) {
                                const float pr = is_logits ? sigmoid_exact(p[k]) : p[k];
                                if (pr > best) {
                                    best = pr;
                                    best_c = k * VEC + (vi & 3);
                                }
                                continue;
                            }
                        } else if (VEC == 4 && is_logits) {
                            continue;
                        }
#pragma unroll
                        for (int e = 0; e < VEC; ++e) {
                            const int c = k * VEC + e;
                            if (c < C) {
                                const float x = row[c];
                                if (x >= t) {
                                    const float pr = is_logits ? sigmoid_exact(x) : x;
                                    if (pr > best) {          // strict: first maximal index, like torch.max(dim=1)
                                        best = pr;
                                        best_c = c;
                                    }
                                }
                            }
                        }
                    }
                }
            }
            is_cand = best > score_thresh;                // model.py:536  scores > 0.05
        }
    }
    // aggregated append: the atomic is issued BEFORE the survivors' decode so that its round trip to L2 overlaps the anchor /
    // regression loads and the two exponentials
    const unsigned ballot = __ballot_sync(0xffffffffu, is_cand);
    const int lane = threadIdx.x & 31;
    int warp_base = 0;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t an = a0 + r;
    if (block_append) {
        // Block-aggregated form (the default): ONE atomic per block that has candidates.  The N per-image counters share one
        // 128-byte line, so every append of the launch is serialised by one L2 slice.  When every anchor is a candidate (an
        // untrained model) one atomic per WARP is 200 k atomics per launch, as long as the whole stream: the kernel then ran
        // 0.40 or 0.62 ms from one process to the next on the same GPU (where the allocator put the counters); per block it is
        // 0.417 ms every time, for +0.3 % on a trained-like batch (two barriers per block).
        __shared__ int warp_cnt[kFilterThreads / 32];
        __shared__ int block_base;
        const int warp = threadIdx.x >> 5;
        if (lane == 0) warp_cnt[warp] = __popc(ballot);
        __syncthreads();
        int total = 0, before = 0;
#pragma unroll
        for (int w = 0; w < kFilterThreads / 32; ++w) {
            const int c = warp_cnt[w];
            before += (w < warp) ? c : 0;
            total += c;
        }
        if (total == 0) return;                                  // block-uniform
        if (threadIdx.x == 0) block_base = atomicAdd(&counts[j], total);
        if (is_cand) b = decode_clip(anchors[an], reg[(int64_t)j * A + an], img_w, img_h);
        __syncthreads();
        warp_base = block_base + before;
    } else {
        if (ballot != 0u && lane == 0) warp_base = atomicAdd(&counts[j], __popc(ballot));
        if (is_cand) b = decode_clip(anchors[an], reg[(int64_t)j * A + an], img_w, img_h);
        warp_base = __shfl_sync(0xffffffffu, warp_base, 0);
    }
    if (is_cand) {
        const int64_t slot = (int64_t)warp_base + __popc(ballot & ((1u << lane) - 1u));
        if (slot < capacity) {
            // 32-byte record as two 128-bit stores
            float4* dst = reinterpret_cast<float4*>(cand + (int64_t)j * capacity + slot);
            dst[0] = b;
            dst[1] = make_float4(best, __int_as_float(best_c), __int_as_float((int)an), 0.0f);
            keys[(int64_t)j * capacity + slot] = make_key(best, (int)an);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4h: the same filter on the head's RAW conv outputs (eval side): per pyramid level, classification
// [N, 9*C, H_l, W_l] and regression [N, 36, H_l, W_l] as the output convolutions produce them, i.e. before
// ClassificationModel / RegressionModel's permute + contiguous + view (retinanet/model.py:125-130, 170-184) and
// ResNet.forward's torch.cat (model.py:472-474) -- 16 B/element of layout copies in front of a 4 B/element read.
// Block = one level, one anchor type k, kHeadFilterPos consecutive positions; thread = one position.  The C class rows of
// type k are C contiguous runs of positions, so a warp reads 128 contiguous bytes per class and every thread keeps its own
// running (max, first argmax, second max) -- no shared-memory staging.  Rows whose runner-up is within reach of the maximum
// (see decode_filter_kernel) walk their classes once more to apply torch.max's first-index rule on the exact probabilities.
// Candidates carry the anchor index of the reference's concatenated order, so K5/K6 are unchanged.
// ------------------------------------------------------------------------------------------------
constexpr int kHeadFilterPos = 256;
constexpr int kHeadFilterMaxLevels = 8;

struct HeadFilterLevels {
    int n;
    const float* cls[kHeadFilterMaxLevels];
    const float* reg[kHeadFilterMaxLevels];
    int hw[kHeadFilterMaxLevels];
    int64_t anchor_off[kHeadFilterMaxLevels + 1];
    int chunk_off[kHeadFilterMaxLevels + 1];
    int pos_chunks[kHeadFilterMaxLevels];
};

__global__ void __launch_bounds__(kHeadFilterPos)
decode_filter_head_kernel(const HeadFilterLevels lv, int is_logits, const float4* __restrict__ anchors, int64_t A, int C,
                          float img_w, float img_h, float score_thresh, float prefilter, cldet_candidate* __restrict__ cand,
                          uint64_t* __restrict__ keys, int64_t capacity, int32_t* __restrict__ counts) {
    pdl_launch_dependents();
    const int j = blockIdx.y;
    int l = 0;
    while (l + 1 < lv.n && (int)blockIdx.x >= lv.chunk_off[l + 1]) ++l;
    const int hw = lv.hw[l];
    const int local = (int)blockIdx.x - lv.chunk_off[l];
    const int k = local / lv.pos_chunks[l];
    const int pp = (local - k * lv.pos_chunks[l]) * kHeadFilterPos + (int)threadIdx.x;     // position inside the level's plane
    const bool live = pp < hw;

    bool is_cand = false;
    float best = 0.0f;
    int best_c = 0;
    if (live) {
        const float* col = lv.cls[l] + ((int64_t)j * (kAnchorsPerCell * C) + (int64_t)k * C) * hw + pp;    // + c * hw
        float m = -INFINITY, m2 = -INFINITY;
        int cmax = 0;
        int c = 0;
        for (; c + 8 <= C; c += 8) {                       // eight independent loads in flight per thread
            float x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = __ldg(col + (int64_t)(c + u) * hw);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (x[u] > m) {                            // strict: the first maximal class stays
                    m2 = m;
                    m = x[u];
                    cmax = c + u;
                } else {
                    m2 = fmaxf(m2, x[u]);
                }
            }
        }
        for (; c < C; ++c) {
            const float x = __ldg(col + (int64_t)c * hw);
            if (x > m) {
                m2 = m;
                m = x;
                cmax = c;
            } else {
                m2 = fmaxf(m2, x);
            }
        }
        if (m > prefilter) {
            const float t = is_logits ? ((m > 10.05f) ? 10.0f : m - 0.05f) : m;
            if (!is_logits || m2 < t) {
                // only the maximal raw value can attain the maximal probability (ties in the raw value: first index)
                best = is_logits ? sigmoid_exact(m) : m;
                best_c = cmax;
            } else {
                best = -1.0f;
                for (int c2 = 0; c2 < C; ++c2) {
                    const float x = __ldg(col + (int64_t)c2 * hw);
                    if (x >= t) {
                        const float pr = sigmoid_exact(x);
                        if (pr > best) {                   // strict: first maximal index, like torch.max(dim=1)
                            best = pr;
                            best_c = c2;
                        }
                    }
                }
            }
            is_cand = best > score_thresh;                 // model.py:536  scores > 0.05
        }
    }
    // warp-aggregated append; the atomic's round trip overlaps the survivors' decode (see decode_filter_kernel)
    const unsigned ballot = __ballot_sync(0xffffffffu, is_cand);
    const int lane = threadIdx.x & 31;
    int warp_base = 0;
    if (ballot != 0u && lane == 0) warp_base = atomicAdd(&counts[j], __popc(ballot));
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t an = lv.anchor_off[l] + (int64_t)pp * kAnchorsPerCell + k;
    if (is_cand) {
        const float* rp = lv.reg[l] + ((int64_t)j * (kAnchorsPerCell * 4) + k * 4) * hw + pp;
        const float4 d = make_float4(rp[0], rp[hw], rp[2 * (int64_t)hw], rp[3 * (int64_t)hw]);
        b = decode_clip(anchors[an], d, img_w, img_h);
    }
    warp_base = __shfl_sync(0xffffffffu, warp_base, 0);
    if (is_cand) {
        const int64_t slot = (int64_t)warp_base + __popc(ballot & ((1u << lane) - 1u));
        if (slot < capacity) {
            float4* dst = reinterpret_cast<float4*>(cand + (int64_t)j * capacity + slot);
            dst[0] = b;
            dst[1] = make_float4(best, __int_as_float(best_c), __int_as_float((int)an), 0.0f);
            keys[(int64_t)j * capacity + slot] = make_key(best, (int)an);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K5: top-k by radix select on the 32 score bits (3 passes of 11/11/10 bits), then exact ordering of the
// survivors by the full 64-bit key with a tiled rank sort.  Everything is per image; no host sync.
// select state per image: [0] prefix (score bits decided so far), [1] remaining k, [2] survivors written, [3] done flag
// ------------------------------------------------------------------------------------------------
constexpr int kSelBins = 2048;
constexpr int kRankPerBlock = 64;       // candidates ranked by one 256-thread block of rank_sort_kernel (4 threads each)
constexpr int kRankDirect = 2048;      // up to this many candidates per image the O(n^2) rank sort beats radix select + compaction
constexpr int kBucketMax = 2048;       // lists up to this length are ranked by ONE block of rank_sort_kernel through a bucket pass
constexpr int kBuckets = 2048;
constexpr int kRadixMin = 4096;        // longer lists (no top-k) are ordered by the one-launch radix sort instead of the pairwise rank sort

struct SelPass {
    int shift;     // bit position of this digit inside the 32 score bits
    int bits;      // digit width
    int hi_bits;   // number of bits already decided (above this digit)
};

__global__ void select_init_kernel(const int32_t* __restrict__ counts, int64_t capacity, int topk, uint32_t* __restrict__ state,
                                   uint32_t* __restrict__ hist, uint32_t* __restrict__ done, int N) {
    const int j = blockIdx.x;
    for (int b = threadIdx.x; b < kSelBins; b += blockDim.x) hist[(int64_t)j * kSelBins + b] = 0;
    if (threadIdx.x == 0) {
        done[j] = 0;
        const int64_t cnt = min64(counts[j], capacity);
        uint32_t* st = state + 4 * j;
        st[0] = 0;
        st[2] = 0;
        // keep everything: no top-k, fewer candidates than k, or few enough (kRankDirect) that the rank sort can order ALL
        // of them and cut at k itself -- the three histogram/pick passes and the compaction then exit immediately
        if (topk <= 0 || cnt <= topk || cnt <= kRankDirect) {
            st[1] = 0;
            st[3] = 1;
        } else {
            st[1] = (uint32_t)topk;
            st[3] = 0;
        }
    }
}

// One radix pass: per-image histogram of the current digit among the keys that match the prefix decided so far, then the
// image's LAST block to finish (threadfence + counter, as in the loss kernel) picks the digit where the descending cumulative
// count crosses the remaining k, clears the histogram for the next pass and re-arms the counter.  (The pick used to be a
// second launch per pass; in the trained-like regime the stage is a chain of short dependent launches, so each one removed
// is ~5 us.)  In the pick, 256 threads own 8 consecutive bins each (high bins first); a shared-memory suffix scan over the
// 256 partial sums locates the owning thread, which then walks its 8 bins.
__global__ void __launch_bounds__(256)
select_hist_pick_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ counts, int64_t capacity, SelPass ps,
                        uint32_t* __restrict__ state, uint32_t* __restrict__ hist, uint32_t* __restrict__ done) {
    __shared__ uint32_t sh[kSelBins];
    __shared__ uint32_t part[256];
    __shared__ bool is_last;
    const int j = blockIdx.y;
    uint32_t* st = state + 4 * j;
    if (st[3]) return;                                   // everything is kept (block-uniform)
    const int64_t cnt = min64(counts[j], capacity);
    const unsigned int nblk = (unsigned int)min64((int64_t)gridDim.x, (cnt + blockDim.x - 1) / blockDim.x);   // blocks with work
    if (blockIdx.x >= nblk) return;
    const int nb = 1 << ps.bits;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) sh[b] = 0;
    __syncthreads();
    const uint32_t prefix = st[0];
    const uint64_t* k = keys + (int64_t)j * capacity;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t s = (uint32_t)(k[i] >> 32);
        const bool match = ps.hi_bits == 0 || (s >> (32 - ps.hi_bits)) == (prefix >> (32 - ps.hi_bits));
        if (match) atomicAdd(&sh[(s >> ps.shift) & (nb - 1)], 1u);
    }
    __syncthreads();
    uint32_t* h = hist + (int64_t)j * kSelBins;
    for (int b = threadIdx.x; b < nb; b += blockDim.x)
        if (sh[b]) atomicAdd(&h[b], sh[b]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&done[j], 1u) == nblk - 1u);
    __syncthreads();
    if (!is_last) return;

    // ---- the image's last block: pick the digit ----
    __threadfence();
    for (int b = threadIdx.x; b < kSelBins; b += blockDim.x) {
        sh[b] = (b < nb) ? __ldcg(h + b) : 0u;           // written by other blocks' atomics: read through L2
        h[b] = 0;                                        // ready for the next pass
    }
    if (threadIdx.x == 0) done[j] = 0;
    __syncthreads();
    // thread t owns bins [hi-7, hi] with hi = kSelBins-1-8t (descending order)
    const int hi = kSelBins - 1 - 8 * (int)threadIdx.x;
    uint32_t mine = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) mine += sh[hi - q];
    part[threadIdx.x] = mine;
    __syncthreads();
    // inclusive prefix over threads (= suffix over bins), Hillis-Steele in shared memory
    for (int off = 1; off < 256; off <<= 1) {
        const uint32_t add = (threadIdx.x >= (unsigned)off) ? part[threadIdx.x - off] : 0u;
        __syncthreads();
        part[threadIdx.x] += add;
        __syncthreads();
    }
    const uint32_t need0 = st[1];
    const uint32_t incl = part[threadIdx.x];
    const uint32_t excl = incl - mine;
    __syncthreads();
    // the owner is the first thread whose inclusive count reaches `need`; if none does (cannot happen when the image has
    // more candidates than k) the last bin wins
    const bool owner = (excl < need0 && incl >= need0) || (threadIdx.x == 255 && incl < need0);
    if (owner) {
        uint32_t need = need0 - excl;
        int b = hi;
        for (int q = 0; q < 8; ++q, --b) {
            if (sh[b] >= need || b == 0) break;
            need -= sh[b];
        }
        if (b < 0) b = 0;
        st[0] |= (uint32_t)b << ps.shift;
        st[1] = need;                   // how many to take from inside this bucket
    }
}

// The whole top-k SELECT of one image in ONE launch: state init, the three radix passes on the score bits and the compaction
// of the survivors, by a thread-block CLUSTER of kSelCluster CTAs per image (co-scheduled by the hardware, synchronised with
// the cluster barrier, histograms exchanged through distributed shared memory -- no global scratch, no inter-launch gaps).
// An image that keeps everything (no top-k, fewer candidates than k, or few enough for the rank sort to order them all) costs
// one cluster that returns at once: the trained-model regime used to pay five launches that did nothing (init + 3 passes +
// compaction, ~20 us of a 400 us pipeline).  Each CTA streams its share of the keys from L2 with four independent loads per
// thread; CTA 0 adds the cluster's histograms and picks the digit with the same descending cumulative-count walk as
// select_hist_pick_kernel, then publishes (prefix, remaining k) into every CTA's shared memory.
constexpr int kSelThreads = 512;       // two CTAs per SM: 32 images x 8 CTAs are one wave on 148 SMs
constexpr int kSelCluster = 8;
constexpr int kSelUnroll = 8;          // independent key loads in flight per thread
constexpr int kSelSub = 4;             // private histogram copies per CTA (by lane & 3)

__global__ void __cluster_dims__(kSelCluster, 1, 1) __launch_bounds__(kSelThreads)
select_fused_kernel(const cldet_candidate* __restrict__ cand, const uint64_t* __restrict__ keys, const int32_t* __restrict__ counts,
                    int64_t capacity, int topk, uint32_t* __restrict__ state, cldet_candidate* __restrict__ out_cand,
                    uint64_t* __restrict__ out_keys, int64_t out_capacity) {
    pdl_wait();                       // may be scheduled while the previous kernel of the chain drains
    pdl_launch_dependents();
    namespace cg = cooperative_groups;
    // kSelSub private copies of the histogram, chosen by lane: the high score bits of a whole image fall into a handful of bins
    // (sign, exponent, two mantissa bits), so without them most of a warp's 32 atomics serialise on the same few words
    __shared__ uint32_t shs[kSelSub][kSelBins];
    uint32_t* const sh = shs[0];
    __shared__ uint32_t part[256];
    __shared__ uint32_t s_prefix, s_need;
    __shared__ int warp_tot[kSelThreads / 32];
    __shared__ int block_base;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int j = blockIdx.x / kSelCluster;
    const int tid = threadIdx.x;
    uint32_t* st = state + 4 * j;
    const int64_t cnt = min64(counts[j], capacity);
    if (topk <= 0 || cnt <= topk || cnt <= kRankDirect) {       // keep everything (uniform over the cluster)
        if (tid == 0 && rank == 0) {
            st[0] = 0; st[1] = 0; st[2] = 0; st[3] = 1;
        }
        return;
    }
    if (tid == 0) {
        s_prefix = 0;
        s_need = (uint32_t)topk;
        if (rank == 0) st[2] = 0;                               // survivors written (global, atomically advanced below)
    }
    const uint64_t* k = keys + (int64_t)j * capacity;
    const int64_t stride = (int64_t)kSelCluster * kSelUnroll * kSelThreads;
    const SelPass passes[3] = {{21, 11, 0}, {10, 11, 11}, {0, 10, 22}};
#pragma unroll 1
    for (int ps = 0; ps < 3; ++ps) {
        const SelPass P = passes[ps];
        const int nb = 1 << P.bits;
        for (int b = tid; b < kSelSub * kSelBins; b += kSelThreads) shs[0][b] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        uint32_t* const mysh = shs[tid & (kSelSub - 1)];
        for (int64_t i0 = (int64_t)rank * kSelUnroll * kSelThreads; i0 < cnt; i0 += stride) {
            uint32_t sc[kSelUnroll];
#pragma unroll
            for (int u = 0; u < kSelUnroll; ++u) {
                const int64_t i = i0 + u * kSelThreads + tid;
                sc[u] = (i < cnt) ? (uint32_t)(k[i] >> 32) : 0u;
            }
#pragma unroll
            for (int u = 0; u < kSelUnroll; ++u) {
                const int64_t i = i0 + u * kSelThreads + tid;
                const bool match = P.hi_bits == 0 || (sc[u] >> (32 - P.hi_bits)) == (prefix >> (32 - P.hi_bits));
                if (i < cnt && match) atomicAdd(&mysh[(sc[u] >> P.shift) & (nb - 1)], 1u);
            }
        }
        __syncthreads();
        for (int b = tid; b < nb; b += kSelThreads) {            // fold the private copies into copy 0
            uint32_t v = shs[0][b];
#pragma unroll
            for (int q = 1; q < kSelSub; ++q) v += shs[q][b];
            shs[0][b] = v;
        }
        cluster.sync();                                          // every CTA's histogram is complete
        if (rank == 0) {
            // add the other CTAs' histograms (distributed shared memory) into mine
            for (int b = tid; b < nb; b += kSelThreads) {
                uint32_t v = sh[b];
#pragma unroll
                for (int r = 1; r < kSelCluster; ++r) v += cluster.map_shared_rank(sh, r)[b];
                sh[b] = v;
            }
            __syncthreads();
            // pick: thread t < 256 owns bins [hi-7, hi], hi = kSelBins-1-8t (descending); suffix scan over the 256 partial sums
            uint32_t mine = 0;
            const int hi = kSelBins - 1 - 8 * (tid & 255);
            if (tid < 256) {
#pragma unroll
                for (int q = 0; q < 8; ++q) mine += sh[hi - q];
                part[tid] = mine;
            }
            __syncthreads();
            for (int off = 1; off < 256; off <<= 1) {
                uint32_t add = 0;
                if (tid < 256 && tid >= off) add = part[tid - off];
                __syncthreads();
                if (tid < 256) part[tid] += add;
                __syncthreads();
            }
            const uint32_t need0 = s_need;
            __syncthreads();
            if (tid < 256) {
                const uint32_t incl = part[tid], excl = incl - mine;
                const bool owner = (excl < need0 && incl >= need0) || (tid == 255 && incl < need0);
                if (owner) {
                    uint32_t need = need0 - excl;
                    int b = hi;
                    for (int q = 0; q < 8; ++q, --b) {
                        if (sh[b] >= need || b == 0) break;
                        need -= sh[b];
                    }
                    if (b < 0) b = 0;
                    const uint32_t np = prefix | ((uint32_t)b << P.shift);
                    for (int r = 0; r < kSelCluster; ++r) {      // publish to every CTA of the cluster
                        *cluster.map_shared_rank(&s_prefix, r) = np;
                        *cluster.map_shared_rank(&s_need, r) = need;
                    }
                }
            }
        }
        cluster.sync();                                          // (prefix, need) visible everywhere; remote reads of sh are done
    }
    // compaction: every candidate whose score bits reach the k-th score (exact ties included; the ordering pass cuts at k);
    // order is arbitrary (the rank sort orders), so a tile appends with one global atomic
    const uint32_t thr = s_prefix;
    const int lane = tid & 31, warp = tid >> 5;
    for (int64_t i0 = (int64_t)rank * kSelUnroll * kSelThreads; i0 < cnt; i0 += stride) {
        uint64_t key[kSelUnroll];
        bool take[kSelUnroll];
        int mine = 0;
#pragma unroll
        for (int u = 0; u < kSelUnroll; ++u) {
            const int64_t i = i0 + u * kSelThreads + tid;
            key[u] = (i < cnt) ? k[i] : 0ull;
            take[u] = (i < cnt) && (uint32_t)(key[u] >> 32) >= thr;
            mine += take[u] ? 1 : 0;
        }
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < kSelThreads / 32; ++w) {
                const int c = warp_tot[w];
                warp_tot[w] = tot;
                tot += c;
            }
            block_base = tot ? (int)atomicAdd(&st[2], (uint32_t)tot) : 0;
        }
        __syncthreads();
        int64_t slot = (int64_t)block_base + warp_tot[warp] + (incl - mine);
#pragma unroll
        for (int u = 0; u < kSelUnroll; ++u) {
            if (take[u]) {
                const int64_t i = i0 + u * kSelThreads + tid;
                if (slot < out_capacity) {
                    const float4* src = reinterpret_cast<const float4*>(cand + (int64_t)j * capacity + i);
                    float4* dst = reinterpret_cast<float4*>(out_cand + (int64_t)j * out_capacity + slot);
                    dst[0] = src[0];
                    dst[1] = src[1];
                    out_keys[(int64_t)j * out_capacity + slot] = key[u];
                }
                ++slot;
            }
        }
        __syncthreads();
    }
    if (tid == 0 && rank == 0) {
        st[0] = thr; st[1] = 0; st[3] = 0;
    }
}

// survivors: score bits > threshold always; == threshold: all of them (exact ties are cut after the ordering pass).
// 4 candidates per thread, one atomic per block.
constexpr int kCompactPerThread = 4;
__global__ void __launch_bounds__(256)
select_compact_kernel(const cldet_candidate* __restrict__ cand, const uint64_t* __restrict__ keys,
                      const int32_t* __restrict__ counts, int64_t capacity, uint32_t* __restrict__ state,
                      cldet_candidate* __restrict__ out_cand, uint64_t* __restrict__ out_keys, int64_t out_capacity) {
    __shared__ int warp_tot[8];
    __shared__ int block_base;
    const int j = blockIdx.y;
    uint32_t* st = state + 4 * j;
    const int64_t cnt = min64(counts[j], capacity);
    const int64_t i0 = (int64_t)blockIdx.x * (256 * kCompactPerThread);
    if (i0 >= cnt) return;
    if (st[3] != 0) return;              // everything is kept: the rank sort reads the original arrays
    const bool all = false;
    const uint32_t thr = st[0];
    uint64_t key[kCompactPerThread];
    bool take[kCompactPerThread];
    int mine = 0;
#pragma unroll
    for (int u = 0; u < kCompactPerThread; ++u) {
        const int64_t i = i0 + u * 256 + threadIdx.x;
        key[u] = (i < cnt) ? keys[(int64_t)j * capacity + i] : 0ull;
        take[u] = (i < cnt) && (all || (uint32_t)(key[u] >> 32) >= thr);
        mine += take[u] ? 1 : 0;
    }
    // exclusive prefix of `mine` over the block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) {
            const int c = warp_tot[w];
            warp_tot[w] = tot;
            tot += c;
        }
        block_base = tot ? (int)atomicAdd(&st[2], (uint32_t)tot) : 0;
    }
    __syncthreads();
    int64_t slot = (int64_t)block_base + warp_tot[warp] + (incl - mine);
#pragma unroll
    for (int u = 0; u < kCompactPerThread; ++u) {
        if (take[u]) {
            const int64_t i = i0 + u * 256 + threadIdx.x;
            if (slot < out_capacity) {
                const float4* src = reinterpret_cast<const float4*>(cand + (int64_t)j * capacity + i);
                float4* dst = reinterpret_cast<float4*>(out_cand + (int64_t)j * out_capacity + slot);
                dst[0] = src[0];
                dst[1] = src[1];
                out_keys[(int64_t)j * out_capacity + slot] = key[u];
            }
            ++slot;
        }
    }
}

// Tiled rank sort: rank_i = #{ l : key_l > key_i } (keys are distinct).  Writes candidate i to position rank_i when
// rank_i < limit (top-k cut), and the final count min(n, limit).  Blocks stride over the candidates, so the grid can be
// sized for the expected count and still be correct for any count.
__global__ void __launch_bounds__(256)
rank_sort_kernel(const cldet_candidate* __restrict__ cand, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ state,
                 const int32_t* __restrict__ counts_in, int64_t in_capacity, int topk, cldet_candidate* __restrict__ sorted,
                 int64_t out_capacity, int32_t* __restrict__ sorted_counts, const cldet_candidate* __restrict__ orig_cand = nullptr,
                 const uint64_t* __restrict__ orig_keys = nullptr, int64_t orig_capacity = 0, int64_t max_n = 0, int64_t bucket_n = 0) {
    pdl_wait();                       // may be scheduled while the previous kernel of the chain drains
    pdl_launch_dependents();
    __shared__ uint64_t tile[kBucketMax];       // pairwise path: a tile of 1024 keys; bucket path: the keys in bucket order
    const int j = blockIdx.y;
    int64_t n = state ? (int64_t)state[4 * j + 2] : (int64_t)counts_in[j];
    if (state && state[4 * j + 3] && orig_cand) {
        // this image kept every candidate (select_init_kernel): nothing was compacted, order the original arrays
        cand = orig_cand;
        keys = orig_keys;
        in_capacity = orig_capacity;
        n = (int64_t)counts_in[j];
    }
    n = min64(n, in_capacity);
    if (max_n > 0 && n > max_n) return;      // a long list: the radix sort kernel launched next to this one orders it
    if (topk <= 0 && n > out_capacity) {     // no top-k and more candidates than the sorted list can hold: report an EMPTY list;
        if (blockIdx.x == 0 && threadIdx.x == 0) sorted_counts[j] = 0;      // the caller sees counts[j] > capacity and re-runs
        return;
    }
    const int64_t limit = (topk > 0) ? min64(n, topk) : n;
    if (blockIdx.x == 0 && threadIdx.x == 0) sorted_counts[j] = (int32_t)min64(limit, out_capacity);
    const uint64_t* k = keys + (int64_t)j * in_capacity;
    if (n <= bucket_n) {
        // Short list (the trained-model regime, with or without a top-k): ONE block ranks the whole image through a bucket pass
        // instead of n^2 64-bit compares spread over the grid (23 us for 32 x ~1300 candidates, issue-bound on every SM -- the
        // longest kernel of the post-filter chain).  bucket = (key - min) >> shift with shift chosen so that the image's key
        // range spans kBuckets buckets; rank = (keys in higher buckets) + (larger keys of the own bucket).  Keys are distinct,
        // so this is the pairwise rank exactly; only the own bucket (a few keys, unless scores tie en masse) is compared.
        if (blockIdx.x != 0) return;
        __shared__ uint32_t bcnt[kBuckets], bstart[kBuckets];
        __shared__ uint64_t red_lo[8], red_hi[8];
        __shared__ uint32_t scan_s[8];
        constexpr int kPer = kBucketMax / 256;
        constexpr int kOwn = kBuckets / 256;
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int nn = (int)n;
        uint64_t kreg[kPer];
        uint64_t lo = ~0ull, hi = 0ull;
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int i = tid + 256 * u;
            kreg[u] = (i < nn) ? k[i] : 0ull;
            if (i < nn) {
                lo = (kreg[u] < lo) ? kreg[u] : lo;
                hi = (kreg[u] > hi) ? kreg[u] : hi;
            }
        }
        for (int b = tid; b < kBuckets; b += 256) bcnt[b] = 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint64_t ol = __shfl_xor_sync(0xffffffffu, lo, o), oh = __shfl_xor_sync(0xffffffffu, hi, o);
            lo = (ol < lo) ? ol : lo;
            hi = (oh > hi) ? oh : hi;
        }
        if (lane == 0) {
            red_lo[warp] = lo;
            red_hi[warp] = hi;
        }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            lo = (red_lo[w] < lo) ? red_lo[w] : lo;
            hi = (red_hi[w] > hi) ? red_hi[w] : hi;
        }
        const uint64_t range = (hi >= lo) ? hi - lo : 0ull;                   // nn == 0: every loop below is empty
        const int shift = (range < (uint64_t)kBuckets) ? 0 : (64 - __clzll((long long)range)) - 11;      // range >> shift < 2048
        int bkt[kPer], loc[kPer];
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int i = tid + 256 * u;
            bkt[u] = (i < nn) ? (int)((kreg[u] - lo) >> shift) : 0;
            loc[u] = (i < nn) ? (int)atomicAdd(&bcnt[bkt[u]], 1u) : 0;       // arrival order inside the bucket (arbitrary)
        }
        __syncthreads();
        // bstart[b] = number of keys in HIGHER buckets (descending order): thread t owns buckets kOwn*t .. kOwn*t + kOwn-1
        uint32_t mine[kOwn], tot = 0;
#pragma unroll
        for (int q = 0; q < kOwn; ++q) {
            mine[q] = bcnt[tid * kOwn + q];
            tot += mine[q];
        }
        uint32_t incl = tot;                                                  // inclusive SUFFIX sum over the warp's threads
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_down_sync(0xffffffffu, incl, o);
            if (lane + o < 32) incl += v;
        }
        if (lane == 0) scan_s[warp] = incl;
        __syncthreads();
        uint32_t above = incl - tot;                                          // keys owned by later threads of this warp
        for (int w = warp + 1; w < 8; ++w) above += scan_s[w];
#pragma unroll
        for (int q = kOwn - 1; q >= 0; --q) {
            bstart[tid * kOwn + q] = above;
            above += mine[q];
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int i = tid + 256 * u;
            if (i < nn) {
                const int pos = (int)bstart[bkt[u]] + loc[u];
                tile[pos] = kreg[u];
            }
        }
        __syncthreads();
        const float4* src = reinterpret_cast<const float4*>(cand + (int64_t)j * in_capacity);
        float4* dst = reinterpret_cast<float4*>(sorted + (int64_t)j * out_capacity);
        int rk[kPer];
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int i = tid + 256 * u;
            rk[u] = 0x7fffffff;
            if (i < nn) {
                const int b0 = (int)bstart[bkt[u]], b1 = b0 + (int)bcnt[bkt[u]];
                int r = b0;
                for (int q = b0; q < b1; ++q) r += (tile[q] > kreg[u]) ? 1 : 0;
                rk[u] = r;
            }
        }
        // the 32-byte records: all loads of a half in flight before the first store (one L2 round trip per half, not per record)
#pragma unroll
        for (int h = 0; h < kPer; h += 4) {
            float4 a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = tid + 256 * (h + u);
                const bool on = rk[h + u] < limit && rk[h + u] < out_capacity;
                a[u] = on ? src[2 * (int64_t)i] : make_float4(0.f, 0.f, 0.f, 0.f);
                b[u] = on ? src[2 * (int64_t)i + 1] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (rk[h + u] < limit && rk[h + u] < out_capacity) {
                    dst[2 * (int64_t)rk[h + u]] = a[u];
                    dst[2 * (int64_t)rk[h + u] + 1] = b[u];
                }
            }
        }
        return;
    }
    // a block ranks kRankPerBlock candidates; the four threads of a candidate count a quarter of every tile each
    __shared__ int rank_s[kRankPerBlock];

    const int c = threadIdx.x & (kRankPerBlock - 1), q = threadIdx.x / kRankPerBlock;
    for (int64_t i0 = (int64_t)blockIdx.x * kRankPerBlock; i0 < n; i0 += (int64_t)gridDim.x * kRankPerBlock) {   // block-uniform
        const int64_t i = i0 + c;
        const uint64_t mine = (i < n) ? k[i] : ~0ull;
        if (q == 0) rank_s[c] = 0;
        int r = 0;
        for (int64_t t0 = 0; t0 < n; t0 += 1024) {
            const int m = (int)min64(1024, n - t0);
            __syncthreads();
            for (int t = threadIdx.x; t < m; t += blockDim.x) tile[t] = k[t0 + t];
            __syncthreads();
            const int per = (m + 3) >> 2;
            const int lo = q * per, hi = min(m, lo + per);
#pragma unroll 8
            for (int t = lo; t < hi; ++t) r += (tile[t] > mine) ? 1 : 0;
        }
        atomicAdd(&rank_s[c], r);
        __syncthreads();
        const int rank = rank_s[c];
        if (q == 0 && i < n && rank < limit && rank < out_capacity) {
            const float4* src = reinterpret_cast<const float4*>(cand + (int64_t)j * in_capacity + i);
            float4* dst = reinterpret_cast<float4*>(sorted + (int64_t)j * out_capacity + rank);
            dst[0] = src[0];
            dst[1] = src[1];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// K5 (large lists): segmented LSD radix sort of each image's candidates by the full 64-bit key, DESCENDING -- the order the
// rank sort produces, without its O(n^2) pair count.  One CTA per image runs ALL passes in one launch: the image's keys live
// in L2, 8-bit digits, stable by construction (tiles are processed in order; inside a tile the 32 warps own consecutive
// 128-key runs, a warp ranks its keys with match.any and per-(warp, digit) counters, an exclusive scan over (digit, warp)
// turns the counters into tile-local offsets).  The digit histograms of all 8 passes come from ONE sweep over the keys;
// a pass whose digit is constant over the image (the high score bits, the unused anchor bits) is skipped -- typically 6-7 of
// the 8 passes run.  The last pass's index permutation gathers the 32-byte records straight into the sorted list.
// ------------------------------------------------------------------------------------------------
constexpr int kRsThreads = 1024;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 4;
constexpr int kRsTile = kRsThreads * kRsItems;
constexpr int kRsBins = 256;

__global__ void __launch_bounds__(kRsThreads)
radix_sort_kernel(const cldet_candidate* __restrict__ cand, const uint64_t* __restrict__ keys, const int32_t* __restrict__ counts,
                  int64_t in_capacity, int64_t max_count, cldet_candidate* __restrict__ sorted, int64_t out_capacity,
                  int32_t* __restrict__ sorted_counts, uint64_t* __restrict__ kbuf, uint32_t* __restrict__ ibuf, int64_t min_n = 0) {
    pdl_wait();                       // may be scheduled while the previous kernel of the chain drains
    pdl_launch_dependents();
    __shared__ uint32_t hist[8][kRsBins];
    __shared__ uint32_t cnt[kRsWarps][kRsBins];
    __shared__ uint32_t base[kRsBins];
    __shared__ uint32_t tile_tot[kRsBins];
    __shared__ int skip[8];
    const int j = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (min_n > 0 && min64(counts[j], in_capacity) <= min_n) return;      // a short list: the rank sort launched before this kernel ordered it
    if (counts[j] > max_count) {            // more candidates than the buffers were sized for: report an EMPTY list; the caller
        if (threadIdx.x == 0) sorted_counts[j] = 0;      // sees counts[j] > capacity and repeats the call at the exact size
        return;
    }
    const int n = (int)min64(min64(counts[j], in_capacity), max_count);
    const uint64_t* k_in = keys + (int64_t)j * in_capacity;
    uint64_t* kb[2] = {kbuf + (int64_t)j * 2 * max_count, kbuf + (int64_t)j * 2 * max_count + max_count};
    uint32_t* ib[2] = {ibuf + (int64_t)j * 2 * max_count, ibuf + (int64_t)j * 2 * max_count + max_count};
    if (tid == 0) sorted_counts[j] = (int32_t)min64(n, out_capacity);
    if (n == 0) return;

    // ---- one sweep: the digit histograms of all 8 passes ----
    for (int i = tid; i < 8 * kRsBins; i += kRsThreads) (&hist[0][0])[i] = 0u;
    if (tid < 8) skip[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kRsThreads) {
        const uint64_t k = k_in[i];
#pragma unroll
        for (int p = 0; p < 8; ++p) atomicAdd(&hist[p][(uint32_t)(k >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    for (int i = tid; i < 8 * kRsBins; i += kRsThreads)
        if ((&hist[0][0])[i] == (uint32_t)n) skip[i / kRsBins] = 1;       // the whole image shares this digit: identity pass
    __syncthreads();

    int cur = -1;                                  // -1: the data still sits in the input arrays (index = position)
    for (int p = 0; p < 8; ++p) {
        if (skip[p]) continue;                     // block-uniform
        // descending order: bin d starts after all larger digits
        if (tid < kRsBins) base[tid] = hist[p][tid];
        __syncthreads();
        for (int off = 1; off < kRsBins; off <<= 1) {          // inclusive suffix sum (Hillis-Steele)
            uint32_t add = 0;
            if (tid < kRsBins && tid + off < kRsBins) add = base[tid + off];
            __syncthreads();
            if (tid < kRsBins) base[tid] += add;
            __syncthreads();
        }
        if (tid < kRsBins) base[tid] -= hist[p][tid];           // exclusive
        __syncthreads();
        const uint64_t* src_k = (cur < 0) ? k_in : kb[cur];
        const uint32_t* src_i = (cur < 0) ? nullptr : ib[cur];
        const int nxt = (cur < 0) ? 0 : (cur ^ 1);
        uint64_t* dst_k = kb[nxt];
        uint32_t* dst_i = ib[nxt];
        const int shift = 8 * p;
        for (int t0 = 0; t0 < n; t0 += kRsTile) {
            for (int i = tid; i < kRsWarps * kRsBins; i += kRsThreads) (&cnt[0][0])[i] = 0u;
            __syncthreads();
            uint64_t key[kRsItems];
            uint32_t idx[kRsItems], loc[kRsItems];
            int dig[kRsItems];
#pragma unroll
            for (int r = 0; r < kRsItems; ++r) {
                const int i = t0 + warp * (32 * kRsItems) + r * 32 + lane;
                const bool valid = i < n;
                key[r] = valid ? src_k[i] : 0ull;
                idx[r] = valid ? (src_i ? src_i[i] : (uint32_t)i) : 0u;
                dig[r] = valid ? (int)((uint32_t)(key[r] >> shift) & 255u) : -1;
            }
#pragma unroll
            for (int r = 0; r < kRsItems; ++r) {
                const unsigned m = __match_any_sync(0xffffffffu, dig[r]);
                const int leader = __ffs(m) - 1;
                uint32_t prev = 0;
                if (lane == leader && dig[r] >= 0) {
                    prev = cnt[warp][dig[r]];
                    cnt[warp][dig[r]] = prev + (uint32_t)__popc(m);
                }
                prev = __shfl_sync(0xffffffffu, prev, leader);
                loc[r] = prev + (uint32_t)__popc(m & ((1u << lane) - 1u));
                __syncwarp();
            }
            __syncthreads();
            if (tid < kRsBins) {
                uint32_t run = 0;
#pragma unroll 8
                for (int w = 0; w < kRsWarps; ++w) {
                    const uint32_t t = cnt[w][tid];
                    cnt[w][tid] = run;
                    run += t;
                }
                tile_tot[tid] = run;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < kRsItems; ++r) {
                if (dig[r] >= 0) {
                    const uint32_t pos = base[dig[r]] + cnt[warp][dig[r]] + loc[r];
                    dst_k[pos] = key[r];
                    dst_i[pos] = idx[r];
                }
            }
            __syncthreads();
            if (tid < kRsBins) base[tid] += tile_tot[tid];
            // (the next tile's first barrier orders this update before any use)
        }
        __syncthreads();          // this pass's global writes are visible to the whole block before the next pass reads them
        cur = nxt;
    }
    // ---- gather the records in sorted order ----
    const int lim = (int)min64(n, out_capacity);
    const float4* src = reinterpret_cast<const float4*>(cand + (int64_t)j * in_capacity);
    float4* dst = reinterpret_cast<float4*>(sorted + (int64_t)j * out_capacity);
    for (int i = tid; i < lim; i += kRsThreads) {
        const uint32_t from = (cur < 0) ? (uint32_t)i : ib[cur][i];
        dst[2 * (int64_t)i] = src[2 * (int64_t)from];
        dst[2 * (int64_t)i + 1] = src[2 * (int64_t)from + 1];
    }
}

// ------------------------------------------------------------------------------------------------
// K6: NMS.  (1) per-image max coordinate + mode (2) 64x64 suppression bitmask over the SORTED list
// (3) sequential resolve, one block per image, 64 boxes per step.
// nms_info per image: [0] max coordinate (float bits), [1] mode actually used (1 trick, 2 vanilla) | kNmsLabelsDisjoint
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kNmsLabelsDisjoint = 0x100u;
constexpr int kChainRounds = 8;             // parallel decision rounds per 64-box chunk of the greedy chain before the box-by-box fallback

constexpr int kPrepThreads = 1024;
constexpr int kPrepUnroll = 4;             // independent record loads in flight per thread: one block walks a whole image's list

__global__ void __launch_bounds__(kPrepThreads)
nms_prepare_kernel(const cldet_candidate* __restrict__ sorted, const int32_t* __restrict__ counts, int64_t capacity, int mode,
                   int64_t vanilla_numel_limit, uint32_t* __restrict__ info) {
    pdl_wait();                       // may be scheduled while the previous kernel of the chain drains
    pdl_launch_dependents();
    __shared__ float red[kPrepThreads / 32], red_min[kPrepThreads / 32];
    __shared__ int red_lab[kPrepThreads / 32];
    const int j = blockIdx.x;
    const int n = (int)min64(counts[j], capacity);
    float m = -INFINITY, mn = INFINITY;
    int lab_max = 0, lab_min = 0;
    const cldet_candidate* c = sorted + (int64_t)j * capacity;
    // (a 256-thread block with one dependent load per iteration took 21 us for 8 k boxes: 32 L2 round trips in a row)
    for (int i0 = threadIdx.x; i0 < n; i0 += kPrepThreads * kPrepUnroll) {
        float4 bx[kPrepUnroll];
        int lb[kPrepUnroll];
#pragma unroll
        for (int u = 0; u < kPrepUnroll; ++u) {
            const int i = i0 + u * kPrepThreads;
            if (i < n) {
                const float4* rec = reinterpret_cast<const float4*>(c + i);
                bx[u] = rec[0];
                lb[u] = __float_as_int(rec[1].y);
            }
        }
#pragma unroll
        for (int u = 0; u < kPrepUnroll; ++u) {
            if (i0 + u * kPrepThreads < n) {
                m = fmaxf(m, fmaxf(fmaxf(bx[u].x, bx[u].y), fmaxf(bx[u].z, bx[u].w)));
                mn = fminf(mn, fminf(fminf(bx[u].x, bx[u].y), fminf(bx[u].z, bx[u].w)));
                lab_max = max(lab_max, lb[u]);
                lab_min = min(lab_min, lb[u]);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        lab_max = max(lab_max, __shfl_xor_sync(0xffffffffu, lab_max, o));
        lab_min = min(lab_min, __shfl_xor_sync(0xffffffffu, lab_min, o));
    }
    if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5] = m;
        red_min[threadIdx.x >> 5] = mn;
        red_lab[threadIdx.x >> 5] = (lab_min < 0) ? 0x7fffffff : lab_max;      // a negative label disables the shortcut below
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kPrepThreads / 32; ++w) {
            m = fmaxf(m, red[w]);
            mn = fminf(mn, red_min[w]);
        }
        int lmax = red_lab[0];
        for (int w = 1; w < kPrepThreads / 32; ++w) lmax = max(lmax, red_lab[w]);
        info[2 * j] = __float_as_uint(m);
        int used = mode;
        if (mode == 0) used = ((int64_t)n * 4 > vanilla_numel_limit) ? 2 : 1;   // torchvision ops/boxes.py batched_nms
        // Coordinate trick, provably disjoint classes: with every coordinate in [0, M] and (label_max + 1) * (M + 1) <= 2^21 the
        // fp32 offsets label * (M + 1) and the shifted coordinates are rounded by at most 0.125 each, so boxes of different
        // labels stay >= 0.5 apart in x: their intersection width is negative and torchvision's kernel never suppresses across
        // labels (for thr >= 0).  The mask kernel may then skip such pairs on the label test alone.  (With huge offsets the
        // trick DOES let rounding merge classes; those inputs keep the full test.)
        const bool disjoint = (used == 1) && n > 0 && mn >= 0.0f && lmax < 0x7fffffff &&
                              ((double)lmax + 1.0) * ((double)m + 1.0) <= 2097152.0;
        info[2 * j + 1] = (uint32_t)used | (disjoint ? kNmsLabelsDisjoint : 0u);
    }
}

__device__ __forceinline__ bool suppresses(const float4 a, float area_a, const float4 b, float area_b, float thr) {
    // torchvision nms: inter / ((area_a + area_b) - inter) > thr, all fp32, one rounding per op.
    // Empty intersections (the vast majority of pairs) give IoU +0 (or NaN), never > thr for thr >= 0: exit before the divide.
    const float w = fminf(a.z, b.z) - fmaxf(a.x, b.x);
    const float h = fminf(a.w, b.w) - fmaxf(a.y, b.y);
    if (!(w > 0.0f) || !(h > 0.0f)) {
        if (thr >= 0.0f) return false;
        const float inter0 = fmaxf(w, 0.0f) * fmaxf(h, 0.0f);
        return (inter0 / ((area_a + area_b) - inter0)) > thr;
    }
    const float inter = w * h;
    return __fdiv_rn(inter, (area_a + area_b) - inter) > thr;
}

__global__ void __launch_bounds__(256)
nms_mask_kernel(const cldet_candidate* __restrict__ sorted, const int32_t* __restrict__ counts, int64_t capacity,
                const uint32_t* __restrict__ info, float thr, uint64_t* __restrict__ mask, int64_t mask_stride_img,
                int col_blocks_alloc) {
    pdl_wait();                       // may be scheduled while the previous kernel of the chain drains
    pdl_launch_dependents();
    const int j = blockIdx.y;
    // never past what the mask workspace was sized for (max_count), even if the caller's counts are larger
    const int n = (int)min64(min64(counts[j], capacity), (int64_t)col_blocks_alloc * 64);
    // blockIdx.x enumerates the UPPER-TRIANGULAR 64x64 tiles row by row: tile t of row r starts at r*cb - r(r-1)/2
    // (a square grid launched twice as many blocks, half of which only returned: this kernel is block-scheduling bound)
    const int cb = col_blocks_alloc;
    const int cbn = (n + 63) / 64;                                    // column blocks that hold boxes
    // tiles beyond the live ones belong to rows/columns >= n: blocks stride over the tile index space, which the host caps
    const long long tiles_total = (long long)cb * (cb + 1) / 2;
    for (long long t = blockIdx.x; t < tiles_total; t += gridDim.x) {
    __syncthreads();                                                  // shared staging is reused by the next tile
    int row_blk = (int)floor(((2.0 * cb + 1.0) - sqrt((2.0 * cb + 1.0) * (2.0 * cb + 1.0) - 8.0 * (double)t)) * 0.5);
    row_blk = max(0, min(row_blk, cb - 1));
    while (row_blk > 0 && (long long)row_blk * cb - (long long)row_blk * (row_blk - 1) / 2 > t) --row_blk;
    while ((long long)(row_blk + 1) * cb - (long long)(row_blk + 1) * row_blk / 2 <= t) ++row_blk;
    const int col_blk = row_blk + (int)(t - ((long long)row_blk * cb - (long long)row_blk * (row_blk - 1) / 2));
    if (row_blk >= cbn) break;                                        // rows are enumerated in order: nothing live follows
    if (col_blk >= cbn) continue;
    const int mode = (int)(info[2 * j + 1] & 0xffu);
    // labels provably never interact (see nms_prepare_kernel): pairs of different labels are skipped on the label test alone
    const bool by_label = (mode == 2) || ((info[2 * j + 1] & kNmsLabelsDisjoint) != 0u && thr >= 0.0f);
    const float off_unit = __uint_as_float(info[2 * j]) + 1.0f;       // max_coordinate + 1
    const cldet_candidate* c = sorted + (int64_t)j * capacity;

    __shared__ float4 cbox[64];
    __shared__ float carea[64];
    __shared__ int clab[64];
    // 256 threads = 64 rows x 4 column quarters (16 columns each); the first 64 threads stage the column boxes
    const int rr = threadIdx.x & 63, qq = threadIdx.x >> 6;
    __shared__ unsigned int part[4][64];
    const int ci = col_blk * 64 + threadIdx.x;
    if (threadIdx.x < 64 && ci < n) {
        const float4* rec = reinterpret_cast<const float4*>(c + ci);
        float4 v = rec[0];
        const int lab = __float_as_int(rec[1].y);
        if (mode == 1) {                                              // boxes + (label * (max + 1))[:, None]
            const float off = (float)lab * off_unit;
            v.x += off; v.y += off; v.z += off; v.w += off;
        }
        cbox[threadIdx.x] = v;
        carea[threadIdx.x] = (v.z - v.x) * (v.w - v.y);
        clab[threadIdx.x] = lab;
    }
    __syncthreads();
    const int ri = row_blk * 64 + rr;
    const bool live = ri < n;
    const float4* rrec = reinterpret_cast<const float4*>(c + (live ? ri : 0));
    float4 me = rrec[0];
    const int my_lab = __float_as_int(rrec[1].y);
    if (mode == 1) {
        const float off = (float)my_lab * off_unit;
        me.x += off; me.y += off; me.z += off; me.w += off;
    }
    const float my_area = (me.z - me.x) * (me.w - me.y);
    const int ncol = min(64, n - col_blk * 64);
    unsigned int bits = 0;                                            // this quarter's 16 columns
    const int start = (row_blk == col_blk) ? rr + 1 : 0;
    const int t_lo = max(start, 16 * qq), t_hi = min(ncol, 16 * qq + 16);
    if (live) {
        for (int t = t_lo; t < t_hi; ++t) {
            if (by_label && clab[t] != my_lab) continue;
            if (suppresses(me, my_area, cbox[t], carea[t], thr)) bits |= 1u << (t - 16 * qq);
        }
    }
    part[qq][rr] = bits;
    __syncthreads();
    if (qq == 0 && live) {
        const uint64_t all = (uint64_t)part[0][rr] | ((uint64_t)part[1][rr] << 16) | ((uint64_t)part[2][rr] << 32) |
                             ((uint64_t)part[3][rr] << 48);
        mask[(int64_t)j * mask_stride_img + (int64_t)ri * col_blocks_alloc + col_blk] = all;
    }
    }
}

// Sequential part of greedy NMS, one block per image, 64 sorted boxes per step:
//   (1) 64 threads fetch the chunk's diagonal mask words into shared memory,
//   (2) one warp resolves the chunk serially in registers (who survives inside the chunk),
//   (3) all 256 threads OR the surviving rows' mask words into the running "removed" bitmap of the later chunks:
//       thread = (column word, row group), 16 x 16, so every thread issues its <= 4 loads at once (the first version
//       walked 64 dependent L2 loads per thread and took 13 us per chunk).
__global__ void __launch_bounds__(256)
nms_resolve_kernel(const int32_t* __restrict__ counts, int64_t capacity, const uint64_t* __restrict__ mask,
                   int64_t mask_stride_img, int col_blocks_alloc, uint64_t* __restrict__ remv_ws, int32_t* __restrict__ keep,
                   int32_t* __restrict__ keep_counts) {
    pdl_wait();                       // may be scheduled while the previous kernel of the chain drains
    pdl_launch_dependents();
    __shared__ uint64_t diag[64];
    __shared__ uint64_t kept_bits;
    __shared__ int kept_total;
    __shared__ unsigned long long acc[16];
    const int j = blockIdx.x;
    if (min64(counts[j], capacity) > (int64_t)col_blocks_alloc * 64) {     // more boxes than the mask workspace was sized for
        if (threadIdx.x == 0) keep_counts[j] = -1;                         // report, never truncate silently
        return;
    }
    const int n = (int)min64(counts[j], capacity);
    const int col_blocks = (n + 63) / 64;
    uint64_t* remv = remv_ws + (int64_t)j * col_blocks_alloc;
    const uint64_t* m = mask + (int64_t)j * mask_stride_img;
    for (int t = threadIdx.x; t < col_blocks; t += blockDim.x) remv[t] = 0;
    if (threadIdx.x == 0) kept_total = 0;
    __syncthreads();
    for (int c = 0; c < col_blocks; ++c) {
        const int rows = min(64, n - c * 64);
        if (threadIdx.x < 64)
            diag[threadIdx.x] = (threadIdx.x < rows) ? m[(int64_t)(c * 64 + threadIdx.x) * col_blocks_alloc + c] : 0ull;
        __syncthreads();
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            uint64_t cur = remv[c];
            uint64_t kept = 0;
#pragma unroll 16
            for (int b = 0; b < 64; ++b) {
                const uint64_t w = diag[b];                       // broadcast read, independent of the chain
                if (b < rows && !((cur >> b) & 1ull)) {
                    kept |= 1ull << b;
                    cur |= w;
                }
            }
            // ordered output of this chunk's survivors
            const int base = kept_total;
            const uint64_t lo0 = (1ull << lane) - 1ull;
            if ((kept >> lane) & 1ull) keep[(int64_t)j * capacity + base + __popcll(kept & lo0)] = c * 64 + lane;
            const uint64_t lo1 = (1ull << (lane + 32)) - 1ull;
            if ((kept >> (lane + 32)) & 1ull) keep[(int64_t)j * capacity + base + __popcll(kept & lo1)] = c * 64 + lane + 32;
            __syncwarp();
            if (lane == 0) {
                kept_bits = kept;
                kept_total = base + __popcll(kept);
            }
        }
        __syncthreads();
        const uint64_t kept = kept_bits;
        // absorb: later column words, 16 at a time; thread (tc, tg) handles word t0+tc and rows tg, tg+16, tg+32, tg+48
        const int tc = threadIdx.x & 15, tg = threadIdx.x >> 4;
        for (int t0 = c + 1; t0 < col_blocks; t0 += 16) {
            if (threadIdx.x < 16) acc[threadIdx.x] = 0ull;
            __syncthreads();
            const int t = t0 + tc;
            if (t < col_blocks) {
                uint64_t v = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int b = tg + 16 * q;
                    if ((kept >> b) & 1ull) v |= m[(int64_t)(c * 64 + b) * col_blocks_alloc + t];
                }
                if (v) atomicOr(&acc[tc], (unsigned long long)v);
            }
            __syncthreads();
            if (threadIdx.x < 16 && t0 + (int)threadIdx.x < col_blocks) remv[t0 + threadIdx.x] |= acc[threadIdx.x];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) keep_counts[j] = kept_total;
}

// Tail of the resolve kernels: the block that resolved image j also gathers its kept candidates into the dense outputs
// (scores / labels / boxes rows of `capacity` entries) -- one launch fewer on a chain that is bound by launch latency.
__device__ __forceinline__ void gather_image(const cldet_candidate* __restrict__ sorted, const int32_t* __restrict__ keep, int j,
                                             int64_t capacity, int kept, float* __restrict__ scores, int64_t* __restrict__ labels,
                                             float4* __restrict__ boxes) {
    for (int i = threadIdx.x; i < kept; i += blockDim.x) {
        const float4* rec = reinterpret_cast<const float4*>(sorted + (int64_t)j * capacity + keep[(int64_t)j * capacity + i]);
        const float4 b = rec[0], t = rec[1];
        scores[(int64_t)j * capacity + i] = t.x;
        labels[(int64_t)j * capacity + i] = (int64_t)__float_as_int(t.y);
        boxes[(int64_t)j * capacity + i] = b;
    }
}

// Small-K resolve (K <= kSmemResolveMax, e.g. after the top-1000 stage): the whole block pulls the image's upper-triangular
// mask (<= 185 KB) into shared memory with independent coalesced loads, then ONE warp runs the sequential part with the
// "removed" bitmap in registers (lane t owns column word t).  The dependent chain sees no global load and no block barrier:
// ~1 us per 64 boxes instead of ~4 us in nms_resolve_kernel.
constexpr int kSmemResolveMax = 1216;      // 1216 rows x 19 words x 8 B = 184.8 KB of the 227 KB a CTA may use

constexpr int kSmemResolveThreads = 1024;

__global__ void __launch_bounds__(kSmemResolveThreads)
nms_resolve_smem_kernel(const int32_t* __restrict__ counts, int64_t capacity, const uint64_t* __restrict__ mask,
                        int64_t mask_stride_img, int col_blocks_alloc, int32_t* __restrict__ keep,
                        int32_t* __restrict__ keep_counts, const cldet_candidate* __restrict__ sorted = nullptr,
                        float* __restrict__ out_scores = nullptr, int64_t* __restrict__ out_labels = nullptr,
                        float4* __restrict__ out_boxes = nullptr) {
    pdl_wait();                       // may be scheduled while the previous kernel of the chain drains
    pdl_launch_dependents();
    extern __shared__ uint64_t sm_mask[];                                // [64*cb][cb], rows >= n zero
    __shared__ int kept_total_s;
    const int j = blockIdx.x;
    if (min64(counts[j], capacity) > (int64_t)col_blocks_alloc * 64) {     // more boxes than the mask workspace was sized for
        if (threadIdx.x == 0) keep_counts[j] = -1;                         // report, never truncate silently
        return;
    }
    const int n = (int)min64(counts[j], capacity);
    const int cb = (n + 63) / 64;
    const uint64_t* m = mask + (int64_t)j * mask_stride_img;
    const int total = 64 * cb * cb;
    // four independent loads in flight per thread (the copy is latency-bound: one CTA per image)
    for (int i0 = threadIdx.x; i0 < total; i0 += 4 * kSmemResolveThreads) {
        uint64_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = i0 + u * kSmemResolveThreads;
            const int row = idx / cb, t = idx - row * cb;
            // the lower triangle is never written by nms_mask_kernel and never read below; rows past n read as zero
            v[u] = (idx < total && row < n && t >= (row >> 6)) ? m[(int64_t)row * col_blocks_alloc + t] : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = i0 + u * kSmemResolveThreads;
            if (idx < total) sm_mask[idx] = v[u];
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    uint64_t remv = 0;                                                    // lane t: removed bits of boxes 64t .. 64t+63
    int kept_total = 0;
    for (int c = 0; c < cb; ++c) {
        const int rows = min(64, n - c * 64);
        uint64_t cur = __shfl_sync(0xffffffffu, remv, c);
        if (rows < 64) cur |= ~0ull << rows;                              // slots past the end can never be kept
        uint64_t kept = 0;
        const uint64_t* diag = sm_mask + (int64_t)(c * 64) * cb + c;
        // 16 rows at a time: the (broadcast) loads first, then a branch-free chain of constant-position bit tests
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint64_t w[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) w[q] = diag[(g * 16 + q) * cb];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int b = g * 16 + q;
                const bool take = !((cur >> b) & 1ull);
                kept |= take ? (1ull << b) : 0ull;
                cur |= take ? w[q] : 0ull;
            }
        }
        const uint64_t lo0 = (1ull << lane) - 1ull;
        if ((kept >> lane) & 1ull) keep[(int64_t)j * capacity + kept_total + __popcll(kept & lo0)] = c * 64 + lane;
        const uint64_t lo1 = (1ull << (lane + 32)) - 1ull;
        if ((kept >> (lane + 32)) & 1ull) keep[(int64_t)j * capacity + kept_total + __popcll(kept & lo1)] = c * 64 + lane + 32;
        kept_total += __popcll(kept);
        if (lane > c && lane < cb) {                                      // absorb the survivors' rows into the later words
            const uint64_t* rowp = sm_mask + (int64_t)(c * 64) * cb + lane;
            uint64_t v0 = 0, v1 = 0;
#pragma unroll
            for (int b = 0; b < 64; b += 2) {                             // unconditional, independent loads, masked by the bit
                v0 |= rowp[b * cb] & (0ull - ((kept >> b) & 1ull));
                v1 |= rowp[(b + 1) * cb] & (0ull - ((kept >> (b + 1)) & 1ull));
            }
            remv |= v0 | v1;
        }
    }
    if (lane == 0) {
        keep_counts[j] = kept_total;
        kept_total_s = kept_total;
    }
    }
    if (!out_scores) return;
    __syncthreads();
    gather_image(sorted, keep, j, capacity, kept_total_s, out_scores, out_labels, out_boxes);
}

// Any-K resolve: one CTA per image, the "removed" bitmap in shared memory, and a two-stage software pipeline over the 64-box
// chunks so that no global-memory latency sits on the sequential chain:
//   warp 0 (the chain)  chunk c: cur = removed[c]; 64-step greedy scan over the chunk's DIAGONAL mask block (prefetched into
//                       shared memory one chunk ahead) -> kept bits; write the chunk's survivors; fold the kept rows' words of the
//                       NEXT column block (c, c+1) (also prefetched) into removed[c+1] -- all the next chunk needs from this one;
//   warps 1..31         meanwhile absorb chunk c-1's kept rows into removed[c+1 ..] with independent coalesced loads (every
//                       thread owns column words), and prefetch the diagonal and next blocks of chunk c+1.
// One block barrier per chunk.  removed[c] is complete when the chain reads it: chunk c-1 reached it through the folded next
// block, chunks <= c-2 through absorb passes that ended at earlier barriers.  Replaces both the 64-dependent-loads-per-chunk
// nms_resolve_kernel (K > 1216) and, when it wins, the whole-mask-in-shared-memory variant.
constexpr int kStreamThreads = 1024;
constexpr int kStreamAbsorbWarps = 24;     // warps 1..24 absorb, 25..28 prefetch, warp 0 runs the chain
constexpr int kStreamUnroll = 16;           // independent loads in flight per absorber thread (192 column words per round)

__global__ void __launch_bounds__(kStreamThreads)
nms_resolve_stream_kernel(const int32_t* __restrict__ counts, int64_t capacity, const uint64_t* __restrict__ mask,
                          int64_t mask_stride_img, int col_blocks_alloc, int32_t* __restrict__ keep,
                          int32_t* __restrict__ keep_counts, const cldet_candidate* __restrict__ sorted = nullptr,
                          float* __restrict__ out_scores = nullptr, int64_t* __restrict__ out_labels = nullptr,
                          float4* __restrict__ out_boxes = nullptr) {
    pdl_wait();                       // may be scheduled while the previous kernel of the chain drains
    pdl_launch_dependents();
    extern __shared__ unsigned long long removed[];                      // [col_blocks]
    __shared__ int kept_total_s;
    __shared__ uint64_t diag[2][64];
    __shared__ uint64_t nextb[2][64];
    __shared__ uint64_t kept_s[2];
    const int j = blockIdx.x;
    const int64_t cnt_j = min64(counts[j], capacity);
    if (cnt_j > (int64_t)col_blocks_alloc * 64) {                         // more boxes than the mask workspace was sized for:
        if (threadIdx.x == 0) keep_counts[j] = -1;                        // report, never truncate silently
        return;
    }
    const int n = (int)cnt_j;
    const int cb = (n + 63) / 64;
    const uint64_t* m = mask + (int64_t)j * mask_stride_img;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int t = tid; t < cb; t += kStreamThreads) removed[t] = 0ull;
    if (tid < 2) kept_s[tid] = 0ull;
    // prefetch for chunk 0
    if (tid < 128) {
        const int b = tid & 63, which = tid >> 6;                         // 0: block (0,0), 1: block (0,1)
        const int row = b, word = which;
        uint64_t v = 0ull;
        if (row < n && word < cb) v = m[(int64_t)row * col_blocks_alloc + word];
        if (which == 0) diag[0][b] = v; else nextb[0][b] = v;
    }
    __syncthreads();
    int kept_total = 0;
    for (int c = 0; c < cb; ++c) {
        const int buf = c & 1;
        if (warp == 0) {
            const int rows = min(64, n - c * 64);
            uint64_t cur = removed[c];
            if (rows < 64) cur |= ~0ull << rows;                           // slots past the end can never be kept
            // Greedy scan of the chunk, decided in parallel rounds instead of 64 dependent steps (see nms_fused_kernel): lane l
            // holds the diagonal words of rows l and l+32 (bits > own row only); an undecided box that NO undecided box could
            // suppress is kept (the first undecided one always is), the boxes the newly kept ones suppress leave the undecided
            // set; two warp-wide ORs per round, box by box after kChainRounds rounds.  Identical to the sequential scan.
            const uint64_t d0 = diag[buf][lane], d1 = diag[buf][lane + 32];
            uint64_t und = ~cur, kept = 0ull;
            int round = 0;
            while (und) {
                if (++round > kChainRounds) {
                    while (und) {
                        const int b = __ffsll((long long)und) - 1;
                        const uint64_t dsel = (b < 32) ? d0 : d1;
                        const uint64_t d = __shfl_sync(0xffffffffu, dsel, b & 31);
                        kept |= 1ull << b;
                        und &= ~((1ull << b) | d);
                    }
                    break;
                }
                const uint64_t mine = (((und >> lane) & 1ull) ? d0 : 0ull) | (((und >> (lane + 32)) & 1ull) ? d1 : 0ull);
                const uint64_t threat = ((uint64_t)__reduce_or_sync(0xffffffffu, (unsigned)(mine >> 32)) << 32) |
                                        __reduce_or_sync(0xffffffffu, (unsigned)mine);
                if (!(threat & und)) {                                    // nobody undecided is threatened: keep them all
                    kept |= und;
                    break;
                }
                const uint64_t fresh = und & ~threat;
                const uint64_t hit = (((fresh >> lane) & 1ull) ? d0 : 0ull) | (((fresh >> (lane + 32)) & 1ull) ? d1 : 0ull);
                const uint64_t gone = ((uint64_t)__reduce_or_sync(0xffffffffu, (unsigned)(hit >> 32)) << 32) |
                                      __reduce_or_sync(0xffffffffu, (unsigned)hit);
                kept |= fresh;
                und &= ~(fresh | gone);
            }
            if (c + 1 < cb) {                                             // what chunk c+1 needs from this chunk
                const uint64_t v = (((kept >> lane) & 1ull) ? nextb[buf][lane] : 0ull) |
                                   (((kept >> (lane + 32)) & 1ull) ? nextb[buf][lane + 32] : 0ull);
                const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v);
                const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
                if (lane == 0) atomicOr(&removed[c + 1], ((unsigned long long)hi << 32) | lo);
            }
            if (lane == 0) kept_s[buf] = kept;
            const uint64_t lo0 = (1ull << lane) - 1ull;
            if ((kept >> lane) & 1ull) keep[(int64_t)j * capacity + kept_total + __popcll(kept & lo0)] = c * 64 + lane;
            const uint64_t lo1 = (1ull << (lane + 32)) - 1ull;
            if ((kept >> (lane + 32)) & 1ull) keep[(int64_t)j * capacity + kept_total + __popcll(kept & lo1)] = c * 64 + lane + 32;
            kept_total += __popcll(kept);
        } else if (warp >= 1 + kStreamAbsorbWarps) {
            // four warps prefetch the diagonal and next blocks of chunk c+1 (three warps idle)
            const int q = tid - 32 * (1 + kStreamAbsorbWarps);
            if (q < 128) {
                const int b = q & 63, which = q >> 6;
                const int row = (c + 1) * 64 + b, word = c + 1 + which;
                uint64_t v = 0ull;
                if (row < n && word < cb) v = m[(int64_t)row * col_blocks_alloc + word];
                if (which == 0) diag[buf ^ 1][b] = v; else nextb[buf ^ 1][b] = v;
            }
        } else if (c >= 1 && c + 1 < cb) {
            // Absorb chunk c-1's kept rows into the column words c+1 ..  A warp = 8 rows x 4 consecutive words (one 32-byte
            // sector per row); the 24 warps = 8 row groups x 3 sector slots.  Every thread issues kStreamUnroll INDEPENDENT
            // loads per round (one L2 latency per round, not per row) and publishes the nonzero words -- few: a box
            // suppresses a handful of others -- with shared-memory atomics.
            const int aw = warp - 1;
            const int b = (aw & 7) * 8 + (lane >> 2), wq = lane & 3, ss = aw >> 3;
            const uint64_t kp = kept_s[buf ^ 1];
            const bool on = (kp >> b) & 1ull;
            const uint64_t* rowp = m + ((int64_t)(c - 1) * 64 + b) * col_blocks_alloc;
            constexpr int kSlots = kStreamAbsorbWarps / 8;                // sector slots per round
            if (on) {
                for (int w0 = c + 1 + 4 * ss + wq; w0 < cb; w0 += 4 * kSlots * kStreamUnroll) {
                    uint64_t v[kStreamUnroll];
#pragma unroll
                    for (int u = 0; u < kStreamUnroll; ++u) {
                        const int wd = w0 + 4 * kSlots * u;
                        v[u] = (wd < cb) ? rowp[wd] : 0ull;
                    }
#pragma unroll
                    for (int u = 0; u < kStreamUnroll; ++u)
                        if (v[u]) atomicOr(&removed[w0 + 4 * kSlots * u], (unsigned long long)v[u]);     // rare: suppression is sparse
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        keep_counts[j] = kept_total;
        kept_total_s = kept_total;
    }
    if (!out_scores) return;
    __syncthreads();
    gather_image(sorted, keep, j, capacity, kept_total_s, out_scores, out_labels, out_boxes);
}

// ------------------------------------------------------------------------------------------------
// K6 in ONE launch for short lists (<= kFuseMax boxes per image: the pre-NMS top-k regime): a thread-block CLUSTER of
// kFuseCluster CTAs per image does what nms_prepare + nms_mask + the resolve + the gather do as three dependent launches.
//   every CTA   stages the image's sorted boxes in its own shared memory, forms the per-image facts of nms_prepare_kernel
//               (max coordinate, mode, "labels provably disjoint") redundantly -- max/min are order-independent, so the four
//               CTAs agree bit for bit -- and applies the coordinate-trick offsets;
//   mask        the upper-triangular 64x64 tiles are dealt round-robin to the cluster's 64 two-warp slots; a thread owns one
//               row of its tile.  When only equal labels can interact it first builds the 64-bit set of equal-label columns
//               (16 broadcast 128-bit loads of labels, no divergence) and then visits only those columns -- the tile kernel
//               above pays a divergent branch per column and runs the IoU test whenever ANY lane's label matches.  The row's
//               64-bit word goes straight into CTA 0's shared memory (distributed shared memory store);
//   resolve     after one cluster barrier CTA 0 holds the whole mask (<= 144 KB) on chip: warp 0 runs the greedy chain with
//               no global-memory latency anywhere -- a 64-box chunk is decided in parallel rounds of warp-wide ORs (see below)
//               instead of 64 dependent steps --, folds the kept rows' words of the next column block itself and leaves the
//               later column words to warps 1..4, one chunk behind; the five warps meet at a named barrier per chunk, the
//               other warps wait for the gather.
// Same arithmetic as the three-kernel path (suppresses(), the offsets, the areas), so keep lists are bit-identical to it.
// ------------------------------------------------------------------------------------------------
constexpr int kFuseMax = 1024;
constexpr int kFuseCb = kFuseMax / 64;         // column words per mask row
constexpr int kFuseCluster = 4;                // 32 images x 4 CTAs: one wave on 148 SMs (one CTA per SM: 172 KB of shared memory)
constexpr int kFuseThreads = 1024;
constexpr int kFuseSlots = kFuseThreads / 64;  // two-warp tile slots per CTA
constexpr int kFuseChainWarps = 5;             // warp 0 runs the chain, warps 1..4 absorb 16 rows each, one chunk behind
constexpr size_t kFuseSmemBytes =
    (size_t)kFuseMax * ((kFuseCb + 2) * sizeof(uint64_t) + sizeof(float4) + sizeof(float) + 2 * sizeof(int));

__global__ void __cluster_dims__(kFuseCluster, 1, 1) __launch_bounds__(kFuseThreads)
nms_fused_kernel(const cldet_candidate* __restrict__ sorted, const int32_t* __restrict__ counts, int64_t capacity, int mode,
                 int64_t vanilla_numel_limit, float thr, int max_rows, int32_t* __restrict__ keep, int32_t* __restrict__ keep_counts,
                 float* __restrict__ out_scores, int64_t* __restrict__ out_labels, float4* __restrict__ out_boxes) {
    pdl_wait();                       // may be scheduled while the previous kernel of the chain drains
    pdl_launch_dependents();
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char fuse_smem[];
    uint64_t* s_mask = reinterpret_cast<uint64_t*>(fuse_smem);                       // [kFuseMax][kFuseCb], used in CTA 0 only
    uint64_t* s_diag = s_mask + (size_t)kFuseMax * kFuseCb;                          // [kFuseMax] word (row block, row block) of every row
    uint64_t* s_next = s_diag + kFuseMax;                                            // [kFuseMax] word (row block, row block + 1)
    float4* s_box = reinterpret_cast<float4*>(s_next + kFuseMax);
    float* s_area = reinterpret_cast<float*>(s_box + kFuseMax);
    int* s_lab = reinterpret_cast<int*>(s_area + kFuseMax);
    int* s_keep = s_lab + kFuseMax;
    __shared__ float red_max[32], red_min[32];
    __shared__ int red_lab[32];
    __shared__ unsigned long long s_removed[kFuseCb];
    __shared__ unsigned long long s_kept[2];
    __shared__ int s_kept_total;
    const int rank = (int)cluster.block_rank();
    const int j = blockIdx.x / kFuseCluster;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t cnt_j = min64(counts[j], capacity);
    if (cnt_j > max_rows) {                                   // more boxes than the caller's max_count (rounded up to 64, <= kFuseMax):
        if (rank == 0 && tid == 0) keep_counts[j] = -1;       // report, never truncate silently (uniform over the cluster)
        return;
    }
    const int n = (int)cnt_j;
    if (n <= 0) {
        if (rank == 0 && tid == 0) keep_counts[j] = 0;
        return;
    }
    // ---- stage the boxes; the facts nms_prepare_kernel publishes ----
    float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
    int my_lab = 0;
    float m = -INFINITY, mn = INFINITY;
    int lab_max = 0, lab_min = 0;
    if (tid < n) {
        const float4* rec = reinterpret_cast<const float4*>(sorted + (int64_t)j * capacity + tid);
        me = rec[0];
        my_lab = __float_as_int(rec[1].y);
        m = fmaxf(fmaxf(me.x, me.y), fmaxf(me.z, me.w));
        mn = fminf(fminf(me.x, me.y), fminf(me.z, me.w));
        lab_max = max(0, my_lab);
        lab_min = min(0, my_lab);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        lab_max = max(lab_max, __shfl_xor_sync(0xffffffffu, lab_max, o));
        lab_min = min(lab_min, __shfl_xor_sync(0xffffffffu, lab_min, o));
    }
    if (lane == 0) {
        red_max[warp] = m;
        red_min[warp] = mn;
        red_lab[warp] = (lab_min < 0) ? 0x7fffffff : lab_max;      // a negative label disables the label shortcut
    }
    if (tid < kFuseCb) s_removed[tid] = 0ull;
    if (tid < 2) s_kept[tid] = 0ull;
    __syncthreads();
    m = red_max[lane];
    mn = red_min[lane];
    int lmax = red_lab[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    }
    int used = mode;
    if (mode == 0) used = ((int64_t)n * 4 > vanilla_numel_limit) ? 2 : 1;       // torchvision ops/boxes.py batched_nms
    const bool disjoint = (used == 1) && mn >= 0.0f && lmax < 0x7fffffff && ((double)lmax + 1.0) * ((double)m + 1.0) <= 2097152.0;
    const bool by_label = (used == 2) || (disjoint && thr >= 0.0f);              // see nms_prepare_kernel / nms_mask_kernel
    if (tid < n) {
        if (used == 1) {                                                         // boxes + (label * (max + 1))[:, None]
            const float off = (float)my_lab * (m + 1.0f);
            me.x += off; me.y += off; me.z += off; me.w += off;
        }
        s_box[tid] = me;
        s_area[tid] = (me.z - me.x) * (me.w - me.y);
    }
    s_lab[tid] = (tid < n) ? my_lab : 0;
    __syncthreads();
    // ---- suppression mask, written into CTA 0's shared memory ----
    const int cbn = (n + 63) >> 6;
    const int tiles = cbn * (cbn + 1) / 2;
    uint64_t* mask0 = cluster.map_shared_rank(s_mask, 0);
    uint64_t* diag0 = cluster.map_shared_rank(s_diag, 0);
    uint64_t* next0 = cluster.map_shared_rank(s_next, 0);
    const int rr = tid & 63;
    for (int t = rank * kFuseSlots + (tid >> 6); t < tiles; t += kFuseCluster * kFuseSlots) {
        int row_blk = 0, rem = t, len = cbn;                  // upper-triangular tiles, row by row
        while (rem >= len) {
            rem -= len;
            --len;
            ++row_blk;
        }
        const int col_blk = row_blk + rem;
        const int ri = row_blk * 64 + rr;
        if (ri >= n) continue;
        const int c0 = col_blk * 64;
        const int ncol = min(64, n - c0);
        uint64_t todo = (ncol == 64) ? ~0ull : ((1ull << ncol) - 1ull);
        if (row_blk == col_blk) todo &= (rr == 63) ? 0ull : (~0ull << (rr + 1));      // only later boxes of the own block
        const int lab = s_lab[ri];
        if (by_label) {
            const int4* l4 = reinterpret_cast<const int4*>(s_lab + c0);
            uint32_t elo = 0, ehi = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int4 a = l4[q], b = l4[q + 8];
                elo |= ((a.x == lab) ? 1u : 0u) << (4 * q) | ((a.y == lab) ? 2u : 0u) << (4 * q) | ((a.z == lab) ? 4u : 0u) << (4 * q) |
                       ((a.w == lab) ? 8u : 0u) << (4 * q);
                ehi |= ((b.x == lab) ? 1u : 0u) << (4 * q) | ((b.y == lab) ? 2u : 0u) << (4 * q) | ((b.z == lab) ? 4u : 0u) << (4 * q) |
                       ((b.w == lab) ? 8u : 0u) << (4 * q);
            }
            todo &= ((uint64_t)ehi << 32) | elo;
        }
        const float4 mine = s_box[ri];
        const float my_area = s_area[ri];
        uint64_t bits = 0ull;
        while (todo) {
            const int tc = __ffsll((long long)todo) - 1;
            todo &= todo - 1ull;
            if (suppresses(mine, my_area, s_box[c0 + tc], s_area[c0 + tc], thr)) bits |= 1ull << tc;
        }
        // the two words the chain itself reads go to compact arrays (conflict-free for a warp), the rest into the row-major mask
        if (col_blk == row_blk) diag0[ri] = bits;
        else if (col_blk == row_blk + 1) next0[ri] = bits;
        else mask0[(size_t)ri * kFuseCb + col_blk] = bits;
    }
    cluster.sync();                                           // every word of the mask has landed in CTA 0
    if (rank != 0) return;
    // ---- greedy resolve from shared memory: warp 0 = the chain, warps 1..4 = absorb one chunk behind ----
    if (warp < kFuseChainWarps) {
        int kept_total = 0;
        for (int c = 0; c < cbn; ++c) {
            const int buf = c & 1;
            if (warp == 0) {
                const int rows = min(64, n - c * 64);
                uint64_t cur = s_removed[c];
                if (rows < 64) cur |= ~0ull << rows;                          // slots past the end can never be kept
                // Lane l holds the diagonal words of rows l and l+32 (bits > own row only).  Decide the chunk in parallel rounds:
                // an undecided box that NO undecided box could suppress is kept (the first undecided one always is); the boxes
                // the newly kept ones suppress leave the undecided set.  Identical to the sequential greedy scan; sparse
                // suppression takes 1-3 rounds (4 warp-wide ORs each) instead of 64 dependent steps.
                const uint64_t d0 = s_diag[c * 64 + lane], d1 = s_diag[c * 64 + 32 + lane];
                uint64_t und = ~cur, kept = 0ull;
                int round = 0;
                while (und) {
                    if (++round > kChainRounds) {                              // a long dependency chain: finish it box by box
                        while (und) {
                            const int b = __ffsll((long long)und) - 1;
                            const uint64_t dsel = (b < 32) ? d0 : d1;
                            const uint64_t d = __shfl_sync(0xffffffffu, dsel, b & 31);
                            kept |= 1ull << b;
                            und &= ~((1ull << b) | d);
                        }
                        break;
                    }
                    const uint64_t mine = (((und >> lane) & 1ull) ? d0 : 0ull) | (((und >> (lane + 32)) & 1ull) ? d1 : 0ull);
                    const uint64_t threat = ((uint64_t)__reduce_or_sync(0xffffffffu, (unsigned)(mine >> 32)) << 32) |
                                            __reduce_or_sync(0xffffffffu, (unsigned)mine);
                    if (!(threat & und)) {                                    // nobody undecided is threatened: keep them all (the
                        kept |= und;                                          // common chunk: one reduction stage instead of two)
                        break;
                    }
                    const uint64_t fresh = und & ~threat;                      // never empty: the first undecided box is in it
                    const uint64_t hit = (((fresh >> lane) & 1ull) ? d0 : 0ull) | (((fresh >> (lane + 32)) & 1ull) ? d1 : 0ull);
                    const uint64_t gone = ((uint64_t)__reduce_or_sync(0xffffffffu, (unsigned)(hit >> 32)) << 32) |
                                          __reduce_or_sync(0xffffffffu, (unsigned)hit);
                    kept |= fresh;
                    und &= ~(fresh | gone);
                }
                if (c + 1 < cbn) {                                            // what chunk c+1 needs from this chunk
                    const uint64_t v = (((kept >> lane) & 1ull) ? s_next[c * 64 + lane] : 0ull) |
                                       (((kept >> (lane + 32)) & 1ull) ? s_next[c * 64 + 32 + lane] : 0ull);
                    const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v);
                    const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
                    if (lane == 0) atomicOr(&s_removed[c + 1], ((unsigned long long)hi << 32) | lo);
                }
                if (lane == 0) s_kept[buf] = kept;
                const uint64_t lo0 = (1ull << lane) - 1ull;
                if ((kept >> lane) & 1ull) {
                    const int pos = kept_total + __popcll(kept & lo0);
                    keep[(int64_t)j * capacity + pos] = c * 64 + lane;
                    s_keep[pos] = c * 64 + lane;
                }
                const uint64_t lo1 = (1ull << (lane + 32)) - 1ull;
                if ((kept >> (lane + 32)) & 1ull) {
                    const int pos = kept_total + __popcll(kept & lo1);
                    keep[(int64_t)j * capacity + pos] = c * 64 + lane + 32;
                    s_keep[pos] = c * 64 + lane + 32;
                }
                kept_total += __popcll(kept);
            } else if (c >= 1 && c + 1 < cbn) {
                // chunk c-1's kept rows into the column words c+1 ..: warp w takes 16 of the 64 rows, lane = (word, 8 of those rows)
                const int wd = c + 1 + (lane & 15), r0 = (warp - 1) * 16 + (lane >> 4) * 8;
                if (wd < cbn) {
                    const uint64_t kp = s_kept[buf ^ 1] >> r0;
                    const uint64_t* rowp = s_mask + (size_t)((c - 1) * 64 + r0) * kFuseCb + wd;
                    uint64_t v0 = 0ull, v1 = 0ull;
#pragma unroll
                    for (int b = 0; b < 8; b += 2) {                          // unconditional independent loads, masked by the bit
                        v0 |= rowp[b * kFuseCb] & (0ull - ((kp >> b) & 1ull));
                        v1 |= rowp[(b + 1) * kFuseCb] & (0ull - ((kp >> (b + 1)) & 1ull));
                    }
                    const uint64_t v = v0 | v1;
                    if (v) atomicOr(&s_removed[wd], (unsigned long long)v);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * kFuseChainWarps) : "memory");      // the chain warp and the absorbers only
        }
        if (tid == 0) {
            keep_counts[j] = kept_total;
            s_kept_total = kept_total;
        }
    }
    if (!out_scores) return;
    __syncthreads();
    const int kept_n = s_kept_total;
    for (int i = tid; i < kept_n; i += kFuseThreads) {
        const float4* rec = reinterpret_cast<const float4*>(sorted + (int64_t)j * capacity + s_keep[i]);
        const float4 b = rec[0], t = rec[1];
        out_scores[(int64_t)j * capacity + i] = t.x;
        out_labels[(int64_t)j * capacity + i] = (int64_t)__float_as_int(t.y);
        out_boxes[(int64_t)j * capacity + i] = b;
    }
}

__global__ void __launch_bounds__(256)
gather_detections_kernel(const cldet_candidate* __restrict__ sorted, const int32_t* __restrict__ keep,
                         const int32_t* __restrict__ keep_counts, int64_t capacity, float* __restrict__ scores,
                         int64_t* __restrict__ labels, float4* __restrict__ boxes) {
    const int j = blockIdx.y;
    const int n = (int)min64(keep_counts[j], capacity);           // a negative count (error marker) gathers nothing
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const cldet_candidate c = sorted[(int64_t)j * capacity + keep[(int64_t)j * capacity + i]];
    scores[(int64_t)j * capacity + i] = c.score;
    labels[(int64_t)j * capacity + i] = (int64_t)c.label;
    boxes[(int64_t)j * capacity + i] = make_float4(c.x1, c.y1, c.x2, c.y2);
}

// batched_nms on caller-provided arrays: pack -> (rank sort) -> NMS -> original indices
__global__ void __launch_bounds__(256)
pack_boxes_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, const int64_t* __restrict__ idxs, int64_t K,
                  cldet_candidate* __restrict__ cand, uint64_t* __restrict__ keys, int32_t* __restrict__ count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count = (int32_t)K;
    if (i >= K) return;
    const float4 b = boxes[i];
    cldet_candidate c;
    c.x1 = b.x; c.y1 = b.y; c.x2 = b.z; c.y2 = b.w;
    c.score = scores[i];
    c.label = idxs ? (int32_t)idxs[i] : 0;
    c.anchor = (int32_t)i;
    c.pad = 0;
    cand[i] = c;
    keys[i] = make_key(c.score, (int)i);
}

__global__ void __launch_bounds__(256)
keep_to_original_kernel(const cldet_candidate* __restrict__ sorted, const int32_t* __restrict__ keep,
                        const int32_t* __restrict__ keep_count, int64_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *keep_count) out[i] = (int64_t)sorted[keep[i]].anchor;
}

// ------------------------------------------------------------------------------------------------
// evaluator post-processing (evaluator.py:329-361): boxes /= scale (true fp32 division, CPU ATen
// semantics), xyxy -> xywh, keep score >= threshold; records ordered by image, then by rank.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
coco_offsets_kernel(const float* __restrict__ scores, const int32_t* __restrict__ counts, int N, int64_t capacity, float thr,
                    int32_t* __restrict__ offsets /* [N+1] */) {
    // one block: per-image number of detections with score >= thr, then an exclusive scan in image order
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    for (int j0 = 0; j0 < N; j0 += blockDim.x) {
        const int j = j0 + threadIdx.x;
        int pass = 0;
        if (j < N) {
            const int n = (int)min64(counts[j], capacity);
            const float* s = scores + (int64_t)j * capacity;
            for (int i = 0; i < n; ++i) pass += (s[i] >= thr) ? 1 : 0;     // `if score < threshold: continue`
        }
        // block-wide exclusive scan of `pass` (Hillis-Steele over a shared array)
        __shared__ int sh[256];
        sh[threadIdx.x] = pass;
        __syncthreads();
        for (int off = 1; off < (int)blockDim.x; off <<= 1) {
            const int add = (threadIdx.x >= (unsigned)off) ? sh[threadIdx.x - off] : 0;
            __syncthreads();
            sh[threadIdx.x] += add;
            __syncthreads();
        }
        if (j < N) offsets[j] = total + sh[threadIdx.x] - pass;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) total += sh[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[N] = total;
}

__global__ void __launch_bounds__(256)
coco_records_kernel(const float* __restrict__ scores, const int64_t* __restrict__ labels, const float4* __restrict__ boxes,
                    const int32_t* __restrict__ counts, const float* __restrict__ scales, int64_t capacity, float thr,
                    const int32_t* __restrict__ offsets, cldet_coco_record* __restrict__ out) {
    const int j = blockIdx.y;
    const int n = (int)min64(counts[j], capacity);
    const float* s = scores + (int64_t)j * capacity;
    // scores are in NMS (descending) order, but ties and the >= test are handled generally: rank among the passing ones
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float sc = s[i];
        if (!(sc >= thr)) continue;
        int rank = 0;
        for (int t = 0; t < i; ++t) rank += (s[t] >= thr) ? 1 : 0;
        const float4 b = boxes[(int64_t)j * capacity + i];
        const float scale = scales[j];
        cldet_coco_record r;
        const float x1 = __fdiv_rn(b.x, scale), y1 = __fdiv_rn(b.y, scale), x2 = __fdiv_rn(b.z, scale), y2 = __fdiv_rn(b.w, scale);
        r.image = j;
        r.label = (int32_t)labels[(int64_t)j * capacity + i];
        r.score = sc;
        r.x = x1;
        r.y = y1;
        r.w = __fsub_rn(x2, x1);
        r.h = __fsub_rn(y2, y1);
        r.pad = 0;
        out[offsets[j] + rank] = r;
    }
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// A/B switches for experiments only (read once): CLDET_NMS_RESOLVE = stream | smem | legacy, CLDET_FORCE_RANK_SORT = 1
static int resolve_choice() {
    static const int v = [] {
        const char* e = getenv("CLDET_NMS_RESOLVE");
        if (!e) return 0;
        if (e[0] == 's' && e[1] == 't') return 1;
        if (e[0] == 's' && e[1] == 'm') return 2;
        if (e[0] == 'l') return 3;
        return 0;
    }();
    return v;
}
static bool select_multi_launch() {          // CLDET_SELECT_MULTI=1: the round-1 chain of five select launches (A/B only)
    static const bool v = [] {
        const char* e = getenv("CLDET_SELECT_MULTI");
        return e && e[0] == '1';
    }();
    return v;
}
static bool nms_fused_enabled() {             // CLDET_NMS_FUSED=0: the three-launch K6 for short lists too (A/B only)
    static const bool v = [] {
        const char* e = getenv("CLDET_NMS_FUSED");
        return !(e && e[0] == '0');
    }();
    return v;
}
static int64_t bucket_n() {                   // CLDET_BUCKET_RANK=0: the pairwise rank for short lists too (A/B only)
    static const int64_t v = [] {
        const char* e = getenv("CLDET_BUCKET_RANK");
        const char* f = getenv("CLDET_FORCE_RANK_SORT");
        return ((e && e[0] == '0') || (f && f[0] == '1')) ? (int64_t)0 : (int64_t)kBucketMax;
    }();
    return v;
}
static int k4_block_append() {                // CLDET_K4_BLOCK_APPEND=0: one append atomic per warp instead of per block (A/B only)
    static const int v = [] {
        const char* e = getenv("CLDET_K4_BLOCK_APPEND");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    return v;
}
static bool force_rank_sort() {
    static const bool v = [] {
        const char* e = getenv("CLDET_FORCE_RANK_SORT");
        return e && e[0] == '1';
    }();
    return v;
}

struct NmsWs {
    uint32_t* info;
    uint64_t* remv;
    uint64_t* mask;
    int col_blocks;
    int64_t mask_stride_img;
    size_t total;
};

static NmsWs nms_ws_layout(void* base, int N, int64_t max_count) {
    NmsWs w;
    w.col_blocks = (int)((max_count + 63) / 64);
    if (w.col_blocks < 1) w.col_blocks = 1;
    const int64_t rows = (int64_t)w.col_blocks * 64;
    w.mask_stride_img = rows * w.col_blocks;
    char* p = reinterpret_cast<char*>(base);
    size_t off = 0;
    w.info = reinterpret_cast<uint32_t*>(p + off);
    off = align_up(off + (size_t)N * 2 * sizeof(uint32_t), 256);
    w.remv = reinterpret_cast<uint64_t*>(p + off);
    off = align_up(off + (size_t)N * w.col_blocks * sizeof(uint64_t), 256);
    w.mask = reinterpret_cast<uint64_t*>(p + off);
    off = align_up(off + (size_t)N * w.mask_stride_img * sizeof(uint64_t), 256);
    w.total = off;
    return w;
}

}  // namespace cldet

using namespace cldet;

extern "C" {

int cldet_decode_boxes(const float* d_anchors, const float* d_reg, int num_images, int64_t num_anchors, int clip, int height,
                       int width, float* d_boxes, void* stream) {
    if (!d_anchors || !d_reg || !d_boxes || num_images <= 0 || num_anchors <= 0) return CLDET_ERR_INVALID_ARGUMENT;
    const int64_t total = (int64_t)num_images * num_anchors;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
    if (clip)
        decode_boxes_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(d_anchors), reinterpret_cast<const float4*>(d_reg), num_anchors, total, (float)width,
            (float)height, reinterpret_cast<float4*>(d_boxes));
    else
        decode_boxes_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(d_anchors), reinterpret_cast<const float4*>(d_reg), num_anchors, total, (float)width,
            (float)height, reinterpret_cast<float4*>(d_boxes));
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

int cldet_clip_boxes(float* d_boxes, int64_t num_boxes, int height, int width, void* stream) {
    if (!d_boxes || num_boxes < 0) return CLDET_ERR_INVALID_ARGUMENT;
    if (num_boxes == 0) return CLDET_OK;
    const int blocks = (int)std::min<int64_t>((num_boxes + 255) / 256, (int64_t)sm_count() * 16);
    clip_boxes_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(d_boxes), num_boxes, (float)width,
                                                                (float)height);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

int cldet_decode_filter(const float* d_cls, int is_logits, const float* d_reg, const float* d_anchors, int num_images,
                        int64_t num_anchors, int num_classes, int height, int width, float score_thresh,
                        cldet_candidate* d_candidates, uint64_t* d_keys, int64_t capacity, int32_t* d_counts, void* stream) {
    if (!d_cls || !d_reg || !d_anchors || !d_candidates || !d_keys || !d_counts) return CLDET_ERR_INVALID_ARGUMENT;
    if (num_images <= 0 || num_images > 65535 || num_anchors <= 0 || num_classes <= 0 || capacity <= 0)
        return CLDET_ERR_INVALID_ARGUMENT;
    if (num_classes >= kFilterMaxPartials) return CLDET_ERR_UNSUPPORTED;
    const int vec = (num_classes % 4 == 0 && (((uintptr_t)d_cls) & 15) == 0) ? 4 : 1;
    const int ppr = (num_classes + vec - 1) / vec;
    const int stride = ppr | 1;                  // odd row stride: phase 2 reads are bank-conflict free
    // one anchor row per thread in phase 2; fewer rows only when a row's partials would not fit in shared memory
    const int rows = std::min(kFilterThreads, std::max(1, kFilterMaxPartials / stride));
    // a pre-filter on the RAW maximum: a row can only pass if its best class probability exceeds the threshold
    float prefilter;
    if (is_logits) {
        if (score_thresh <= 0.0f) prefilter = -INFINITY;
        else if (score_thresh >= 1.0f) prefilter = INFINITY;
        else prefilter = (float)(log((double)score_thresh / (1.0 - (double)score_thresh)) - 0.01);
    } else {
        prefilter = score_thresh;                // exact: the score IS the raw maximum
    }
    dim3 grid((unsigned)((num_anchors + rows - 1) / rows), (unsigned)num_images);
    const size_t smem = (size_t)rows * stride * (sizeof(float) + 1) + 16;      // per-vector maxima + per-vector info bytes
    cudaStream_t s = (cudaStream_t)stream;
    if (vec == 4)
        decode_filter_kernel<4><<<grid, kFilterThreads, smem, s>>>(
            d_cls, is_logits, reinterpret_cast<const float4*>(d_reg), reinterpret_cast<const float4*>(d_anchors), num_anchors,
            num_classes, rows, stride, (float)width, (float)height, score_thresh, prefilter, d_candidates, d_keys, capacity,
            d_counts, k4_block_append());
    else
        decode_filter_kernel<1><<<grid, kFilterThreads, smem, s>>>(
            d_cls, is_logits, reinterpret_cast<const float4*>(d_reg), reinterpret_cast<const float4*>(d_anchors), num_anchors,
            num_classes, rows, stride, (float)width, (float)height, score_thresh, prefilter, d_candidates, d_keys, capacity,
            d_counts, k4_block_append());
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

int cldet_decode_filter_head(const float* const* h_cls_levels, const float* const* h_reg_levels, int num_levels,
                             int image_height, int image_width, int is_logits, const float* d_anchors, int num_images,
                             int num_classes, float score_thresh, cldet_candidate* d_candidates, uint64_t* d_keys,
                             int64_t capacity, int32_t* d_counts, void* stream) {
    if (!h_cls_levels || !h_reg_levels || !d_anchors || !d_candidates || !d_keys || !d_counts) return CLDET_ERR_INVALID_ARGUMENT;
    if (num_levels != kNumLevels || image_height <= 0 || image_width <= 0) return CLDET_ERR_INVALID_ARGUMENT;
    if (num_images <= 0 || num_images > 65535 || num_classes <= 0 || capacity <= 0) return CLDET_ERR_INVALID_ARGUMENT;
    HeadFilterLevels lv;
    lv.n = num_levels;
    int64_t aoff = 0;
    int coff = 0;
    for (int l = 0; l < kHeadFilterMaxLevels; ++l) {
        lv.cls[l] = lv.reg[l] = nullptr;
        lv.hw[l] = 0;
        lv.pos_chunks[l] = 1;
    }
    for (int l = 0; l < num_levels; ++l) {
        const int sh = 3 + l;
        const int64_t hw = (int64_t)((image_height + (1 << sh) - 1) >> sh) * ((image_width + (1 << sh) - 1) >> sh);
        if (hw <= 0 || hw > (1 << 26) || !h_cls_levels[l] || !h_reg_levels[l]) return CLDET_ERR_INVALID_ARGUMENT;
        lv.cls[l] = h_cls_levels[l];
        lv.reg[l] = h_reg_levels[l];
        lv.hw[l] = (int)hw;
        lv.anchor_off[l] = aoff;
        lv.chunk_off[l] = coff;
        lv.pos_chunks[l] = (int)((hw + kHeadFilterPos - 1) / kHeadFilterPos);
        aoff += hw * kAnchorsPerCell;
        coff += kAnchorsPerCell * lv.pos_chunks[l];
    }
    for (int l = num_levels; l <= kHeadFilterMaxLevels; ++l) {
        lv.anchor_off[l] = aoff;
        lv.chunk_off[l] = coff;
    }
    float prefilter;
    if (is_logits) {
        if (score_thresh <= 0.0f) prefilter = -INFINITY;
        else if (score_thresh >= 1.0f) prefilter = INFINITY;
        else prefilter = (float)(log((double)score_thresh / (1.0 - (double)score_thresh)) - 0.01);
    } else {
        prefilter = score_thresh;
    }
    dim3 grid((unsigned)coff, (unsigned)num_images);
    decode_filter_head_kernel<<<grid, kHeadFilterPos, 0, (cudaStream_t)stream>>>(
        lv, is_logits, reinterpret_cast<const float4*>(d_anchors), aoff, num_classes, (float)image_width, (float)image_height,
        score_thresh, prefilter, d_candidates, d_keys, capacity, d_counts);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

size_t cldet_sort_workspace_bytes(int num_images, int64_t max_count, int topk) {
    if (num_images <= 0 || max_count < 0) return 0;
    size_t off = align_up((size_t)num_images * 4 * sizeof(uint32_t), 256);            // select state
    off = align_up(off + (size_t)num_images * kSelBins * sizeof(uint32_t), 256);      // histograms
    off = align_up(off + (size_t)num_images * sizeof(uint32_t), 256);                 // per-image "blocks done" counters
    if (topk > 0) {                                                                  // compacted survivors (with ties)
        off = align_up(off + (size_t)num_images * max_count * sizeof(cldet_candidate), 256);
        off = align_up(off + (size_t)num_images * max_count * sizeof(uint64_t), 256);
    } else if (max_count > kRadixMin) {                                              // radix sort: key + index ping-pong buffers
        off = align_up(off + (size_t)num_images * 2 * max_count * sizeof(uint64_t), 256);
        off = align_up(off + (size_t)num_images * 2 * max_count * sizeof(uint32_t), 256);
    }
    return off + 256;
}

int cldet_sort_candidates(const cldet_candidate* d_candidates, const uint64_t* d_keys, const int32_t* d_counts,
                          int num_images, int64_t capacity, int64_t max_count, int topk, cldet_candidate* d_sorted,
                          int64_t sorted_capacity, int32_t* d_sorted_counts, void* d_workspace, size_t workspace_bytes,
                          void* stream) {
    if (!d_candidates || !d_keys || !d_counts || !d_sorted || !d_sorted_counts || !d_workspace) return CLDET_ERR_INVALID_ARGUMENT;
    if (num_images <= 0 || num_images > 65535 || capacity <= 0 || sorted_capacity <= 0) return CLDET_ERR_INVALID_ARGUMENT;
    if (max_count > capacity) max_count = capacity;
    if (max_count <= 0) {
        CLDET_CUDA_TRY(cudaMemsetAsync(d_sorted_counts, 0, sizeof(int32_t) * num_images, (cudaStream_t)stream));
        return CLDET_OK;
    }
    if (workspace_bytes < cldet_sort_workspace_bytes(num_images, max_count, topk)) return CLDET_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t s = (cudaStream_t)stream;
    char* p = reinterpret_cast<char*>(d_workspace);
    uint32_t* state = reinterpret_cast<uint32_t*>(p);
    size_t off = align_up((size_t)num_images * 4 * sizeof(uint32_t), 256);
    uint32_t* hist = reinterpret_cast<uint32_t*>(p + off);
    off = align_up(off + (size_t)num_images * kSelBins * sizeof(uint32_t), 256);
    uint32_t* done = reinterpret_cast<uint32_t*>(p + off);
    off = align_up(off + (size_t)num_images * sizeof(uint32_t), 256);
    const int blocks_x = (int)std::min<int64_t>((max_count + 255) / 256, 4096);

    if (topk > 0) {
        cldet_candidate* sel_cand = reinterpret_cast<cldet_candidate*>(p + off);
        off = align_up(off + (size_t)num_images * max_count * sizeof(cldet_candidate), 256);
        uint64_t* sel_keys = reinterpret_cast<uint64_t*>(p + off);
        if (!select_multi_launch()) {
            // one launch: init + three radix passes + compaction, one cluster of kSelCluster CTAs per image
            CLDET_CUDA_TRY(launch_pdl(select_fused_kernel, dim3(num_images * kSelCluster), dim3(kSelThreads), 0, s, d_candidates, d_keys,
                                      d_counts, capacity, topk, state, sel_cand, sel_keys, max_count));
        } else {
            select_init_kernel<<<num_images, 256, 0, s>>>(d_counts, capacity, topk, state, hist, done, num_images);
            CLDET_LAUNCH_CHECK();
            const SelPass passes[3] = {{21, 11, 0}, {10, 11, 11}, {0, 10, 22}};
            for (int ps = 0; ps < 3; ++ps) {
                dim3 g((unsigned)std::max(1, std::min(blocks_x / 8, 16)), (unsigned)num_images);
                select_hist_pick_kernel<<<g, 256, 0, s>>>(d_keys, d_counts, capacity, passes[ps], state, hist, done);
                CLDET_LAUNCH_CHECK();
            }
            const int64_t per_block = 256 * kCompactPerThread;
            dim3 gc((unsigned)((max_count + per_block - 1) / per_block), (unsigned)num_images);
            select_compact_kernel<<<gc, 256, 0, s>>>(d_candidates, d_keys, d_counts, capacity, state, sel_cand, sel_keys, max_count);
            CLDET_LAUNCH_CHECK();
        }
        // survivors are ~topk (plus exact score ties): size the grid for 2*topk, the kernel strides if there are more
        dim3 gr((unsigned)std::max<int64_t>(1, std::min<int64_t>((max_count + kRankPerBlock - 1) / kRankPerBlock,
                                                                   (2 * (int64_t)topk + kRankPerBlock - 1) / kRankPerBlock)),
                (unsigned)num_images);
        CLDET_CUDA_TRY(launch_pdl(rank_sort_kernel, gr, dim3(256), 0, s, (const cldet_candidate*)sel_cand, (const uint64_t*)sel_keys,
                                  (const uint32_t*)state, d_counts, max_count, topk, d_sorted, sorted_capacity, d_sorted_counts,
                                  d_candidates, d_keys, capacity, (int64_t)0, bucket_n()));
    } else if (max_count > kRadixMin && !force_rank_sort()) {
        // the reference's mode (no top-k) with long lists: one launch, one CTA per image, every radix pass inside it
        uint64_t* kbuf = reinterpret_cast<uint64_t*>(p + off);
        off = align_up(off + (size_t)num_images * 2 * max_count * sizeof(uint64_t), 256);
        uint32_t* ibuf = reinterpret_cast<uint32_t*>(p + off);
        // the host only knows an upper bound of the counts: short lists (the trained-model regime) go to the pairwise rank
        // sort, whose grid is sized for kRadixMin candidates; lists longer than that are skipped there and ordered by the
        // radix kernel -- each image takes exactly one of the two
        dim3 gshort((unsigned)((kRadixMin + kRankPerBlock - 1) / kRankPerBlock), (unsigned)num_images);
        CLDET_CUDA_TRY(launch_pdl(rank_sort_kernel, gshort, dim3(256), 0, s, d_candidates, d_keys, (const uint32_t*)nullptr, d_counts,
                                  capacity, 0, d_sorted, sorted_capacity, d_sorted_counts, (const cldet_candidate*)nullptr,
                                  (const uint64_t*)nullptr, (int64_t)0, (int64_t)kRadixMin, bucket_n()));
        CLDET_CUDA_TRY(launch_pdl(radix_sort_kernel, dim3(num_images), dim3(kRsThreads), 0, s, d_candidates, d_keys, d_counts, capacity,
                                  max_count, d_sorted, sorted_capacity, d_sorted_counts, kbuf, ibuf, (int64_t)kRadixMin));
    } else {
        dim3 gc((unsigned)std::min<int64_t>((max_count + kRankPerBlock - 1) / kRankPerBlock, 65535 * 16), (unsigned)num_images);
            CLDET_CUDA_TRY(launch_pdl(rank_sort_kernel, gc, dim3(256), 0, s, d_candidates, d_keys, (const uint32_t*)nullptr, d_counts, capacity,
                                      0, d_sorted, sorted_capacity, d_sorted_counts, (const cldet_candidate*)nullptr,
                                      (const uint64_t*)nullptr, (int64_t)0, (int64_t)0, bucket_n()));
    }
    return CLDET_OK;
}

size_t cldet_nms_workspace_bytes(int num_images, int64_t max_count) {
    if (num_images <= 0 || max_count < 0) return 0;
    return nms_ws_layout(nullptr, num_images, max_count).total + 256;
}

static int nms_sorted_impl(const cldet_candidate* d_sorted, const int32_t* d_sorted_counts, int num_images, int64_t capacity,
                           int64_t max_count, float iou_thresh, int mode, int64_t vanilla_numel_limit, int32_t* d_keep,
                           int32_t* d_keep_counts, void* d_workspace, size_t workspace_bytes, void* stream, float* d_scores,
                           int64_t* d_labels, float* d_boxes);

int cldet_nms_sorted(const cldet_candidate* d_sorted, const int32_t* d_sorted_counts, int num_images, int64_t capacity,
                     int64_t max_count, float iou_thresh, int mode, int64_t vanilla_numel_limit, int32_t* d_keep,
                     int32_t* d_keep_counts, void* d_workspace, size_t workspace_bytes, void* stream) {
    return nms_sorted_impl(d_sorted, d_sorted_counts, num_images, capacity, max_count, iou_thresh, mode, vanilla_numel_limit, d_keep,
                           d_keep_counts, d_workspace, workspace_bytes, stream, nullptr, nullptr, nullptr);
}

int cldet_nms_gather_sorted(const cldet_candidate* d_sorted, const int32_t* d_sorted_counts, int num_images, int64_t capacity,
                            int64_t max_count, float iou_thresh, int mode, int64_t vanilla_numel_limit, int32_t* d_keep,
                            int32_t* d_keep_counts, float* d_scores, int64_t* d_labels, float* d_boxes, void* d_workspace,
                            size_t workspace_bytes, void* stream) {
    if (!d_scores || !d_labels || !d_boxes) return CLDET_ERR_INVALID_ARGUMENT;
    return nms_sorted_impl(d_sorted, d_sorted_counts, num_images, capacity, max_count, iou_thresh, mode, vanilla_numel_limit, d_keep,
                           d_keep_counts, d_workspace, workspace_bytes, stream, d_scores, d_labels, d_boxes);
}

static int nms_sorted_impl(const cldet_candidate* d_sorted, const int32_t* d_sorted_counts, int num_images, int64_t capacity,
                           int64_t max_count, float iou_thresh, int mode, int64_t vanilla_numel_limit, int32_t* d_keep,
                           int32_t* d_keep_counts, void* d_workspace, size_t workspace_bytes, void* stream, float* d_scores,
                           int64_t* d_labels, float* d_boxes) {
    if (!d_sorted || !d_sorted_counts || !d_keep || !d_keep_counts || !d_workspace) return CLDET_ERR_INVALID_ARGUMENT;
    if (num_images <= 0 || num_images > 65535 || capacity <= 0 || mode < 0 || mode > 2) return CLDET_ERR_INVALID_ARGUMENT;
    if (max_count > capacity) max_count = capacity;
    cudaStream_t s = (cudaStream_t)stream;
    if (max_count <= 0) {
        CLDET_CUDA_TRY(cudaMemsetAsync(d_keep_counts, 0, sizeof(int32_t) * num_images, s));
        return CLDET_OK;
    }
    if (workspace_bytes < cldet_nms_workspace_bytes(num_images, max_count)) return CLDET_ERR_WORKSPACE_TOO_SMALL;
    if (max_count <= kFuseMax && nms_fused_enabled()) {
        // short lists (the pre-NMS top-k regime): prepare + mask + resolve + gather in one launch, one cluster per image
        CLDET_CUDA_TRY(cudaFuncSetAttribute(nms_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFuseSmemBytes));
        CLDET_CUDA_TRY(launch_pdl(nms_fused_kernel, dim3((unsigned)num_images * kFuseCluster), dim3(kFuseThreads), kFuseSmemBytes, s, d_sorted,
                                  d_sorted_counts, capacity, mode, vanilla_numel_limit, iou_thresh, (int)((max_count + 63) / 64 * 64), d_keep,
                                  d_keep_counts, d_scores, d_labels,
                                  reinterpret_cast<float4*>(d_boxes)));
        return CLDET_OK;
    }
    const NmsWs w = nms_ws_layout(d_workspace, num_images, max_count);
    if (w.col_blocks > 65535) return CLDET_ERR_UNSUPPORTED;
    CLDET_CUDA_TRY(launch_pdl(nms_prepare_kernel, dim3(num_images), dim3(kPrepThreads), 0, s, d_sorted, d_sorted_counts, capacity, mode,
                              vanilla_numel_limit, w.info));
    const long long tiles = (long long)w.col_blocks * (w.col_blocks + 1) / 2;          // upper-triangular tiles per image
    if (tiles > 2147483647ll) return CLDET_ERR_UNSUPPORTED;
    // blocks stride over the tile index space: a generous capacity does not cost empty blocks
    dim3 grid((unsigned)std::min<long long>(tiles, std::max<long long>(1, (long long)sm_count() * 64 / num_images)), (unsigned)num_images);
    CLDET_CUDA_TRY(launch_pdl(nms_mask_kernel, grid, dim3(256), 0, s, d_sorted, d_sorted_counts, capacity, (const uint32_t*)w.info, iou_thresh,
                              w.mask, w.mask_stride_img, w.col_blocks));
    const int resolve = resolve_choice();          // 0 default, 1 stream, 2 whole mask in shared memory, 3 legacy (A/B experiments)
    if (resolve == 0 || resolve == 1 || max_count > kSmemResolveMax) {          // default: the pipelined kernel, for every K
        const size_t smem = (size_t)w.col_blocks * sizeof(unsigned long long);
        if (smem > 200 * 1024) return CLDET_ERR_UNSUPPORTED;
        if (smem > 40 * 1024)
            CLDET_CUDA_TRY(cudaFuncSetAttribute(nms_resolve_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CLDET_CUDA_TRY(launch_pdl(nms_resolve_stream_kernel, dim3(num_images), dim3(kStreamThreads), smem, s, d_sorted_counts, capacity,
                                  (const uint64_t*)w.mask, w.mask_stride_img, w.col_blocks, d_keep, d_keep_counts, d_sorted, d_scores,
                                  d_labels, reinterpret_cast<float4*>(d_boxes)));
    } else if (max_count <= kSmemResolveMax && resolve != 3) {
        const size_t cbm = (size_t)((max_count + 63) / 64);
        const size_t smem = 64 * cbm * cbm * sizeof(uint64_t);
        if (smem > 48 * 1024)
            CLDET_CUDA_TRY(cudaFuncSetAttribute(nms_resolve_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CLDET_CUDA_TRY(launch_pdl(nms_resolve_smem_kernel, dim3(num_images), dim3(kSmemResolveThreads), smem, s, d_sorted_counts, capacity,
                                  (const uint64_t*)w.mask, w.mask_stride_img, w.col_blocks, d_keep, d_keep_counts, d_sorted, d_scores,
                                  d_labels, reinterpret_cast<float4*>(d_boxes)));
    } else {
        nms_resolve_kernel<<<num_images, 256, 0, s>>>(d_sorted_counts, capacity, w.mask, w.mask_stride_img, w.col_blocks, w.remv,
                                                      d_keep, d_keep_counts);
        CLDET_LAUNCH_CHECK();
        if (d_scores) return cldet_gather_detections(d_sorted, d_keep, d_keep_counts, num_images, capacity, max_count, d_scores, d_labels,
                                                     d_boxes, stream);
    }
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

int cldet_gather_detections(const cldet_candidate* d_sorted, const int32_t* d_keep, const int32_t* d_keep_counts,
                            int num_images, int64_t capacity, int64_t max_keep, float* d_scores, int64_t* d_labels,
                            float* d_boxes, void* stream) {
    if (!d_sorted || !d_keep || !d_keep_counts || !d_scores || !d_labels || !d_boxes) return CLDET_ERR_INVALID_ARGUMENT;
    if (num_images <= 0 || num_images > 65535 || capacity <= 0) return CLDET_ERR_INVALID_ARGUMENT;
    if (max_keep > capacity) max_keep = capacity;
    if (max_keep <= 0) return CLDET_OK;
    dim3 grid((unsigned)((max_keep + 255) / 256), (unsigned)num_images);
    gather_detections_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_sorted, d_keep, d_keep_counts, capacity, d_scores,
                                                                    d_labels, reinterpret_cast<float4*>(d_boxes));
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

int cldet_coco_results(const float* d_scores, const int64_t* d_labels, const float* d_boxes, const int32_t* d_counts,
                       const float* d_scales, int num_images, int64_t capacity, float score_threshold,
                       cldet_coco_record* d_records, int32_t* d_offsets, void* stream) {
    if (!d_scores || !d_labels || !d_boxes || !d_counts || !d_scales || !d_records || !d_offsets) return CLDET_ERR_INVALID_ARGUMENT;
    if (num_images <= 0 || num_images > 65535 || capacity <= 0) return CLDET_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    coco_offsets_kernel<<<1, 256, 0, s>>>(d_scores, d_counts, num_images, capacity, score_threshold, d_offsets);
    CLDET_LAUNCH_CHECK();
    dim3 grid((unsigned)std::min<int64_t>((capacity + 255) / 256, 64), (unsigned)num_images);
    coco_records_kernel<<<grid, 256, 0, s>>>(d_scores, d_labels, reinterpret_cast<const float4*>(d_boxes), d_counts, d_scales,
                                             capacity, score_threshold, d_offsets, d_records);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

size_t cldet_batched_nms_workspace_bytes(int64_t num_boxes) {
    if (num_boxes <= 0) return 256;
    size_t off = 0;
    off = align_up(off + (size_t)num_boxes * sizeof(cldet_candidate), 256);   // packed
    off = align_up(off + (size_t)num_boxes * sizeof(uint64_t), 256);          // keys
    off = align_up(off + (size_t)num_boxes * sizeof(cldet_candidate), 256);   // sorted
    off = align_up(off + (size_t)num_boxes * sizeof(int32_t), 256);           // keep positions
    off = align_up(off + 4 * sizeof(int32_t), 256);                           // counts
    if (num_boxes > kRadixMin) {                                              // radix sort ping-pong buffers
        off = align_up(off + (size_t)2 * num_boxes * sizeof(uint64_t), 256);
        off = align_up(off + (size_t)2 * num_boxes * sizeof(uint32_t), 256);
    }
    off += cldet_nms_workspace_bytes(1, num_boxes);
    return off + 256;
}

int cldet_batched_nms(const float* d_boxes, const float* d_scores, const int64_t* d_idxs, int64_t num_boxes,
                      float iou_thresh, int mode, int64_t vanilla_numel_limit, int64_t* d_keep, int32_t* d_keep_count,
                      void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!d_keep_count || !d_workspace || num_boxes < 0 || mode < 0 || mode > 2) return CLDET_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    if (num_boxes == 0) {
        CLDET_CUDA_TRY(cudaMemsetAsync(d_keep_count, 0, sizeof(int32_t), s));
        return CLDET_OK;
    }
    if (!d_boxes || !d_scores || !d_keep) return CLDET_ERR_INVALID_ARGUMENT;
    if (num_boxes > 0x7fffffff) return CLDET_ERR_UNSUPPORTED;
    if (workspace_bytes < cldet_batched_nms_workspace_bytes(num_boxes)) return CLDET_ERR_WORKSPACE_TOO_SMALL;
    char* p = reinterpret_cast<char*>(d_workspace);
    size_t off = 0;
    cldet_candidate* packed = reinterpret_cast<cldet_candidate*>(p + off);
    off = align_up(off + (size_t)num_boxes * sizeof(cldet_candidate), 256);
    uint64_t* keys = reinterpret_cast<uint64_t*>(p + off);
    off = align_up(off + (size_t)num_boxes * sizeof(uint64_t), 256);
    cldet_candidate* sorted = reinterpret_cast<cldet_candidate*>(p + off);
    off = align_up(off + (size_t)num_boxes * sizeof(cldet_candidate), 256);
    int32_t* keep_pos = reinterpret_cast<int32_t*>(p + off);
    off = align_up(off + (size_t)num_boxes * sizeof(int32_t), 256);
    int32_t* cnts = reinterpret_cast<int32_t*>(p + off);   // [0] packed count, [1] sorted count
    off = align_up(off + 4 * sizeof(int32_t), 256);
    uint64_t* kbuf = nullptr;
    uint32_t* ibuf = nullptr;
    if (num_boxes > kRadixMin) {
        kbuf = reinterpret_cast<uint64_t*>(p + off);
        off = align_up(off + (size_t)2 * num_boxes * sizeof(uint64_t), 256);
        ibuf = reinterpret_cast<uint32_t*>(p + off);
        off = align_up(off + (size_t)2 * num_boxes * sizeof(uint32_t), 256);
    }
    void* nms_ws = p + off;
    const unsigned blocks = (unsigned)((num_boxes + 255) / 256);
    pack_boxes_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(d_boxes), d_scores, d_idxs, num_boxes, packed, keys,
                                             cnts);
    CLDET_LAUNCH_CHECK();
    if (kbuf && !force_rank_sort()) {
        radix_sort_kernel<<<1, kRsThreads, 0, s>>>(packed, keys, cnts, num_boxes, num_boxes, sorted, num_boxes, cnts + 1, kbuf, ibuf);
    } else {
        dim3 g((unsigned)((num_boxes + kRankPerBlock - 1) / kRankPerBlock), 1);
        rank_sort_kernel<<<g, 256, 0, s>>>(packed, keys, nullptr, cnts, num_boxes, 0, sorted, num_boxes, cnts + 1, nullptr, nullptr, 0, 0,
                                           bucket_n());
    }
    CLDET_LAUNCH_CHECK();
    int rc = cldet_nms_sorted(sorted, cnts + 1, 1, num_boxes, num_boxes, iou_thresh, d_idxs ? mode : 1, vanilla_numel_limit,
                              keep_pos, d_keep_count, nms_ws, workspace_bytes - off, stream);
    if (rc) return rc;
    keep_to_original_kernel<<<blocks, 256, 0, s>>>(sorted, keep_pos, d_keep_count, d_keep);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

}  // extern "C"
