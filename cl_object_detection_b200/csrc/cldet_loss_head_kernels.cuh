// K3h: the fused focal + smooth-L1 forward/backward reading the detection head's RAW conv outputs, one tensor per pyramid
// level in the layout the convolutions produce: classification [N, 9*C, H_l, W_l] and regression [N, 9*4, H_l, W_l] (NCHW).
//
// Reference: ClassificationModel.forward / RegressionModel.forward (retinanet/model.py:133-214) turn every level's output into
// [N, H_l*W_l*9, C] with permute(0,2,3,1) + contiguous() + view, ResNet.forward concatenates the five levels
// (model.py:472-474), and autograd mirrors all of it in backward: 8 B/element for the permute copy, 8 for the cat, and the
// same again for their gradients -- more traffic than the loss itself (8 B/element).  SURVEY 8(f) row f1 names this as the
// step immediately before the path.  Here the layout change is index arithmetic: the element of anchor (y, x, k) and class c
// lives at channel k*C + c, position y*W_l + x of its level's plane, and its gradient is written to the same place of a
// gradient tensor of the same layout, which is exactly what the output convolution's backward wants.
//
// Work decomposition: a block owns kHeadPos consecutive positions of ONE anchor type k of ONE level of ONE image: C channel
// rows (k*C .. k*C+C-1) of kHeadPos contiguous floats (2 KB) each.  Consecutive blocks walk the channel planes in address
// order, so the blocks in flight touch ~100 MB of contiguous memory like the concatenated-layout kernel does (a first version
// gave each block all 9*C channels of 64 positions: 720 rows 67 KB apart per block, twice as slow -- TLB reach).  One warp
// sweeps one row at a time with 128-bit accesses, four in flight per lane (planes whose size is not a multiple of 4 floats --
// the tiny top levels -- use 32-bit accesses).  The assignment words of the block's anchors sit in shared memory by position; a
// bitmask marks the positions whose anchor is plain background, so the hot path (four background elements) costs one mask
// test per vector.  The 9 type-blocks of a position range read the same key / word sectors; L2 serves the repeats.
#pragma once
#include "cldet_loss_kernels.cuh"

namespace cldet {

#ifndef CLDET_HEAD_POS
#define CLDET_HEAD_POS 512
#endif
constexpr int kHeadPos = CLDET_HEAD_POS;  // positions per block (a multiple of 128)
static_assert(kHeadPos % 128 == 0 && kHeadPos <= 1024, "kHeadPos: whole 128-position groups, at most 8 vectors per lane");
constexpr int kHeadMaskWords = kHeadPos / 32;
#ifndef CLDET_HEAD_MINBLOCKS
#define CLDET_HEAD_MINBLOCKS 5
#endif
// Bulk-copy (TMA) sweep: every warp owns a ring of kHeadStages row buffers (kHeadPos floats each) in dynamic shared memory
#ifndef CLDET_HEAD_STAGES
#define CLDET_HEAD_STAGES 3
#endif
#ifndef CLDET_HEAD_TMA_MINBLOCKS
#define CLDET_HEAD_TMA_MINBLOCKS 4
#endif
constexpr int kHeadStages = CLDET_HEAD_STAGES;
constexpr int kHeadWarps = kLossThreads / 32;
constexpr size_t kHeadStageBytes = (size_t)kHeadWarps * kHeadStages * kHeadPos * sizeof(float);
constexpr int kHeadMaxLevels = 8;
constexpr int kHeadTypes = 9;           // anchors per position (3 ratios x 3 scales, retinanet/anchors.py:10-19)

struct HeadLevels {
    int n;                                       // pyramid levels
    const float* cls[kHeadMaxLevels];            // [N, 9*C, hw]
    const float* reg[kHeadMaxLevels];            // [N, 36, hw]
    float* gcls[kHeadMaxLevels];
    float* greg[kHeadMaxLevels];
    int hw[kHeadMaxLevels];                      // H_l * W_l
    int64_t anchor_off[kHeadMaxLevels + 1];      // level offsets inside the concatenated anchor index space
    int chunk_off[kHeadMaxLevels + 1];           // level offsets inside the per-image chunk index space
    int pos_chunks[kHeadMaxLevels];              // ceil(hw / kHeadPos); a level has 9 * pos_chunks chunks, type-major
};

// One element with target 0 in the general (gamma != 2 or IL variants) configuration, or target 1: defer to cls_element.
// mode 0: losses + gradients; mode 1: gradients only (backward with different upstream weights).
template <bool GAMMA2, bool VARIANTS, bool GRAD, bool LOGITS>
__device__ __forceinline__ float head_element(float x, int c, uint32_t m, int64_t anchor_abs, const LossArgs& a,
                                              const ImageScales& sc, float as_bg, bool need_iou, Acc& acc, int lane4) {
    const uint32_t st = meta_state(m);
    if (st == CLDET_STATE_IGNORE) return 0.0f;
    const float p = LOGITS ? sigmoid_exact(x) : x;
    float g;
    const bool target1 = (st == CLDET_STATE_POS) && ((uint32_t)c == meta_label(m));
    if (GAMMA2 && !VARIANTS && !target1) {
        g = neg_element_raw<GRAD>(p, as_bg, acc.raw);
    } else {
        float iou = 1.0f;
        if (need_iou && st == CLDET_STATE_POS) iou = a.iou_max[anchor_abs];
        g = cls_element<GAMMA2, VARIANTS, GRAD>(p, c, m, a, sc, iou, acc);
    }
    if (GRAD && LOGITS) g = sigmoid_bwd(g, p);
    return g;
}

// Target-0 PAIR with 0/1 weights (0: the element is ignored or is a positive anchor's own class, handled separately):
// neg_pair_raw with the weight folded into the loss term and the gradient scale -- no branch, packed arithmetic.
template <bool GRAD>
__device__ __forceinline__ f32x2 neg_pair_raw_w(float p0_raw, float p1_raw, float as, f32x2 w, f32x2& rawn) {
    const float p0 = fminf(fmaxf(p0_raw, 1e-4f), 0.9999f);
    const float p1 = fminf(fmaxf(p1_raw, 1e-4f), 0.9999f);
    const f32x2 p = pk2(p0, p1);
    const f32x2 q = fma2(p, bc2(-1.0f), bc2(1.0f));
    float q0, q1;
    upk2(q, q0, q1);
    const f32x2 L = log_fast2(q0, q1);
    rawn = fma2(mul2(mul2(p, p), w), L, rawn);
    if (!GRAD) return 0ull;
    const f32x2 t = fma2(L, bc2(-2.0f), mul2(p, pk2(__frcp_rn_fast(q0), __frcp_rn_fast(q1))));
    const f32x2 g = mul2(mul2(mul2(bc2(as), w), p), t);
    float g0, g1;
    upk2(g, g0, g1);
    return pk2((p0 == p0_raw) ? g0 : 0.0f, (p1 == p1_raw) ? g1 : 0.0f);
}

// Four consecutive positions of one class row whose anchors are ALL plain background (the common case by far): the bare
// packed pairs, no weights, no look-ups.
template <bool GRAD, bool LOGITS>
__device__ __forceinline__ float4 head_vec_plain(const float4 x, float as_bg, Acc& acc) {
    float p0 = x.x, p1 = x.y, p2 = x.z, p3 = x.w;
    f32x2 pa = 0ull, pb = 0ull;
    if (LOGITS) {
        pa = sigmoid_exact2(p0, p1);
        pb = sigmoid_exact2(p2, p3);
        upk2(pa, p0, p1);
        upk2(pb, p2, p3);
    }
    f32x2 ga = neg_pair_raw<GRAD>(p0, p1, as_bg, acc.rawn[0]);
    f32x2 gb = neg_pair_raw<GRAD>(p2, p3, as_bg, acc.rawn[1]);
    if (GRAD && LOGITS) {
        ga = sigmoid_bwd2(ga, pa);
        gb = sigmoid_bwd2(gb, pb);
    }
    float4 g;
    upk2(ga, g.x, g.y);
    upk2(gb, g.z, g.w);
    return g;
}

// A positive anchor's own class (target 1) when no IL variant is active: f = 1 - p in all three reference branches
// (losses.py:352-366 with decrease_positive == 1).  Rare (one element per positive anchor) and fat (logf): kept out of line,
// arguments by value, so the hot loop's instruction footprint stays small.
static __device__ __noinline__ float2 own_class_element(float p_raw, float alpha, float scale) {
    const float p = fminf(fmaxf(p_raw, 1e-4f), 0.9999f);
    const float f = 1.0f - p;
    const float nl = -logf(p);
    const float pw = f * f;                                   // ATen evaluates pow(x, 2.0) as x*x
    const float loss = (alpha * pw) * nl;
    const float g = alpha * ((2.0f * f) * -1.0f * nl - pw / p) * scale;
    return make_float2(loss, (p == p_raw) ? g : 0.0f);
}

// Four consecutive positions of one class row.  gamma == 2 without IL variants: every element runs the weighted target-0
// form (weight 0 for ignored anchors and for a positive anchor's own class, looked up only for non-plain positions), then the
// rare own-class elements are recomputed out of line.  One code path for all lanes: non-background anchors cluster around
// the GT boxes, so a per-vector branch between a bare and a general path would make most warps execute both, and two inlined
// copies of the math per unrolled vector overflow the instruction cache (measured: 2.4 warps per issue stalled on fetch).
template <bool GAMMA2, bool VARIANTS, bool GRAD, bool LOGITS>
__device__ __forceinline__ float4 head_vec(const float4 x, int c, uint32_t plain4, const uint32_t* mp, int64_t anchor_abs,
                                           const LossArgs& a, const ImageScales& sc, float as_bg, bool need_iou, Acc& acc) {
    float4 g;
    if (GAMMA2 && !VARIANTS) {
        float p0 = x.x, p1 = x.y, p2 = x.z, p3 = x.w;
        f32x2 pa = 0ull, pb = 0ull;
        if (LOGITS) {
            pa = sigmoid_exact2(p0, p1);
            pb = sigmoid_exact2(p2, p3);
            upk2(pa, p0, p1);
            upk2(pb, p2, p3);
        }
        float w0 = 1.0f, w1 = 1.0f, w2 = 1.0f, w3 = 1.0f;
        uint32_t own = 0;                                               // elements that are a positive anchor's own class
        if (plain4 != 0xFu) {
            float wgt[4] = {1.0f, 1.0f, 1.0f, 1.0f};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (!((plain4 >> e) & 1u)) {
                    const uint32_t m = mp[e];                           // not plain background: ignored or positive
                    if (meta_state(m) == CLDET_STATE_IGNORE) {
                        wgt[e] = 0.0f;
                    } else if (meta_state(m) == CLDET_STATE_POS && (uint32_t)c == meta_label(m)) {
                        wgt[e] = 0.0f;
                        own |= 1u << e;
                    }
                }
            }
            w0 = wgt[0]; w1 = wgt[1]; w2 = wgt[2]; w3 = wgt[3];
        }
        const f32x2 ga = neg_pair_raw_w<GRAD>(p0, p1, as_bg, pk2(w0, w1), acc.rawn[0]);
        const f32x2 gb = neg_pair_raw_w<GRAD>(p2, p3, as_bg, pk2(w2, w3), acc.rawn[1]);
        upk2(ga, g.x, g.y);
        upk2(gb, g.z, g.w);
        if (own) {
            const float pv[4] = {p0, p1, p2, p3};
            float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll 1
            for (int e = 0; e < 4; ++e) {
                if ((own >> e) & 1u) {
                    const float2 lg = own_class_element(pv[e], a.p.alpha, sc.s_fg);
                    acc.fg += lg.x;
                    gv[e] = GRAD ? lg.y : 0.0f;
                }
            }
            g = make_float4(gv[0], gv[1], gv[2], gv[3]);
        }
        if (GRAD && LOGITS) {
            upk2(sigmoid_bwd2(pk2(g.x, g.y), pa), g.x, g.y);
            upk2(sigmoid_bwd2(pk2(g.z, g.w), pb), g.z, g.w);
        }
    } else {
        g.x = head_element<GAMMA2, VARIANTS, GRAD, LOGITS>(x.x, c, mp[0], anchor_abs, a, sc, as_bg, need_iou, acc, 0);
        g.y = head_element<GAMMA2, VARIANTS, GRAD, LOGITS>(x.y, c, mp[1], anchor_abs + kHeadTypes, a, sc, as_bg, need_iou, acc, 1);
        g.z = head_element<GAMMA2, VARIANTS, GRAD, LOGITS>(x.z, c, mp[2], anchor_abs + 2 * kHeadTypes, a, sc, as_bg, need_iou, acc, 2);
        g.w = head_element<GAMMA2, VARIANTS, GRAD, LOGITS>(x.w, c, mp[3], anchor_abs + 3 * kHeadTypes, a, sc, as_bg, need_iou, acc, 3);
    }
    return g;
}

// State of one warp's bulk-copy row loop (see head_rows_bulk).
struct BulkRows {
    float* ring;            // this lane's first vector in ring slot 0 (generic address)
    uint32_t buf0, bar0;    // shared-memory addresses of the warp's ring and of its mbarriers
    uint32_t row_bytes;
    int my_rows;            // rows c0, c0 + kHeadWarps, ...
    int64_t stride;         // floats between two of this warp's rows: kHeadWarps * hw
    float* dst_row;         // gradient row of the current iteration
    const float* src_fetch; // class row to fetch next (row index i + kHeadStages - 1)
    int c0;
    int nvec;               // vectors of this lane inside the chunk (ragged chunks)
    uint32_t pm;            // plain-background bits of the lane's kU vectors
    const uint32_t* mp;
    int64_t ab;
};

// The bulk-copy sweep of one warp.  Row i lives in ring slot i % kHeadStages: wait for its bytes, compute the gradients IN
// PLACE in shared memory (every lane rewrites exactly the vectors it read), hand the slot to the copy engine as a bulk store,
// and refill the slot freed one iteration ago with row i + kHeadStages - 1.  No block-wide barrier and no register-held load
// in the loop: kHeadStages - 1 rows per warp are always in flight.  All addresses advance by constants.
template <bool GAMMA2, bool VARIANTS, bool GRAD, bool LOGITS, bool PLAIN, bool FULL>
__device__ __forceinline__ void head_rows_bulk(BulkRows br, const LossArgs& a, const ImageScales& sc, float as_bg, bool need_iou,
                                               Acc& acc) {
    constexpr int kU = kHeadPos / 128;
    constexpr uint32_t kSlotBytes = kHeadPos * 4;
    const bool lane0 = (threadIdx.x & 31) == 0;
    int slot = 0;
    uint32_t parity = 0;
    int c = br.c0;
    for (int i = 0; i < br.my_rows; ++i, c += kHeadWarps) {
        mbar_wait(br.bar0 + 8u * slot, parity);
        float* b = br.ring + slot * kHeadPos;
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            if (FULL || u < br.nvec) {
                float4* v = reinterpret_cast<float4*>(b + 128 * u);
                if (PLAIN) {
                    *v = head_vec_plain<GRAD, LOGITS>(*v, as_bg, acc);
                } else {
                    *v = head_vec<GAMMA2, VARIANTS, GRAD, LOGITS>(*v, c, (br.pm >> (4 * u)) & 0xFu, br.mp + 128 * u,
                                                                 br.ab + (int64_t)(128 * u) * kHeadTypes, a, sc, as_bg, need_iou, acc);
                }
            }
        }
        fence_async_smem();                                      // this lane's writes -> visible to the copy engine
        __syncwarp();
        if (lane0) {
            bulk_store(br.dst_row, br.buf0 + kSlotBytes * slot, br.row_bytes);
            bulk_commit();
            if (i + kHeadStages - 1 < br.my_rows) {              // next row to fetch goes into the slot of row i - 1
                if (i >= 1) bulk_wait_read<1>();                 // ... once the store of row i - 1 has drained it
                const uint32_t rs = (slot == 0) ? kHeadStages - 1 : slot - 1;
                mbar_expect_tx(br.bar0 + 8u * rs, br.row_bytes);
                bulk_load(br.buf0 + kSlotBytes * rs, br.src_fetch, br.row_bytes, br.bar0 + 8u * rs);
            }
        }
        br.dst_row += br.stride;
        br.src_fetch += br.stride;
        if (++slot == kHeadStages) {
            slot = 0;
            parity ^= 1u;
        }
    }
    // the ring must outlive the stores that read it (their global writes complete with the grid)
#ifdef CLDET_HEAD_WAIT_ALL
    if (lane0) bulk_wait_all<0>();
#else
    if (lane0) bulk_wait_read<0>();
#endif
}

// TMA: compile the bulk-copy sweep (fused forward+backward launch only; `stage` = the block's dynamic shared memory,
// `bars` = kHeadWarps * kHeadStages mbarriers).  It is taken for chunks of planes whose rows are 16-byte aligned (H_l*W_l a
// multiple of 4 floats: 94 % of a COCO-shaped batch); the other chunks run the register sweeps below.
template <bool GAMMA2, bool VARIANTS, bool GRAD, bool LOGITS, bool TMA = false>
__device__ __forceinline__ void head_chunk(const LossArgs& a, const HeadLevels& lv, int j, int chunk, int mode,
                                           const ImageScales& sc, Acc& acc, uint32_t* smeta, uint32_t* plain,
                                           float* stage = nullptr, unsigned long long* bars = nullptr) {
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kLossThreads / 32;
    int l = 0;
    while (l + 1 < lv.n && chunk >= lv.chunk_off[l + 1]) ++l;
    const int hw = lv.hw[l];
    const int local = chunk - lv.chunk_off[l];
    const int k = local / lv.pos_chunks[l];                                 // anchor type of this block
    const int p0 = (local - k * lv.pos_chunks[l]) * kHeadPos;
    const int np = min(kHeadPos, hw - p0);
    const int64_t an0 = lv.anchor_off[l] + (int64_t)p0 * kHeadTypes + k;    // anchor of position p0; stride kHeadTypes
    const int C = a.C;
    const bool need_iou = VARIANTS && a.p.incremental && a.p.decrease_positive_by_iou;
    const float* src = lv.cls[l] + ((int64_t)j * (kHeadTypes * C) + (int64_t)k * C) * hw + p0;
    float* dst = GRAD ? lv.gcls[l] + ((int64_t)j * (kHeadTypes * C) + (int64_t)k * C) * hw + p0 : nullptr;
    // 128-bit / bulk accesses need 16-byte aligned addresses; a row starts at float offset (channel * hw + p0) from the tensor base
    const bool base_ok = (((uintptr_t)lv.cls[l] | (uintptr_t)lv.gcls[l]) & 15) == 0;

    // ---- bulk-copy sweep, part 1: this warp's first rows start moving into its ring before the prologue runs ----
    bool use_tma = false;
    int my_rows = 0;
    uint32_t row_bytes = 0, bar0 = 0, buf0 = 0;
    if constexpr (TMA) {
        use_tma = base_ok && (hw & 3) == 0;                              // block-uniform
        if (use_tma) {
            my_rows = warp < C ? (C - warp + kWarps - 1) / kWarps : 0;   // rows warp, warp + kWarps, ...
            row_bytes = (uint32_t)np * 4u;                               // np % 4 == 0 here
            bar0 = smem_u32(bars + warp * kHeadStages);
            buf0 = smem_u32(stage + (size_t)warp * kHeadStages * kHeadPos);
            if (lane == 0) {
#pragma unroll
                for (int st = 0; st < kHeadStages; ++st) mbar_init(bar0 + 8u * st, 1u);
                fence_async_smem();
#pragma unroll
                for (int i = 0; i < kHeadStages - 1; ++i) {
                    if (i < my_rows) {
                        mbar_expect_tx(bar0 + 8u * i, row_bytes);
                        bulk_load(buf0 + (uint32_t)(i * kHeadPos * 4), src + (int64_t)(warp + kWarps * i) * hw, row_bytes, bar0 + 8u * i);
                    }
                }
            }
        }
    }

    // ---- per-anchor prologue: assignment word, outputs keyed by anchor, smooth-L1 for the positives; every position writes its
    // four regression-gradient values (zeros unless positive) itself -- coalesced across the block, no separate zero pass ----
    float* greg_rows = GRAD ? lv.greg[l] + ((int64_t)j * (kHeadTypes * 4) + k * 4) * hw + p0 : nullptr;
    const int nvalid_j = (mode == 0 && a.best) ? a.nvalid[j] : 1;
    const float* reg_rows = lv.reg[l] + ((int64_t)j * (kHeadTypes * 4) + k * 4) * hw + p0;
    constexpr int kRounds = kHeadPos / kLossThreads;                        // uniform trip count: the ballot needs whole warps
    static_assert(kHeadPos % kLossThreads == 0, "kHeadPos must be a multiple of the block size");
    // all of a thread's key / word loads are issued before the first is used: the prologue costs one memory latency
    unsigned long long keys[kRounds];
    uint32_t words[kRounds];
#pragma unroll
    for (int rd = 0; rd < kRounds; ++rd) {
        const int pp = rd * kLossThreads + tid;
        keys[rd] = 0ull;
        words[rd] = 0u;
        if (pp < np) {
            const int64_t gi = (int64_t)j * a.A + an0 + (int64_t)pp * kHeadTypes;
            if (mode == 0 && a.best) keys[rd] = a.best[gi];
            else words[rd] = a.meta[gi];
        }
    }
#pragma unroll
    for (int rd = 0; rd < kRounds; ++rd) {
        const int pp = rd * kLossThreads + tid;
        bool is_plain = pp >= np;                                           // positions past the end count as plain (never swept)
        if (pp < np) {
            const int64_t an = an0 + (int64_t)pp * kHeadTypes;
            const int64_t gi = (int64_t)j * a.A + an;
            uint32_t m = words[rd];
            if (mode == 0 && a.best) {
                const unsigned long long key = keys[rd];
                if (key) a.best[gi] = 0ull;                   // leave the scratch zeroed for the next call
                m = word_from_best(a, j, key, nvalid_j);
                a.meta_out[gi] = m;
                if (a.iou_out) a.iou_out[gi] = __uint_as_float((uint32_t)(key >> 32));
            }
            smeta[pp] = m;
            const uint32_t st = meta_state(m);
            is_plain = (st == CLDET_STATE_BG || st == CLDET_STATE_EMPTY);
            if (mode == 0 && a.bg_mask) a.bg_mask[gi] = (st != CLDET_STATE_POS) ? 1 : 0;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (st == CLDET_STATE_POS) {
                const float* rp = reg_rows + pp;
                const float4 r = make_float4(rp[0], rp[hw], rp[2 * (int64_t)hw], rp[3 * (int64_t)hw]);
                acc.reg += reg_anchor<GRAD>(a, j, an, m, r, sc.s_reg, g);
                if (mode == 0 && a.status && meta_label(m) == CLDET_BAD_LABEL) *a.status = 1;
            }
            if (GRAD) {
                float* gp = greg_rows + pp;
                gp[0] = g.x;
                gp[hw] = g.y;
                gp[2 * (int64_t)hw] = g.z;
                gp[3 * (int64_t)hw] = g.w;
            }
        }
        const uint32_t bits = __ballot_sync(0xffffffffu, is_plain);
        if (lane == 0) plain[pp >> 5] = bits;
    }
    if (tid == 0) plain[kHeadMaskWords] = 0u;
    __syncthreads();

    // ---- classification rows: C rows of np contiguous positions, one warp per row ----
    const float alpha_img = (meta_state(smeta[0]) == CLDET_STATE_EMPTY) ? 1.0f - a.p.alpha : a.p.alpha;
    const float as_bg = alpha_img * sc.s_bg;
    const int64_t abs0 = (int64_t)j * a.A + an0;
    constexpr int kU = kHeadPos / 128;                                   // kU x 32 lanes x 16 B = one full row per round
    constexpr uint32_t kAllPlain = 0xFFFFFFFFu >> (32 - 4 * kU);
    if constexpr (TMA) {
        if (use_tma) {
            // ---- bulk-copy sweep, part 2 (head_rows_bulk): four loop bodies -- all-plain or not, full or ragged chunk --
            // chosen once per block, so the row loop itself carries no per-vector tests.
            uint32_t pm = 0;
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int q = lane + 32 * u;
                pm |= ((plain[q >> 3] >> ((q & 7) * 4)) & 0xFu) << (4 * u);
            }
            const bool all_plain = GAMMA2 && !VARIANTS && __all_sync(0xffffffffu, pm == kAllPlain);
            BulkRows br;
            br.ring = stage + (size_t)warp * kHeadStages * kHeadPos + 4 * lane;
            br.buf0 = buf0;
            br.bar0 = bar0;
            br.row_bytes = row_bytes;
            br.my_rows = my_rows;
            br.stride = (int64_t)kWarps * hw;
            br.dst_row = dst + (int64_t)warp * hw;
            br.src_fetch = src + (int64_t)(warp + kWarps * (kHeadStages - 1)) * hw;
            br.c0 = warp;
            br.nvec = min(kU, max(0, (np - 4 * lane + 127) >> 7));          // this lane's vectors inside the chunk
            br.pm = pm;
            br.mp = smeta + 4 * lane;
            br.ab = abs0 + (int64_t)(4 * lane) * kHeadTypes;
            const bool full = np == kHeadPos;
            if (all_plain) {
                if (full) head_rows_bulk<GAMMA2, VARIANTS, GRAD, LOGITS, true, true>(br, a, sc, as_bg, need_iou, acc);
                else head_rows_bulk<GAMMA2, VARIANTS, GRAD, LOGITS, true, false>(br, a, sc, as_bg, need_iou, acc);
            } else {
                if (full) head_rows_bulk<GAMMA2, VARIANTS, GRAD, LOGITS, false, true>(br, a, sc, as_bg, need_iou, acc);
                else head_rows_bulk<GAMMA2, VARIANTS, GRAD, LOGITS, false, false>(br, a, sc, as_bg, need_iou, acc);
            }
            return;
        }
    }
    if (base_ok && (hw & 3) == 0 && np == kHeadPos) {
        // Full rows of an aligned plane (94 % of a COCO-shaped batch): no bounds checks, and the plain-background bits of this
        // lane's kU vectors do not depend on the class row, so they are looked up once per chunk.
        uint32_t pm = 0;
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int q = lane + 32 * u;
            pm |= ((plain[q >> 3] >> ((q & 7) * 4)) & 0xFu) << (4 * u);
        }
        // Non-background anchors cluster around the GT boxes: for most blocks ALL kHeadPos positions of this anchor type are
        // plain background, and then the sweep is the bare packed math (a warp-uniform choice made once per block, so each
        // warp runs exactly one of the two loops).
        if (GAMMA2 && !VARIANTS && __all_sync(0xffffffffu, pm == kAllPlain)) {
            for (int c = warp; c < C; c += kWarps) {
                const float* sp = src + (int64_t)c * hw + 4 * lane;
                float* dp = GRAD ? dst + (int64_t)c * hw + 4 * lane : nullptr;
                float4 x[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    asm volatile(CLDET_LD_QUAL ".v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(x[u].x), "=f"(x[u].y), "=f"(x[u].z), "=f"(x[u].w)
                                 : "l"(sp + 128 * u));
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const float4 g = head_vec_plain<GRAD, LOGITS>(x[u], as_bg, acc);
                    if (GRAD) {
                        asm volatile(CLDET_ST_QUAL ".v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dp + 128 * u), "f"(g.x), "f"(g.y), "f"(g.z), "f"(g.w)
                                     : "memory");
                    }
                }
            }
            return;
        }
        const uint32_t* mp = smeta + 4 * lane;
        const int64_t ab = abs0 + (int64_t)(4 * lane) * kHeadTypes;
        for (int c = warp; c < C; c += kWarps) {
            const float* sp = src + (int64_t)c * hw + 4 * lane;
            float* dp = GRAD ? dst + (int64_t)c * hw + 4 * lane : nullptr;
            float4 x[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                asm volatile(CLDET_LD_QUAL ".v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(x[u].x), "=f"(x[u].y), "=f"(x[u].z), "=f"(x[u].w)
                             : "l"(sp + 128 * u));
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const float4 g = head_vec<GAMMA2, VARIANTS, GRAD, LOGITS>(x[u], c, (pm >> (4 * u)) & 0xFu, mp + 128 * u,
                                                                         ab + (int64_t)(128 * u) * kHeadTypes, a, sc, as_bg, need_iou, acc);
                if (GRAD) {
                    asm volatile(CLDET_ST_QUAL ".v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dp + 128 * u), "f"(g.x), "f"(g.y), "f"(g.z), "f"(g.w)
                                 : "memory");
                }
            }
        }
        return;
    }
    // Ragged ends and planes that are not a multiple of four floats (the small top levels, ~6 % of a COCO-shaped batch):
    // compact code, two vectors in flight.
    constexpr int kUg = 2;
#pragma unroll 1
    for (int c = warp; c < C; c += kWarps) {
        const int64_t row_off = ((int64_t)j * (kHeadTypes * C) + (int64_t)k * C + c) * hw + p0;
        const float* sp = src + (int64_t)c * hw;
        float* dp = GRAD ? dst + (int64_t)c * hw : nullptr;
        const int peel = base_ok ? min(np, (int)((4 - (row_off & 3)) & 3)) : np;
        const int nv = (np - peel) >> 2;
        const int tail0 = peel + 4 * nv;
        // leading / trailing scalars (at most 3 + 3 when the base is aligned)
#pragma unroll 1
        for (int pp = lane; pp < peel + (np - tail0); pp += 32) {
            const int pos = pp < peel ? pp : tail0 + (pp - peel);
            const float g = head_element<GAMMA2, VARIANTS, GRAD, LOGITS>(sp[pos], c, smeta[pos], abs0 + (int64_t)pos * kHeadTypes, a, sc,
                                                                         as_bg, need_iou, acc, 0);
            if (GRAD) dp[pos] = g;
        }
#pragma unroll 1
        for (int q0 = lane; q0 < nv; q0 += 32 * kUg) {
            float4 x[kUg];
#pragma unroll
            for (int u = 0; u < kUg; ++u) {
                const int q = q0 + 32 * u;
                if (q < nv) {
                    asm volatile(CLDET_LD_QUAL ".v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(x[u].x), "=f"(x[u].y), "=f"(x[u].z), "=f"(x[u].w)
                                 : "l"(sp + peel + 4 * q));
                }
            }
#pragma unroll
            for (int u = 0; u < kUg; ++u) {
                const int q = q0 + 32 * u;
                if (q < nv) {
                    const int pos = peel + 4 * q;                        // first of the vector's four positions
                    const uint32_t plain4 = __funnelshift_r(plain[pos >> 5], plain[(pos >> 5) + 1], pos & 31) & 0xFu;
                    const float4 g = head_vec<GAMMA2, VARIANTS, GRAD, LOGITS>(x[u], c, plain4, smeta + pos,
                                                                             abs0 + (int64_t)pos * kHeadTypes, a, sc, as_bg, need_iou, acc);
                    if (GRAD) {
                        asm volatile(CLDET_ST_QUAL ".v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dp + pos), "f"(g.x), "f"(g.y), "f"(g.z), "f"(g.w)
                                     : "memory");
                    }
                }
            }
        }
    }
}

// The fused forward+backward launch (GRAD) carries the bulk-copy sweep: kHeadStageBytes of dynamic shared memory per block,
// CLDET_HEAD_TMA_MINBLOCKS blocks per SM.  The forward-only launch keeps the register sweeps and needs no dynamic memory.
template <bool GAMMA2, bool VARIANTS, bool GRAD, bool LOGITS>
__global__ void __launch_bounds__(kLossThreads, GRAD ? CLDET_HEAD_TMA_MINBLOCKS : CLDET_HEAD_MINBLOCKS)
    focal_loss_head_kernel(const LossArgs a, const HeadLevels lv) {
    extern __shared__ __align__(128) float head_stage[];
    __shared__ __align__(8) unsigned long long bars[kHeadWarps * kHeadStages];
    __shared__ float red[4][kLossThreads / 32];
    __shared__ double fin[4][kLossThreads / 32];
    __shared__ bool is_last;
    __shared__ uint32_t smeta[kHeadPos];
    __shared__ uint32_t plain[kHeadMaskWords + 1];      // + 1: the funnel shift reads one word past the last

    const int j = blockIdx.y;
    const int npos = a.npos[j];
    const ImageScales sc = image_scales(a, j, npos);
    Acc acc = acc_zero();
    head_chunk<GAMMA2, VARIANTS, GRAD, LOGITS, GRAD>(a, lv, j, (int)blockIdx.x, 0, sc, acc, smeta, plain, head_stage, bars);
    finish_block(a, j, (int)blockIdx.x, acc, npos, sc, red, fin, &is_last);
}

// Backward with upstream weights that differ from the baked ones: the image's gradients are recomputed (one mode only; the
// positives-only patch of the concatenated layout is not worth a second code path here).
template <bool GAMMA2, bool VARIANTS, bool LOGITS>
__global__ void __launch_bounds__(kLossThreads, 4) focal_head_reweight_kernel(const LossArgs a, const HeadLevels lv) {
    __shared__ uint32_t smeta[kHeadPos];
    __shared__ uint32_t plain[kHeadMaskWords + 1];      // + 1: the funnel shift reads one word past the last
    const int j = blockIdx.y;
    const float* wo = a.baked_weights + j;
    const int N = a.N;
    const float wn[4] = {weight_of(a, 0, j), weight_of(a, 1, j), weight_of(a, 2, j), weight_of(a, 3, j)};
    const bool enh_on = a.p.incremental && a.p.enhance_on_new;
    const bool changed = (wn[0] != wo[0]) || (wn[1] != wo[N]) || (wn[2] != wo[2 * N]) || (enh_on && wn[3] != wo[3 * N]);
    if (!changed) return;
    const ImageScales sc = image_scales(a, j, a.npos[j]);
    for (int chunk = blockIdx.x; chunk < a.bpi; chunk += gridDim.x) {
        Acc acc = acc_zero();
        head_chunk<GAMMA2, VARIANTS, true, LOGITS>(a, lv, j, chunk, 1, sc, acc, smeta, plain);
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int done = atomicAdd(&a.rw_counters[j], 1u);
        if (done == gridDim.x - 1u) {
            float* wb = a.baked_weights + j;
#pragma unroll
            for (int k = 0; k < 4; ++k) wb[k * N] = wn[k];
            a.rw_counters[j] = 0;
        }
    }
}

// new_ignore_past_class pre-pass (losses.py:326-327) on the head layout: one thread per anchor, positions fastest so that the
// loads of one class column are coalesced.
template <bool LOGITS>
__global__ void __launch_bounds__(256) old_class_flag_head_kernel(const HeadLevels lv, int level, int64_t A, int C, int past,
                                                                  uint32_t* __restrict__ meta) {
    const int j = blockIdx.y;
    const int hw = lv.hw[level];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= hw * kHeadTypes) return;
    const int k = t / hw, pp = t - k * hw;
    const int64_t an = lv.anchor_off[level] + (int64_t)pp * kHeadTypes + k;
    uint32_t m = meta[(int64_t)j * A + an] & ~CLDET_META_OLD_ACTIVE;
    if (meta_state(m) == CLDET_STATE_BG) {
        const float* base = lv.cls[level] + ((int64_t)j * (kHeadTypes * C) + (int64_t)k * C) * hw + pp;
        float s = 0.0f;
        for (int c = 0; c < past; ++c) {
            const float x = base[(int64_t)c * hw];
            s = __fadd_rn(s, fminf(fmaxf(LOGITS ? sigmoid_exact(x) : x, 1e-4f), 0.9999f));
        }
        if (s < 0.5f) m |= CLDET_META_OLD_ACTIVE;
    }
    meta[(int64_t)j * A + an] = m;
}

template <bool LOGITS>
void run_head_loss_kernels(const LossArgs& a, const HeadLevels& lv, bool grad, bool gamma2, bool variants, dim3 grid, cudaStream_t s) {
#define CLDET_HEAD_LAUNCH(G2, VAR)                                                                              \
    do {                                                                                                        \
        if (grad) {                                                                                             \
            static const cudaError_t attr = cudaFuncSetAttribute(focal_loss_head_kernel<G2, VAR, true, LOGITS>, \
                                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                                                 (int)kHeadStageBytes);                         \
            (void)attr;                                                                                         \
            focal_loss_head_kernel<G2, VAR, true, LOGITS><<<grid, kLossThreads, kHeadStageBytes, s>>>(a, lv);   \
        } else {                                                                                                \
            focal_loss_head_kernel<G2, VAR, false, LOGITS><<<grid, kLossThreads, 0, s>>>(a, lv);                \
        }                                                                                                       \
    } while (0)
    if (gamma2) {
        if (variants) CLDET_HEAD_LAUNCH(true, true);
        else CLDET_HEAD_LAUNCH(true, false);
    } else {
        if (variants) CLDET_HEAD_LAUNCH(false, true);
        else CLDET_HEAD_LAUNCH(false, false);
    }
#undef CLDET_HEAD_LAUNCH
}

template <bool LOGITS>
void run_head_reweight_kernels(const LossArgs& a, const HeadLevels& lv, bool gamma2, bool variants, dim3 grid, cudaStream_t s) {
    if (gamma2) {
        if (variants) focal_head_reweight_kernel<true, true, LOGITS><<<grid, kLossThreads, 0, s>>>(a, lv);
        else focal_head_reweight_kernel<true, false, LOGITS><<<grid, kLossThreads, 0, s>>>(a, lv);
    } else {
        if (variants) focal_head_reweight_kernel<false, true, LOGITS><<<grid, kLossThreads, 0, s>>>(a, lv);
        else focal_head_reweight_kernel<false, false, LOGITS><<<grid, kLossThreads, 0, s>>>(a, lv);
    }
}

template <bool LOGITS>
void run_head_flag_kernels(const HeadLevels& lv, int N, int64_t A, int C, int past, uint32_t* meta, cudaStream_t s) {
    for (int l = 0; l < lv.n; ++l) {
        dim3 g((unsigned)((lv.hw[l] * kHeadTypes + 255) / 256), (unsigned)N);
        old_class_flag_head_kernel<LOGITS><<<g, 256, 0, s>>>(lv, l, A, C, past, meta);
    }
}

extern template void run_head_loss_kernels<true>(const LossArgs&, const HeadLevels&, bool, bool, bool, dim3, cudaStream_t);
extern template void run_head_loss_kernels<false>(const LossArgs&, const HeadLevels&, bool, bool, bool, dim3, cudaStream_t);
extern template void run_head_reweight_kernels<true>(const LossArgs&, const HeadLevels&, bool, bool, dim3, cudaStream_t);
extern template void run_head_reweight_kernels<false>(const LossArgs&, const HeadLevels&, bool, bool, dim3, cudaStream_t);
extern template void run_head_flag_kernels<true>(const HeadLevels&, int, int64_t, int, int, uint32_t*, cudaStream_t);
extern template void run_head_flag_kernels<false>(const HeadLevels&, int, int64_t, int, int, uint32_t*, cudaStream_t);

}  // namespace cldet
