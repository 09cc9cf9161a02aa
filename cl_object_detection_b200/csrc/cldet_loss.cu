// K3 host side: C-ABI entry points of the fused focal-loss + smooth-L1 forward/backward (kernels: cldet_loss_kernels.cuh).
// This translation unit instantiates the probabilities-in kernels; cldet_loss_logits.cu the logits-in (sigmoid-fused) ones.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "cldet_loss_head_kernels.cuh"

namespace cldet {

template void run_loss_kernels<false>(const LossArgs&, int, bool, bool, bool, dim3, cudaStream_t);
template void run_reweight_kernels<false>(const LossArgs&, int, bool, bool, dim3, cudaStream_t);
template void run_head_loss_kernels<false>(const LossArgs&, const HeadLevels&, bool, bool, bool, dim3, cudaStream_t);
template void run_head_reweight_kernels<false>(const LossArgs&, const HeadLevels&, bool, bool, dim3, cudaStream_t);
template void run_head_flag_kernels<false>(const HeadLevels&, int, int64_t, int, int, uint32_t*, cudaStream_t);

// new_ignore_past_class pre-pass (losses.py:326-327): flag background anchors whose clamped old-class
// probabilities sum to < 0.5.  fp32 sequential sum in column order.
__global__ void __launch_bounds__(256) old_class_flag_kernel(const float* __restrict__ cls, int64_t A, int C, int past,
                                                             int is_logits, uint32_t* __restrict__ meta) {
    const int j = blockIdx.y;
    const int64_t an = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (an >= A) return;
    uint32_t m = meta[(int64_t)j * A + an] & ~CLDET_META_OLD_ACTIVE;
    if (meta_state(m) == CLDET_STATE_BG) {
        const float* row = cls + ((int64_t)j * A + an) * C;
        float s = 0.0f;
        for (int c = 0; c < past; ++c) s = __fadd_rn(s, fminf(fmaxf(is_logits ? sigmoid_exact(row[c]) : row[c], 1e-4f), 0.9999f));
        if (s < 0.5f) m |= CLDET_META_OLD_ACTIVE;
    }
    meta[(int64_t)j * A + an] = m;
}

struct LossPlan {
    int anchors_per_block;
    int bpi;
    uint32_t div_magic;
};

static int64_t gcd64(int64_t x, int64_t y) { return y == 0 ? x : gcd64(y, x % y); }

static LossPlan make_plan(int N, int64_t A, int C, int vec) {
    LossPlan pl;
    // ~24k elements per block; with a vector path the chunk is made a whole number of fully unrolled tiles
    // (anchors_per_block * C/vec divisible by kTile) so the hot loop runs without bounds checks.
    int64_t quantum = 32;
    const int64_t kTile = tile_for(vec);
    if (vec > 1) quantum = kTile / gcd64(C / vec, kTile);
    int64_t target = 24576;                                  // elements per block
    if (const char* e = getenv("CLDET_LOSS_BLOCK_ELEMS")) target = std::max<int64_t>(1024, atoll(e));    // A/B experiments only
    int64_t apb = (target / C + quantum - 1) / quantum * quantum;
    if (apb > kMaxAnchorsPerBlock) apb = kMaxAnchorsPerBlock / quantum * quantum;
    if (apb < 32) apb = (quantum <= kMaxAnchorsPerBlock) ? quantum : 32;
    // small problems: keep at least ~4 blocks per SM in flight (ragged tiles are fine there)
    const int64_t want_blocks = (int64_t)sm_count() * 4;
    int64_t cap = ((int64_t)N * A / want_blocks + 31) / 32 * 32;
    if (cap < 32) cap = 32;
    if (apb > cap) apb = cap;
    if (apb > kMaxAnchorsPerBlock) apb = kMaxAnchorsPerBlock;
    pl.anchors_per_block = (int)apb;
    pl.bpi = (int)((A + apb - 1) / apb);
    const uint64_t span = (uint64_t)apb * C;              // largest dividend + 1
    const uint64_t magic = (1ull << 32) / (uint64_t)C + 1;
    // __umulhi(x, magic) == x / C for all x < span iff span * (magic*C - 2^32) < 2^32
    pl.div_magic = (span * (magic * (uint64_t)C - (1ull << 32)) < (1ull << 32) && magic < (1ull << 32)) ? (uint32_t)magic : 0u;
    return pl;
}

struct Workspace {
    float* partials;
    unsigned int* counters;
};

// workspace: [counters N | npos accumulator N | reweight counters N | images-done counter 1 | pad to 256 B | partials]; the
// header must be zero before the first call and is left zero by every call.
static size_t workspace_header_bytes(int N) { return (((size_t)N * 3 + 1) * sizeof(unsigned int) + 255) / 256 * 256; }

static size_t workspace_partials_bytes(int N, int64_t A) {
    // worst case bpi: 32 anchors per block (concatenated layout); the head layout has 9 * ceil(hw_l / kHeadPos) blocks per
    // level, i.e. at most A / kHeadPos + 9 per level
    const int64_t max_bpi = (A + 31) / 32 + kHeadTypes * kHeadMaxLevels;
    return ((size_t)N * max_bpi * 4 * sizeof(float) + 255) / 256 * 256;
}

// [header | partials | best: N*A uint64 (IoU_max bits, ~row) keys of the GT-centric assignment (zero between calls)]
static size_t workspace_best_bytes(int N, int64_t A) { return ((size_t)N * A * sizeof(unsigned long long) + 255) / 256 * 256; }
static size_t workspace_touched_bytes(int N, int64_t A) { return ((size_t)N * ((A + 31) / 32) * sizeof(uint32_t) + 255) / 256 * 256; }

// ... | touched: N * ceil(A/32) words, one bit per anchor that holds a key (zero between calls)]
static size_t workspace_bytes(int N, int64_t A) {
    return workspace_header_bytes(N) + workspace_partials_bytes(N, A) + workspace_best_bytes(N, A) + workspace_touched_bytes(N, A) + 256;
}

// widest vector the class map allows: rows must be a whole number of vectors and the buffers aligned to the vector
static int pick_vec(int C, const void* cls, const void* gcls) {
    const uintptr_t bits = (uintptr_t)cls | (uintptr_t)gcls;
    if (C % 8 == 0 && (bits & 31) == 0) return 8;
    if (C % 4 == 0 && (bits & 15) == 0) return 4;
    return 1;
}

static bool has_variants(const cldet_loss_params& p) {
    return p.incremental && (p.ignore_past_class || p.decrease_positive_by_iou || p.enhance_on_new ||
                             p.decrease_positive != 1.0f);
}

static int check_common(const float* d_cls, const float* d_anchors, const float* d_annotations, int N, int64_t A, int C, int G,
                        const cldet_loss_params* params) {
    if (!d_cls || !d_anchors || !d_annotations || !params) return CLDET_ERR_INVALID_ARGUMENT;
    if (N <= 0 || N > 65535 || A <= 0 || C <= 0 || C > CLDET_MAX_CLASSES || G <= 0 || G > CLDET_MAX_GT_ROWS)
        return CLDET_ERR_INVALID_ARGUMENT;
    if (A * (int64_t)C >= (1ll << 40)) return CLDET_ERR_INVALID_ARGUMENT;
    if (params->incremental && (params->past_class_num < 0 || params->past_class_num > C)) return CLDET_ERR_INVALID_ARGUMENT;
    return CLDET_OK;
}

// optional per-thread profiling hook: events recorded around the two stages of the NEXT fused call of this thread
static thread_local cudaEvent_t g_prof_events[3] = {nullptr, nullptr, nullptr};

}  // namespace cldet

using namespace cldet;

extern "C" {

int cldet_focal_loss_profile_events(void* ev_begin, void* ev_between, void* ev_end) {
    g_prof_events[0] = (cudaEvent_t)ev_begin;
    g_prof_events[1] = (cudaEvent_t)ev_between;
    g_prof_events[2] = (cudaEvent_t)ev_end;
    return CLDET_OK;
}

size_t cldet_focal_loss_workspace_bytes(int num_images, int64_t num_anchors) {
    if (num_images <= 0 || num_anchors <= 0) return 0;
    return workspace_bytes(num_images, num_anchors);
}

static int loss_stage(const float* d_cls, const float* d_reg, const float* d_anchors, const float* d_annotations,
                      int num_images, int64_t num_anchors, int num_classes, int gt_rows, const cldet_loss_params* params,
                      const float* d_weights, float* d_baked_weights, float* d_grad_cls, float* d_grad_reg, float* d_losses,
                      uint32_t* d_meta, const float* d_iou_max, const int32_t* d_npos, int32_t* d_npos_out, int32_t* d_npos_reset,
                      uint8_t* d_bg_mask, int32_t* d_status, void* d_workspace, size_t ws_bytes, void* stream,
                      const cldet_peer_exchange* peer = nullptr, unsigned long long* d_best = nullptr, const int32_t* d_nvalid = nullptr,
                      float* d_iou_out = nullptr, uint32_t* d_touched = nullptr, float* d_reg_mean = nullptr) {
    int rc = check_common(d_cls, d_anchors, d_annotations, num_images, num_anchors, num_classes, gt_rows, params);
    if (rc) return rc;
    if (!d_reg || !d_losses || !d_meta || !d_npos || !d_workspace) return CLDET_ERR_INVALID_ARGUMENT;
    const bool grad = d_weights != nullptr;
    if (grad != (d_grad_cls != nullptr) || grad != (d_grad_reg != nullptr)) return CLDET_ERR_INVALID_ARGUMENT;
    if (params->incremental && params->decrease_positive_by_iou && !d_iou_max) return CLDET_ERR_INVALID_ARGUMENT;
    if (ws_bytes < workspace_bytes(num_images, num_anchors)) return CLDET_ERR_WORKSPACE_TOO_SMALL;
    if (((uintptr_t)d_cls | (uintptr_t)d_grad_cls | (uintptr_t)d_reg | (uintptr_t)d_grad_reg | (uintptr_t)d_anchors) & 15)
        return CLDET_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;

    const int vec = pick_vec(num_classes, d_cls, d_grad_cls);
    const LossPlan pl = make_plan(num_images, num_anchors, num_classes, vec);
    LossArgs a;
    a.cls = d_cls; a.reg = d_reg; a.anchors = reinterpret_cast<const float4*>(d_anchors); a.ann = d_annotations;
    a.N = num_images; a.A = num_anchors; a.C = num_classes; a.G = gt_rows; a.p = *params;
    for (int k = 0; k < 4; ++k) {
        a.w_ptr[k] = d_weights ? d_weights + (size_t)k * num_images : nullptr;
        a.w_stride[k] = 1;
    }
    a.has_w = d_weights ? 1 : 0;
    a.w_reg_mean = nullptr; a.reg_mean_scale = 0.0f;
    a.baked_weights = d_baked_weights; a.gcls = d_grad_cls; a.greg = d_grad_reg; a.losses = d_losses;
    a.reg_mean = d_reg_mean; a.images_done = reinterpret_cast<unsigned int*>(d_workspace) + 3 * (size_t)num_images;
    a.meta = d_meta; a.iou_max = d_iou_max; a.npos = d_npos; a.bg_mask = d_bg_mask; a.status = d_status;
    a.npos_out = d_npos_out; a.npos_reset = d_npos_reset; a.rw_counters = nullptr;
    a.best = d_best; a.touched = d_touched; a.meta_out = d_meta; a.iou_out = d_iou_out; a.nvalid = d_nvalid;
    a.peer_terms = nullptr; a.peer_flags = nullptr; a.rank = 0; a.world = 1; a.parity = 0;
    a.wait_flags = nullptr; a.wait_terms = nullptr; a.wait_out = nullptr; a.wait_status = nullptr; a.wait_target = 0;
    a.wait_timeout_ns = 0;
    if (peer && peer->world > 1) {
        if (!peer->d_peer_terms || !peer->d_peer_flags || peer->rank < 0 || peer->rank >= peer->world || peer->world > 64 ||
            (peer->parity != 0 && peer->parity != 1))
            return CLDET_ERR_INVALID_ARGUMENT;
        a.peer_terms = reinterpret_cast<float* const*>(peer->d_peer_terms);
        a.peer_flags = reinterpret_cast<unsigned int* const*>(peer->d_peer_flags);
        a.rank = peer->rank; a.world = peer->world; a.parity = peer->parity;
        if (peer->d_wait_out) {
            if (!peer->d_flags_local || !peer->d_terms_local || peer->timeout_ms <= 0 || kLossThreads < peer->world)
                return CLDET_ERR_INVALID_ARGUMENT;
            a.wait_flags = reinterpret_cast<const unsigned int*>(peer->d_flags_local);
            a.wait_terms = reinterpret_cast<const float*>(peer->d_terms_local) + (size_t)peer->parity * peer->world * 4 * num_images;
            a.wait_out = peer->d_wait_out;
            a.wait_status = peer->d_wait_status;
            a.wait_target = peer->target_arrivals;
            a.wait_timeout_ns = (unsigned long long)peer->timeout_ms * 1000000ull;
        }
    }
    a.counters = reinterpret_cast<unsigned int*>(d_workspace);
    a.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(d_workspace) + workspace_header_bytes(num_images));
    a.anchors_per_block = pl.anchors_per_block; a.bpi = pl.bpi; a.div_magic = pl.div_magic;

    const bool variants = has_variants(*params);
    if (params->incremental && params->ignore_past_class && params->new_ignore_past_class && params->past_class_num > 0) {
        dim3 g((unsigned)((num_anchors + 255) / 256), (unsigned)num_images);
        old_class_flag_kernel<<<g, 256, 0, s>>>(d_cls, num_anchors, num_classes, params->past_class_num, params->cls_is_logits, d_meta);
        CLDET_LAUNCH_CHECK();
    }
    dim3 grid((unsigned)pl.bpi, (unsigned)num_images);
    const bool gamma2 = params->gamma == 2.0f;
    if (params->cls_is_logits) run_loss_kernels<true>(a, vec, grad, gamma2, variants, grid, s);
    else run_loss_kernels<false>(a, vec, grad, gamma2, variants, grid, s);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

int cldet_focal_loss_from_assignment(const float* d_cls, const float* d_reg, const float* d_anchors,
                                     const float* d_annotations, int num_images, int64_t num_anchors, int num_classes,
                                     int gt_rows, const cldet_loss_params* params, const float* d_weights,
                                     float* d_baked_weights, float* d_grad_cls, float* d_grad_reg, float* d_losses,
                                     uint32_t* d_meta, const float* d_iou_max, const int32_t* d_npos, uint8_t* d_bg_mask,
                                     int32_t* d_status, void* d_workspace, size_t ws_bytes, void* stream) {
    return loss_stage(d_cls, d_reg, d_anchors, d_annotations, num_images, num_anchors, num_classes, gt_rows, params, d_weights,
                      d_baked_weights, d_grad_cls, d_grad_reg, d_losses, d_meta, d_iou_max, d_npos, nullptr, nullptr, d_bg_mask, d_status,
                      d_workspace, ws_bytes, stream);
}

int cldet_focal_loss_sharded(const float* d_cls, const float* d_reg, const float* d_anchors, const float* d_annotations,
                             int num_images, int64_t num_anchors, int num_classes, int gt_rows,
                             const cldet_loss_params* params, const float* d_weights, float* d_baked_weights,
                             float* d_grad_cls, float* d_grad_reg, float* d_losses,
                             uint32_t* d_meta, float* d_iou_max, int32_t* d_npos, int32_t* d_nvalid,
                             uint8_t* d_bg_mask, int32_t* d_status,
                             void* d_workspace, size_t ws_bytes, const cldet_peer_exchange* peer, float* d_reg_mean, void* stream) {
    int rc = check_common(d_cls, d_anchors, d_annotations, num_images, num_anchors, num_classes, gt_rows, params);
    if (rc) return rc;
    if (!d_npos || !d_nvalid || !d_meta || !d_workspace) return CLDET_ERR_INVALID_ARGUMENT;
    if (ws_bytes < workspace_bytes(num_images, num_anchors)) return CLDET_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t s = (cudaStream_t)stream;
    if (d_status) CLDET_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), s));
    // positives are accumulated in the (zeroed) workspace; the loss kernel's last block per image publishes the count to
    // d_npos and clears the accumulator again, so no memset is needed between calls
    int32_t* npos_acc = reinterpret_cast<int32_t*>(d_workspace) + num_images;
    cudaEvent_t ev[3] = {g_prof_events[0], g_prof_events[1], g_prof_events[2]};
    g_prof_events[0] = g_prof_events[1] = g_prof_events[2] = nullptr;      // one-shot
    if (ev[0]) CLDET_CUDA_TRY(cudaEventRecord(ev[0], s));
    // Standard anchor grid of a known image size: GT-centric assignment (a few hundred anchors per GT box instead of A x G
    // pairs); the loss kernel turns the IoU_max bits into assignment words itself.  The new_ignore_past_class pre-pass needs
    // ready-made words, and arbitrary anchor sets have no grid: both use the anchor-centric kernel.
    unsigned long long* best = nullptr;
    uint32_t* touched = nullptr;
    const bool needs_words = params->incremental && params->ignore_past_class && params->new_ignore_past_class &&
                             params->past_class_num > 0;
    if (params->image_height > 0 && params->image_width > 0 && !needs_words) {
        best = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(d_workspace) + workspace_header_bytes(num_images) +
                                           workspace_partials_bytes(num_images, num_anchors));
        // the key bitmap needs chunks that own whole bitmap words (always true except for C >= 1024, where chunks are tiny)
        const LossPlan pl = make_plan(num_images, num_anchors, num_classes, pick_vec(num_classes, d_cls, d_grad_cls));
        if (pl.anchors_per_block % 32 == 0)
            touched = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(best) + workspace_best_bytes(num_images, num_anchors));
        rc = launch_gt_scatter(params->image_height, params->image_width, d_anchors, num_anchors, d_annotations, num_images,
                               gt_rows, best, touched, npos_acc, d_nvalid, s);
        if (rc == CLDET_ERR_UNSUPPORTED) best = nullptr, touched = nullptr;      // anchors are not this image size's grid
        else if (rc) return rc;
    }
    if (!best) {
        rc = cldet_iou_assign(d_anchors, num_anchors, d_annotations, num_images, gt_rows, num_classes, d_meta, nullptr,
                              d_iou_max, npos_acc, d_nvalid, stream);
        if (rc) return rc;
    }
    if (ev[1]) CLDET_CUDA_TRY(cudaEventRecord(ev[1], s));
    rc = loss_stage(d_cls, d_reg, d_anchors, d_annotations, num_images, num_anchors, num_classes, gt_rows, params, d_weights,
                      d_baked_weights, d_grad_cls, d_grad_reg, d_losses, d_meta, d_iou_max, npos_acc, d_npos, npos_acc, d_bg_mask, d_status,
                      d_workspace, ws_bytes, stream, peer, best, d_nvalid, best ? d_iou_max : nullptr, touched, d_reg_mean);
    if (rc) return rc;
    if (ev[2]) CLDET_CUDA_TRY(cudaEventRecord(ev[2], s));
    return CLDET_OK;
}

int cldet_focal_loss(const float* d_cls, const float* d_reg, const float* d_anchors, const float* d_annotations,
                     int num_images, int64_t num_anchors, int num_classes, int gt_rows,
                     const cldet_loss_params* params, const float* d_weights, float* d_baked_weights,
                     float* d_grad_cls, float* d_grad_reg, float* d_losses,
                     uint32_t* d_meta, float* d_iou_max, int32_t* d_npos, int32_t* d_nvalid,
                     uint8_t* d_bg_mask, int32_t* d_status,
                     void* d_workspace, size_t ws_bytes, void* stream) {
    return cldet_focal_loss_sharded(d_cls, d_reg, d_anchors, d_annotations, num_images, num_anchors, num_classes, gt_rows, params,
                                    d_weights, d_baked_weights, d_grad_cls, d_grad_reg, d_losses, d_meta, d_iou_max, d_npos,
                                    d_nvalid, d_bg_mask, d_status, d_workspace, ws_bytes, nullptr, nullptr, stream);
}

// Consumer side of the fused all-gather.  Thread r < world waits (acquire, system scope) until source rank r's arrival
// counter for this parity has reached `target` -- the counters only ever grow, `target` = N x (number of uses of the parity
// so far), compared wrap-safely -- then the whole block copies this parity's [world][4][N] terms into the caller's PRIVATE
// [4][world*N] tensor (global image order), so nothing the caller keeps aliases memory that peers overwrite two steps later.
// The wait is bounded by `timeout_ns` of %globaltimer: a peer that never arrives poisons ITS rows of the output with NaN and
// raises *status (which may live in mapped host memory, so the host sees it without a synchronisation) instead of hanging the
// GPU; because nothing is ever reset, a late arrival cannot desynchronise the following steps.
__global__ void __launch_bounds__(256)
peer_wait_copy_kernel(const unsigned int* flags, int world, int parity, unsigned int target, unsigned long long timeout_ns,
                      const float* terms, int n, float* out, float* reg_mean, int32_t* status) {
    __shared__ int bad[64];
    peer_wait_copy_block(flags, world, parity, target, timeout_ns, terms, n, out, reg_mean, status, bad);
}

// Buffers every rank of a node can map: allocated with cudaMalloc (whole allocation = one IPC handle, offset 0), exported
// with cudaIpcGetMemHandle and opened by the peers WITH THEIR OWN DEVICE CURRENT (cudaIpcMemLazyEnablePeerAccess), which is
// what makes the mapping dereferenceable from kernels of the opening device.
int cldet_peer_alloc(size_t bytes, void** d_ptr, unsigned char* h_handle64) {
    if (!d_ptr || !h_handle64 || bytes == 0) return CLDET_ERR_INVALID_ARGUMENT;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void* p = nullptr;
    CLDET_CUDA_TRY(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        CLDET_CUDA_TRY(e);
    }
    memcpy(h_handle64, &h, 64);
    *d_ptr = p;
    return CLDET_OK;
}

int cldet_peer_open(const unsigned char* h_handle64, void** d_ptr) {
    if (!h_handle64 || !d_ptr) return CLDET_ERR_INVALID_ARGUMENT;
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle64, 64);
    CLDET_CUDA_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return CLDET_OK;
}

int cldet_peer_close(void* d_ptr) {
    if (!d_ptr) return CLDET_OK;
    CLDET_CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
    return CLDET_OK;
}

int cldet_peer_free(void* d_ptr) {
    if (!d_ptr) return CLDET_OK;
    CLDET_CUDA_TRY(cudaFree(d_ptr));
    return CLDET_OK;
}

int cldet_enable_peer_access(int peer_device) {
    int cur = 0;
    CLDET_CUDA_TRY(cudaGetDevice(&cur));
    if (peer_device == cur) return CLDET_OK;
    int can = 0;
    CLDET_CUDA_TRY(cudaDeviceCanAccessPeer(&can, cur, peer_device));
    if (!can) return CLDET_ERR_UNSUPPORTED;
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        (void)cudaGetLastError();
        return CLDET_OK;
    }
    CLDET_CUDA_TRY(e);
    return CLDET_OK;
}

int cldet_peer_wait(const void* d_flags_local, const float* d_terms_local, int world, int num_images, int parity,
                    uint32_t target_arrivals, int timeout_ms, float* d_out, float* d_reg_mean, int32_t* d_status, void* stream) {
    if (!d_flags_local || world < 1 || world > 64 || (parity != 0 && parity != 1) || num_images <= 0 || timeout_ms <= 0)
        return CLDET_ERR_INVALID_ARGUMENT;
    if ((d_out || d_reg_mean) && !d_terms_local) return CLDET_ERR_INVALID_ARGUMENT;
    const float* terms = d_terms_local ? d_terms_local + (size_t)parity * world * 4 * num_images : nullptr;
    peer_wait_copy_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned int*>(d_flags_local), world, parity,
                                                              target_arrivals, (unsigned long long)timeout_ms * 1000000ull, terms,
                                                              num_images, d_out, d_reg_mean, d_status);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

// ---- head layout: per-level NCHW conv outputs in, gradients of the same layout out ----
// Fills the level table for an (image_height, image_width) input: level l = 3..7 has ceil(H/2^l) x ceil(W/2^l) positions
// (retinanet/anchors.py:25) with 9 anchors each; channels are k*C + c (classification) and k*4 + i (regression).
static int head_levels(const float* const* h_cls, const float* const* h_reg, float* const* h_gcls, float* const* h_greg,
                       int num_levels, int image_height, int image_width, bool grad, HeadLevels* lv, int64_t* num_anchors,
                       int* chunks_per_image) {
    if (!h_cls || !h_reg || num_levels != kNumLevels || image_height <= 0 || image_width <= 0) return CLDET_ERR_INVALID_ARGUMENT;
    if (grad && (!h_gcls || !h_greg)) return CLDET_ERR_INVALID_ARGUMENT;
    lv->n = num_levels;
    int64_t aoff = 0;
    int coff = 0;
    for (int l = 0; l < kHeadMaxLevels; ++l) {
        lv->cls[l] = lv->reg[l] = nullptr;
        lv->gcls[l] = lv->greg[l] = nullptr;
        lv->hw[l] = 0;
        lv->pos_chunks[l] = 1;
    }
    for (int l = 0; l < num_levels; ++l) {
        const int sh = 3 + l;
        const int64_t hw = (int64_t)((image_height + (1 << sh) - 1) >> sh) * ((image_width + (1 << sh) - 1) >> sh);
        if (hw <= 0 || hw > (1 << 26)) return CLDET_ERR_INVALID_ARGUMENT;
        if (!h_cls[l] || !h_reg[l] || (grad && (!h_gcls[l] || !h_greg[l]))) return CLDET_ERR_INVALID_ARGUMENT;
        lv->cls[l] = h_cls[l]; lv->reg[l] = h_reg[l];
        lv->gcls[l] = grad ? h_gcls[l] : nullptr; lv->greg[l] = grad ? h_greg[l] : nullptr;
        lv->hw[l] = (int)hw;
        lv->anchor_off[l] = aoff;
        lv->chunk_off[l] = coff;
        lv->pos_chunks[l] = (int)((hw + kHeadPos - 1) / kHeadPos);
        aoff += hw * kHeadTypes;
        coff += kHeadTypes * lv->pos_chunks[l];
    }
    for (int l = num_levels; l <= kHeadMaxLevels; ++l) {
        lv->anchor_off[l] = aoff;
        lv->chunk_off[l] = coff;
    }
    *num_anchors = aoff;
    *chunks_per_image = coff;
    return CLDET_OK;
}

static void head_args(LossArgs* a, const float* d_anchors, const float* d_annotations, int N, int64_t A, int C, int G,
                      const cldet_loss_params* params, int chunks_per_image, void* d_workspace) {
    memset(a, 0, sizeof(*a));
    a->anchors = reinterpret_cast<const float4*>(d_anchors); a->ann = d_annotations;
    a->N = N; a->A = A; a->C = C; a->G = G; a->p = *params;
    a->world = 1;
    a->counters = reinterpret_cast<unsigned int*>(d_workspace);
    a->partials = reinterpret_cast<float*>(reinterpret_cast<char*>(d_workspace) + workspace_header_bytes(N));
    a->rw_counters = reinterpret_cast<unsigned int*>(d_workspace) + 2 * (size_t)N;
    a->bpi = chunks_per_image;
    a->anchors_per_block = kHeadPos;
    const uint64_t span = (uint64_t)kHeadTypes * C;              // dividends are channel indices < 9*C
    const uint64_t magic = (1ull << 32) / (uint64_t)C + 1;
    a->div_magic = (span * (magic * (uint64_t)C - (1ull << 32)) < (1ull << 32) && magic < (1ull << 32)) ? (uint32_t)magic : 0u;
}

static int reweight_impl(const float* d_cls, const float* d_reg, const float* d_anchors, const float* d_annotations,
                         int num_images, int64_t num_anchors, int num_classes, int gt_rows, const cldet_loss_params* params,
                         const float* const w_ptr[4], const int64_t w_stride[4], float* d_baked_weights, float* d_grad_cls,
                         float* d_grad_reg, const uint32_t* d_meta, const float* d_iou_max, const int32_t* d_npos,
                         void* d_workspace, size_t ws_bytes, void* stream, const float* w_reg_mean = nullptr,
                         float reg_mean_scale = 0.0f) {
    int rc = check_common(d_cls, d_anchors, d_annotations, num_images, num_anchors, num_classes, gt_rows, params);
    if (rc) return rc;
    if (!d_reg || !d_baked_weights || !d_grad_cls || !d_grad_reg || !d_meta || !d_npos || !d_workspace)
        return CLDET_ERR_INVALID_ARGUMENT;
    if (ws_bytes < workspace_bytes(num_images, num_anchors)) return CLDET_ERR_WORKSPACE_TOO_SMALL;
    if (params->incremental && params->decrease_positive_by_iou && !d_iou_max) return CLDET_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    const int vec = pick_vec(num_classes, d_cls, d_grad_cls);
    const LossPlan pl = make_plan(num_images, num_anchors, num_classes, vec);
    LossArgs a;
    a.cls = d_cls; a.reg = d_reg; a.anchors = reinterpret_cast<const float4*>(d_anchors); a.ann = d_annotations;
    a.N = num_images; a.A = num_anchors; a.C = num_classes; a.G = gt_rows; a.p = *params;
    for (int k = 0; k < 4; ++k) {
        a.w_ptr[k] = w_ptr[k];
        a.w_stride[k] = (int)w_stride[k];
    }
    a.has_w = 1;
    a.w_reg_mean = w_reg_mean; a.reg_mean_scale = reg_mean_scale;
    a.reg_mean = nullptr; a.images_done = nullptr;
    a.baked_weights = d_baked_weights; a.gcls = d_grad_cls; a.greg = d_grad_reg;
    a.losses = nullptr; a.meta = d_meta; a.iou_max = d_iou_max; a.npos = d_npos; a.bg_mask = nullptr; a.status = nullptr;
    a.npos_out = nullptr; a.npos_reset = nullptr;
    a.best = nullptr; a.touched = nullptr; a.meta_out = nullptr; a.iou_out = nullptr; a.nvalid = nullptr;
    a.peer_terms = nullptr; a.peer_flags = nullptr; a.rank = 0; a.world = 1; a.parity = 0;
    a.wait_flags = nullptr; a.wait_terms = nullptr; a.wait_out = nullptr; a.wait_status = nullptr; a.wait_target = 0;
    a.wait_timeout_ns = 0;
    a.counters = nullptr; a.partials = nullptr;
    a.rw_counters = reinterpret_cast<unsigned int*>(d_workspace) + 2 * (size_t)num_images;
    a.anchors_per_block = pl.anchors_per_block; a.bpi = pl.bpi; a.div_magic = pl.div_magic;
    // ~4 CTAs per SM over the whole batch; blocks loop over their image's chunks
    const int per_image = std::max(1, (sm_count() * 4 + num_images - 1) / num_images);
    dim3 grid((unsigned)std::min(pl.bpi, per_image), (unsigned)num_images);
    const bool gamma2 = params->gamma == 2.0f;
    const bool variants = has_variants(*params);
    if (params->cls_is_logits) run_reweight_kernels<true>(a, vec, gamma2, variants, grid, s);
    else run_reweight_kernels<false>(a, vec, gamma2, variants, grid, s);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

int cldet_focal_loss_reweight(const float* d_cls, const float* d_reg, const float* d_anchors, const float* d_annotations,
                              int num_images, int64_t num_anchors, int num_classes, int gt_rows,
                              const cldet_loss_params* params, const float* d_new_weights, float* d_baked_weights,
                              float* d_grad_cls, float* d_grad_reg, const uint32_t* d_meta, const float* d_iou_max,
                              const int32_t* d_npos, void* d_workspace, size_t ws_bytes, void* stream) {
    if (!d_new_weights || num_images <= 0) return CLDET_ERR_INVALID_ARGUMENT;
    const float* rows[4] = {d_new_weights, d_new_weights + num_images, d_new_weights + 2 * (size_t)num_images,
                            d_new_weights + 3 * (size_t)num_images};
    const int64_t strides[4] = {1, 1, 1, 1};
    return reweight_impl(d_cls, d_reg, d_anchors, d_annotations, num_images, num_anchors, num_classes, gt_rows, params, rows,
                         strides, d_baked_weights, d_grad_cls, d_grad_reg, d_meta, d_iou_max, d_npos, d_workspace, ws_bytes,
                         stream);
}

int cldet_focal_loss_reweight_rows(const float* d_cls, const float* d_reg, const float* d_anchors, const float* d_annotations,
                                   int num_images, int64_t num_anchors, int num_classes, int gt_rows,
                                   const cldet_loss_params* params, const float* d_w_bg, int64_t stride_bg,
                                   const float* d_w_fg, int64_t stride_fg, const float* d_w_reg, int64_t stride_reg,
                                   const float* d_w_enh, int64_t stride_enh, const float* d_w_reg_mean, float reg_mean_scale,
                                   float* d_baked_weights, float* d_grad_cls,
                                   float* d_grad_reg, const uint32_t* d_meta, const float* d_iou_max, const int32_t* d_npos,
                                   void* d_workspace, size_t ws_bytes, void* stream) {
    const float* rows[4] = {d_w_bg, d_w_fg, d_w_reg, d_w_enh};
    const int64_t strides[4] = {stride_bg, stride_fg, stride_reg, stride_enh};
    for (int k = 0; k < 4; ++k)
        if (strides[k] < 0 || strides[k] > 0x7fffffff) return CLDET_ERR_INVALID_ARGUMENT;
    return reweight_impl(d_cls, d_reg, d_anchors, d_annotations, num_images, num_anchors, num_classes, gt_rows, params, rows,
                         strides, d_baked_weights, d_grad_cls, d_grad_reg, d_meta, d_iou_max, d_npos, d_workspace, ws_bytes,
                         stream, d_w_reg_mean, reg_mean_scale);
}

int cldet_focal_loss_head(const float* const* h_cls_levels, const float* const* h_reg_levels, int num_levels, int image_height,
                          int image_width, const float* d_anchors, const float* d_annotations, int num_images, int num_classes,
                          int gt_rows, const cldet_loss_params* params, const float* d_weights, float* d_baked_weights,
                          float* const* h_grad_cls_levels, float* const* h_grad_reg_levels, float* d_losses, uint32_t* d_meta,
                          float* d_iou_max, int32_t* d_npos, int32_t* d_nvalid, uint8_t* d_bg_mask, int32_t* d_status,
                          void* d_workspace, size_t ws_bytes, void* stream) {
    const bool grad = d_weights != nullptr;
    HeadLevels lv;
    int64_t A = 0;
    int cpi = 0;
    int rc = head_levels(h_cls_levels, h_reg_levels, h_grad_cls_levels, h_grad_reg_levels, num_levels, image_height, image_width,
                         grad, &lv, &A, &cpi);
    if (rc) return rc;
    rc = check_common(lv.cls[0], d_anchors, d_annotations, num_images, A, num_classes, gt_rows, params);
    if (rc) return rc;
    if (!d_losses || !d_meta || !d_npos || !d_nvalid || !d_workspace) return CLDET_ERR_INVALID_ARGUMENT;
    if (params->incremental && params->decrease_positive_by_iou && !d_iou_max) return CLDET_ERR_INVALID_ARGUMENT;
    if (ws_bytes < workspace_bytes(num_images, A)) return CLDET_ERR_WORKSPACE_TOO_SMALL;
    if ((uintptr_t)d_anchors & 15) return CLDET_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    if (d_status) CLDET_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), s));
    int32_t* npos_acc = reinterpret_cast<int32_t*>(d_workspace) + num_images;
    // assignment: GT-centric keys (the anchors must be the standard grid of this image size), except when the
    // new_ignore_past_class pre-pass needs ready-made words
    const bool needs_words = params->incremental && params->ignore_past_class && params->new_ignore_past_class &&
                             params->past_class_num > 0;
    unsigned long long* best = nullptr;
    if (!needs_words) {
        best = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(d_workspace) + workspace_header_bytes(num_images) +
                                                      workspace_partials_bytes(num_images, A));
        rc = launch_gt_scatter(image_height, image_width, d_anchors, A, d_annotations, num_images, gt_rows, best, nullptr, npos_acc,
                               d_nvalid, s);
        if (rc) return rc;
    } else {
        rc = cldet_iou_assign(d_anchors, A, d_annotations, num_images, gt_rows, num_classes, d_meta, nullptr, d_iou_max, npos_acc,
                              d_nvalid, stream);
        if (rc) return rc;
        if (params->cls_is_logits) run_head_flag_kernels<true>(lv, num_images, A, num_classes, params->past_class_num, d_meta, s);
        else run_head_flag_kernels<false>(lv, num_images, A, num_classes, params->past_class_num, d_meta, s);
        CLDET_LAUNCH_CHECK();
    }
    LossArgs a;
    head_args(&a, d_anchors, d_annotations, num_images, A, num_classes, gt_rows, params, cpi, d_workspace);
    for (int k = 0; k < 4; ++k) {
        a.w_ptr[k] = d_weights ? d_weights + (size_t)k * num_images : nullptr;
        a.w_stride[k] = 1;
    }
    a.has_w = grad ? 1 : 0;
    a.baked_weights = d_baked_weights; a.losses = d_losses; a.meta = d_meta; a.iou_max = d_iou_max;
    a.npos = npos_acc; a.npos_out = d_npos; a.npos_reset = npos_acc; a.bg_mask = d_bg_mask; a.status = d_status;
    a.best = best; a.touched = nullptr; a.meta_out = d_meta; a.iou_out = best ? d_iou_max : nullptr; a.nvalid = d_nvalid;
    dim3 grid((unsigned)cpi, (unsigned)num_images);
    const bool gamma2 = params->gamma == 2.0f;
    const bool variants = has_variants(*params);
    if (params->cls_is_logits) run_head_loss_kernels<true>(a, lv, grad, gamma2, variants, grid, s);
    else run_head_loss_kernels<false>(a, lv, grad, gamma2, variants, grid, s);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

int cldet_focal_loss_head_reweight(const float* const* h_cls_levels, const float* const* h_reg_levels, int num_levels,
                                   int image_height, int image_width, const float* d_anchors, const float* d_annotations,
                                   int num_images, int num_classes, int gt_rows, const cldet_loss_params* params,
                                   const float* d_w_bg, int64_t stride_bg, const float* d_w_fg, int64_t stride_fg,
                                   const float* d_w_reg, int64_t stride_reg, const float* d_w_enh, int64_t stride_enh,
                                   float* d_baked_weights, float* const* h_grad_cls_levels, float* const* h_grad_reg_levels,
                                   const uint32_t* d_meta, const float* d_iou_max, const int32_t* d_npos, void* d_workspace,
                                   size_t ws_bytes, void* stream) {
    HeadLevels lv;
    int64_t A = 0;
    int cpi = 0;
    int rc = head_levels(h_cls_levels, h_reg_levels, h_grad_cls_levels, h_grad_reg_levels, num_levels, image_height, image_width,
                         true, &lv, &A, &cpi);
    if (rc) return rc;
    rc = check_common(lv.cls[0], d_anchors, d_annotations, num_images, A, num_classes, gt_rows, params);
    if (rc) return rc;
    if (!d_baked_weights || !d_meta || !d_npos || !d_workspace) return CLDET_ERR_INVALID_ARGUMENT;
    if (params->incremental && params->decrease_positive_by_iou && !d_iou_max) return CLDET_ERR_INVALID_ARGUMENT;
    if (ws_bytes < workspace_bytes(num_images, A)) return CLDET_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t s = (cudaStream_t)stream;
    LossArgs a;
    head_args(&a, d_anchors, d_annotations, num_images, A, num_classes, gt_rows, params, cpi, d_workspace);
    const float* w[4] = {d_w_bg, d_w_fg, d_w_reg, d_w_enh};
    const int64_t st[4] = {stride_bg, stride_fg, stride_reg, stride_enh};
    for (int k = 0; k < 4; ++k) {
        a.w_ptr[k] = w[k];
        a.w_stride[k] = (int)st[k];
    }
    a.has_w = 1;
    a.baked_weights = d_baked_weights; a.meta = d_meta; a.iou_max = d_iou_max; a.npos = d_npos;
    const int per_image = std::max(1, (sm_count() * 4 + num_images - 1) / num_images);
    dim3 grid((unsigned)std::min(cpi, per_image), (unsigned)num_images);
    const bool gamma2 = params->gamma == 2.0f;
    const bool variants = has_variants(*params);
    if (params->cls_is_logits) run_head_reweight_kernels<true>(a, lv, gamma2, variants, grid, s);
    else run_head_reweight_kernels<false>(a, lv, gamma2, variants, grid, s);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

}  // extern "C"
