// Thin torch custom-op layer over the C ABI of libcldet.so (include/cldet.h): `torch.ops.cldet.*`.
//
// What lives here (and nothing else): tensors from torch's caching allocator, the per-(device, stream, N, A) workspace
// cache, the current CUDA stream, status -> exception mapping, and the autograd node of the fused loss.  Every arithmetic
// step of the path runs in the hand-written kernels behind the C ABI; this file contains no kernel and no math.
//
// Replaces the reference's Python-level boundary:
//   cldet::focal_loss  -> FocalLoss.forward + its autograd graph  (retinanet/losses.py:252-452; call sites :460, :572)
//   cldet::detect      -> the part of ResNet.predict after self.forward (retinanet/model.py:507-550) and
//                         Labeler.predict (IL_method/persuado_label.py:99-127), for a whole batch
//   cldet::batched_nms -> torchvision.ops.batched_nms / nms as called at model.py:540, persuado_label.py:116
//
// Why C++: one FocalLoss forward + backward through ctypes + a Python autograd.Function cost ~0.33 ms of host time per
// step (round-1 profiles/bench_api.jsonl) against 0.05 ms of kernels at the VOC shape.  Here a step is two C-ABI calls, a
// handful of allocator hits and one C++ autograd node.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/csrc/autograd/custom_function.h>
#include <torch/library.h>

#include <stdlib.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <unordered_map>
#include <vector>

#include "cldet.h"

namespace {

using at::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

void check_status(int rc, const char* what) {
    if (rc == CLDET_OK) return;
    std::string msg = std::string("libcldet: ") + cldet_status_string(rc);
    if (rc == CLDET_ERR_CUDA) msg += std::string(": ") + cldet_last_cuda_error();
    TORCH_CHECK(false, msg, " (", what, ")");
}

void check_f32_cuda(const Tensor& t, const char* name) {
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor: this path has no CPU implementation");
    TORCH_CHECK(t.scalar_type() == at::kFloat, name, " must be float32");
    TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}

// Zero-initialised scratch of the fused loss, cached per host thread and keyed by (device, stream, N, A): libcldet leaves a
// workspace zero-clean after every call, so it is cleared exactly once, and calls that share one are ordered by the stream.
struct WsKey {
    int device;
    void* stream;
    int64_t n, a;
    bool operator==(const WsKey& o) const { return device == o.device && stream == o.stream && n == o.n && a == o.a; }
};
struct WsHash {
    size_t operator()(const WsKey& k) const {
        return std::hash<int64_t>()(k.n * 1000003 + k.a) ^ std::hash<void*>()(k.stream) ^ (size_t)k.device * 0x9e3779b97f4a7c15ull;
    }
};
thread_local std::unordered_map<WsKey, Tensor, WsHash> g_workspaces;

Tensor loss_workspace(const Tensor& like, void* stream, int64_t n, int64_t a) {
    const WsKey key{(int)like.get_device(), stream, n, a};
    auto it = g_workspaces.find(key);
    if (it != g_workspaces.end()) return it->second;
    if (g_workspaces.size() > 8) g_workspaces.clear();
    const size_t bytes = cldet_focal_loss_workspace_bytes((int)n, a);
    Tensor ws = at::zeros({(int64_t)bytes}, like.options().dtype(at::kByte));
    g_workspaces.emplace(key, ws);
    return ws;
}

void drop_workspaces() { g_workspaces.clear(); }

cldet_loss_params make_params(double alpha, double gamma, int64_t incremental, int64_t past, int64_t ignore_past,
                              int64_t new_ignore_past, int64_t dec_by_iou, int64_t enhance, double dec_pos, int64_t height,
                              int64_t width, bool logits) {
    cldet_loss_params p;
    p.alpha = (float)alpha;
    p.gamma = (float)gamma;
    p.incremental = (int32_t)incremental;
    p.past_class_num = (int32_t)past;
    p.ignore_past_class = (int32_t)ignore_past;
    p.new_ignore_past_class = (int32_t)new_ignore_past;
    p.decrease_positive_by_iou = (int32_t)dec_by_iou;
    p.enhance_on_new = (int32_t)enhance;
    p.decrease_positive = (float)dec_pos;
    p.image_height = (int32_t)height;
    p.image_width = (int32_t)width;
    p.cls_is_logits = logits ? 1 : 0;
    return p;
}

// peer exchange of an image-sharded run, as plain integers (see dist.PeerGather):
// [0] device array of the ranks' gather buffers, [1] device array of the ranks' arrival counters, [2] rank, [3] world,
// [4] parity, [5] this rank's counters, [6] this rank's gather buffer, [7] wait target (arrivals), [8] timeout ms,
// [9] status word (mapped host memory)
struct Peer {
    bool on = false;
    cldet_peer_exchange ex{};
    const void* flags_local = nullptr;
    const float* terms_local = nullptr;
    uint32_t target = 0;
    int timeout_ms = 0;
    int32_t* status = nullptr;
};

Peer parse_peer(at::IntArrayRef v) {
    Peer p;
    if (v.empty()) return p;
    TORCH_CHECK(v.size() == 10, "peer exchange descriptor must have 10 entries");
    p.on = true;
    p.ex.d_peer_terms = reinterpret_cast<void*>(v[0]);
    p.ex.d_peer_flags = reinterpret_cast<void*>(v[1]);
    p.ex.rank = (int32_t)v[2];
    p.ex.world = (int32_t)v[3];
    p.ex.parity = (int32_t)v[4];
    p.flags_local = reinterpret_cast<const void*>(v[5]);
    p.terms_local = reinterpret_cast<const float*>(v[6]);
    p.target = (uint32_t)v[7];
    p.timeout_ms = (int)v[8];
    p.status = reinterpret_cast<int32_t*>(v[9]);
    return p;
}

// What backward needs besides tensors, as ONE saved IValue (an int list): sizes, the peer slice and the params struct's words.
enum { kN = 0, kA, kC, kG, kNg, kLocal0, kLp0 };
constexpr size_t kLpWords = sizeof(cldet_loss_params) / sizeof(int32_t);
static_assert(sizeof(cldet_loss_params) % sizeof(int32_t) == 0, "cldet_loss_params is a sequence of 32-bit fields");

struct FocalLossFn : public torch::autograd::Function<FocalLossFn> {
    // outputs: bg[Ng], fg[Ng], reg_j[Ng], enh_j[Ng], reg_loss[1], npos[N], nvalid[N], meta, bg_mask?, status?
    // (Ng = N, or world*N rows in global image order on an image-sharded run)
    static variable_list forward(AutogradContext* ctx, const Tensor& cls, const Tensor& reg, const Tensor& anchors,
                                 const Tensor& ann, const Tensor& hint, cldet_loss_params lp, bool need_grad, bool want_bg_mask,
                                 bool want_status, Peer peer) {
        const int64_t n = cls.size(0), a = cls.size(1), c = cls.size(2), g = ann.size(1);
        c10::cuda::CUDAGuard guard(cls.device());
        cudaStream_t stream = c10::cuda::getCurrentCUDAStream();
        const auto f32 = cls.options();
        const auto i32 = cls.options().dtype(at::kInt);
        const int64_t ng = peer.on ? n * peer.ex.world : n;
        Tensor losses = at::empty({4, n}, f32);
        Tensor reg_loss = at::empty({1}, f32);
        Tensor meta = at::empty({n, a}, i32);
        Tensor iou_max = lp.decrease_positive_by_iou ? at::empty({n, a}, f32) : Tensor();
        Tensor counts = at::empty({2, n}, i32);
        Tensor bg_mask = want_bg_mask ? at::empty({n, a}, cls.options().dtype(at::kByte)) : Tensor();
        Tensor status = want_status ? at::empty({1}, i32) : Tensor();
        Tensor ws = loss_workspace(cls, stream, n, a);
        Tensor baked, gcls, greg;
        if (need_grad) {
            baked = at::empty({4, n}, f32);
            gcls = at::empty_like(cls);
            greg = at::empty_like(reg);
        }
        int32_t* npos = counts.data_ptr<int32_t>();
        int32_t* nvalid = npos + n;
        auto opt = [](const Tensor& t) -> void* { return t.defined() ? t.data_ptr() : nullptr; };
        // image-sharded run: the loss kernel pushes every image's terms into all ranks' buffers; its last block then waits for
        // all ranks' arrivals and copies the GLOBAL rows into `global` -- a private tensor, never a view of the exchange
        // buffer -- so no wait kernel is launched behind it (CLDET_PEER_SEPARATE_WAIT=1: the two-launch variant, A/B only)
        Tensor global;
        static const bool separate_wait = [] {
            const char* e = getenv("CLDET_PEER_SEPARATE_WAIT");
            return e && e[0] == '1';
        }();
        if (peer.on) {
            global = at::empty({4, ng}, f32);
            if (!separate_wait) {
                peer.ex.timeout_ms = peer.timeout_ms;
                peer.ex.target_arrivals = peer.target;
                peer.ex.d_flags_local = peer.flags_local;
                peer.ex.d_terms_local = peer.terms_local;
                peer.ex.d_wait_out = global.data_ptr<float>();
                peer.ex.d_wait_status = peer.status;
            }
        }
        int rc = cldet_focal_loss_sharded(
            cls.data_ptr<float>(), reg.data_ptr<float>(), anchors.data_ptr<float>(), ann.data_ptr<float>(), (int)n, a, (int)c,
            (int)g, &lp, need_grad ? hint.data_ptr<float>() : nullptr, (float*)opt(baked), (float*)opt(gcls), (float*)opt(greg),
            losses.data_ptr<float>(), (uint32_t*)meta.data_ptr(), (float*)opt(iou_max), npos, nvalid, (uint8_t*)opt(bg_mask),
            (int32_t*)opt(status), ws.data_ptr(), (size_t)ws.numel(), peer.on ? &peer.ex : nullptr, reg_loss.data_ptr<float>(), stream);
        if (rc != CLDET_OK) {
            drop_workspaces();          // a failed call may leave the scratch header dirty
            check_status(rc, "cldet_focal_loss");
        }
        if (peer.on) {
            if (separate_wait)
                check_status(cldet_peer_wait(peer.flags_local, peer.terms_local, peer.ex.world, (int)n, peer.ex.parity, peer.target,
                                             peer.timeout_ms, global.data_ptr<float>(), reg_loss.data_ptr<float>(), peer.status, stream),
                             "cldet_peer_wait");
            losses = global;
        }
        ctx->set_materialize_grads(false);          // absent upstream gradients stay undefined (= zero rows), no fill kernels
        if (need_grad) {
            std::vector<int64_t> info(kLp0 + kLpWords);
            info[kN] = n; info[kA] = a; info[kC] = c; info[kG] = g; info[kNg] = ng;
            info[kLocal0] = peer.on ? (int64_t)peer.ex.rank * n : (int64_t)0;
            const int32_t* words = reinterpret_cast<const int32_t*>(&lp);
            for (size_t i = 0; i < kLpWords; ++i) info[kLp0 + i] = words[i];
            ctx->saved_data["info"] = std::move(info);
            ctx->save_for_backward({cls, reg, anchors, ann, baked, gcls, greg, meta, counts, iou_max, ws});
        }
        std::vector<Tensor> rows = losses.unbind(0);
        std::vector<Tensor> cnt = counts.unbind(0);
        variable_list outs = {rows[0], rows[1], rows[2], rows[3], reg_loss, cnt[0], cnt[1], meta};
        std::vector<Tensor> nondiff = {cnt[0], cnt[1], meta};
        if (want_bg_mask) {
            outs.push_back(bg_mask);
            nondiff.push_back(bg_mask);
        }
        if (want_status) {
            outs.push_back(status);
            nondiff.push_back(status);
        }
        ctx->mark_non_differentiable(nondiff);
        return outs;
    }

    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        const auto saved = ctx->get_saved_variables();
        const Tensor &cls = saved[0], &reg = saved[1], &anchors = saved[2], &ann = saved[3], &baked = saved[4];
        Tensor gcls = saved[5], greg = saved[6];
        const Tensor &meta = saved[7], &counts = saved[8], &iou_max = saved[9], &ws = saved[10];
        const auto info = ctx->saved_data["info"].toIntVector();
        const int64_t n = info[kN], a = info[kA], c = info[kC], g = info[kG], ng = info[kNg], local0 = info[kLocal0];
        cldet_loss_params lp;
        {
            int32_t* words = reinterpret_cast<int32_t*>(&lp);
            for (size_t i = 0; i < kLpWords; ++i) words[i] = (int32_t)info[kLp0 + i];
        }
        c10::cuda::CUDAGuard guard(cls.device());
        cudaStream_t stream = c10::cuda::getCurrentCUDAStream();
        // rows: dL/dbg, dL/dfg, dL/dreg_j, dL/denh as (pointer, element stride); a row may arrive expanded (stride 0) or not
        // at all (undefined = zeros).  The regression term reaches the caller twice -- per image (reg_j) and as the batch mean
        // reg_loss = mean_j reg_j (what the reference returns): the kernel adds dL/dreg_loss * (1/Ng) to the per-image row.
        Tensor rows[5] = {grads[0], grads[1], grads[2], grads[3], grads[4]};
        const float* ptr[4];
        int64_t stride[4];
        for (int k = 0; k < 5; ++k)
            if (rows[k].defined() && (rows[k].scalar_type() != at::kFloat || !rows[k].is_cuda())) rows[k] = rows[k].to(cls.options());
        for (int k = 0; k < 4; ++k) {
            if (!rows[k].defined()) {
                ptr[k] = nullptr;
                stride[k] = 0;
                continue;
            }
            const int64_t st = (rows[k].dim() == 0 || rows[k].numel() == 1) ? 0 : rows[k].stride(0);
            ptr[k] = rows[k].data_ptr<float>() + ((st != 0) ? local0 * st : 0);   // global rows: this rank's images start at local0
            stride[k] = st;
        }
        const float* reg_mean_w = rows[4].defined() ? rows[4].data_ptr<float>() : nullptr;
        const int32_t* npos = counts.data_ptr<int32_t>();
        check_status(cldet_focal_loss_reweight_rows(
                         cls.data_ptr<float>(), reg.data_ptr<float>(), anchors.data_ptr<float>(), ann.data_ptr<float>(), (int)n, a,
                         (int)c, (int)g, &lp, ptr[0], stride[0], ptr[1], stride[1], ptr[2], stride[2], ptr[3], stride[3], reg_mean_w,
                         (float)(1.0 / (double)ng), baked.data_ptr<float>(), gcls.data_ptr<float>(), greg.data_ptr<float>(),
                         (const uint32_t*)meta.data_ptr(), iou_max.defined() ? iou_max.data_ptr<float>() : nullptr, npos,
                         ws.data_ptr(), (size_t)ws.numel(), stream),
                     "cldet_focal_loss_reweight_rows");
        auto it = ctx->saved_data.find("calls");
        const int64_t calls = (it == ctx->saved_data.end() ? 0 : it->second.toInt()) + 1;
        ctx->saved_data["calls"] = calls;
        if (calls > 1) {       // the buffers may already be someone's .grad: hand out copies from now on
            gcls = gcls.clone();
            greg = greg.clone();
        }
        return {gcls, greg, Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
    }
};

std::vector<Tensor> focal_loss(const Tensor& cls, const Tensor& reg, const Tensor& anchors, const Tensor& ann, const Tensor& hint,
                               double alpha, double gamma, int64_t incremental, int64_t past, int64_t ignore_past,
                               int64_t new_ignore_past, int64_t dec_by_iou, int64_t enhance, double dec_pos, int64_t height,
                               int64_t width, bool logits, bool want_bg_mask, bool want_status, at::IntArrayRef peer) {
    check_f32_cuda(cls, "classifications");
    check_f32_cuda(reg, "regressions");
    check_f32_cuda(anchors, "anchors");
    check_f32_cuda(ann, "annotations");
    check_f32_cuda(hint, "upstream_hint");
    TORCH_CHECK(cls.dim() == 3 && reg.dim() == 3 && reg.size(2) == 4 && reg.size(0) == cls.size(0) && reg.size(1) == cls.size(1),
                "classifications must be [N,A,C] and regressions [N,A,4]");
    TORCH_CHECK(anchors.numel() == cls.size(1) * 4, "anchors must be [1,A,4]");
    TORCH_CHECK(ann.dim() == 3 && ann.size(0) == cls.size(0) && ann.size(2) == 5 && ann.size(1) > 0, "annotations must be [N,G>=1,5]");
    TORCH_CHECK(hint.dim() == 2 && hint.size(0) == 4 && hint.size(1) == cls.size(0), "upstream_hint must be [4, N]");
    const bool need_grad = at::GradMode::is_enabled() && (cls.requires_grad() || reg.requires_grad());
    const cldet_loss_params lp =
        make_params(alpha, gamma, incremental, past, ignore_past, new_ignore_past, dec_by_iou, enhance, dec_pos, height, width, logits);
    return FocalLossFn::apply(cls, reg, anchors, ann, hint, lp, need_grad, want_bg_mask, want_status, parse_peer(peer));
}

// ---- eval-mode detection output for a batch: K4 decode+filter -> K5 ordering (+ optional top-k) -> K6 NMS -> gather ----
// Returns the padded result (scores[N,cap], labels[N,cap] int64, boxes[N,cap,4], counts[N] int32, candidates[N] int32) and
// NEVER synchronises.  With pre_nms_topk == 0 (the reference's mode) the per-image capacity `cap` is
// `capacity` when > 0, else optimistic_capacity() (sized so that the batch's NMS masks stay small): the caller reads counts and candidates back together (the one
// read it needs anyway to slice the results) and, only if some image had more candidates than cap, calls again with
// capacity = that count.
constexpr int64_t kOptimisticCapacity = 4096;           // at least this many candidates per image fit the first attempt
constexpr int64_t kOptimisticMaskBytes = 128ll << 20;   // ... and as many as keep the batch's suppression masks under this

int64_t optimistic_capacity(int64_t n, int64_t a) {
    // mask bytes of a batch = n * cap^2 / 8: one image gets ~32 k candidates, 32 images ~5.8 k each
    int64_t cap = (int64_t)std::sqrt((double)kOptimisticMaskBytes * 8.0 / (double)std::max<int64_t>(n, 1));
    cap = std::max<int64_t>(kOptimisticCapacity, cap / 64 * 64);
    return std::min<int64_t>(a, cap);
}

std::vector<Tensor> detect(const Tensor& cls, const Tensor& reg, const Tensor& anchors, int64_t height, int64_t width, bool is_logits,
                           double score_thresh, double iou_thresh, int64_t pre_nms_topk, int64_t nms_mode,
                           int64_t vanilla_numel_limit, int64_t capacity) {
    check_f32_cuda(cls, "classifications");
    check_f32_cuda(reg, "regressions");
    check_f32_cuda(anchors, "anchors");
    TORCH_CHECK(cls.dim() == 3 && reg.dim() == 3 && reg.size(2) == 4 && reg.size(0) == cls.size(0) && reg.size(1) == cls.size(1) &&
                    anchors.numel() == cls.size(1) * 4,
                "expected cls [N,A,C], regressions [N,A,4], anchors [1,A,4]");
    TORCH_CHECK(((uintptr_t)reg.data_ptr() & 15) == 0 && ((uintptr_t)anchors.data_ptr() & 15) == 0,
                "regressions and anchors must be 16-byte aligned");
    const int64_t n = cls.size(0), a = cls.size(1), c = cls.size(2);
    c10::cuda::CUDAGuard guard(cls.device());
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream();
    const auto f32 = cls.options();
    const auto i32 = f32.dtype(at::kInt);
    const auto u8 = f32.dtype(at::kByte);
    const int topk = pre_nms_topk > 0 ? (int)pre_nms_topk : 0;
    Tensor counts = at::zeros({n}, i32);
    Tensor cand = at::empty({n, a, (int64_t)sizeof(cldet_candidate)}, u8);
    Tensor keys = at::empty({n, a}, f32.dtype(at::kLong));
    check_status(cldet_decode_filter(cls.data_ptr<float>(), is_logits ? 1 : 0, reg.data_ptr<float>(), anchors.data_ptr<float>(), (int)n, a,
                                     (int)c, (int)height, (int)width, (float)score_thresh, (cldet_candidate*)cand.data_ptr(),
                                     (uint64_t*)keys.data_ptr(), a, counts.data_ptr<int32_t>(), stream),
                 "cldet_decode_filter");
    int64_t max_count, cap;
    if (topk) {
        max_count = a;
        cap = std::min<int64_t>(topk, a);
    } else {
        cap = max_count = capacity > 0 ? std::min<int64_t>(a, capacity) : optimistic_capacity(n, a);
    }
    Tensor sorted = at::empty({n, cap, (int64_t)sizeof(cldet_candidate)}, u8);
    Tensor sorted_counts = at::empty({n}, i32);
    const size_t sws_bytes = cldet_sort_workspace_bytes((int)n, max_count, topk);
    Tensor sws = at::empty({(int64_t)sws_bytes}, u8);
    check_status(cldet_sort_candidates((const cldet_candidate*)cand.data_ptr(), (const uint64_t*)keys.data_ptr(), counts.data_ptr<int32_t>(),
                                       (int)n, a, max_count, topk, (cldet_candidate*)sorted.data_ptr(), cap,
                                       sorted_counts.data_ptr<int32_t>(), sws.data_ptr(), sws_bytes, stream),
                 "cldet_sort_candidates");
    // the suppression mask is cap^2/8 bytes per image (200 MB at 40 k candidates): with long lists, run the NMS over groups of
    // images that fit a fixed budget instead of sizing one workspace for the whole batch
    Tensor keep = at::empty({n, cap}, i32);
    Tensor keep_counts = at::empty({n}, i32);
    Tensor scores = at::empty({n, cap}, f32);
    Tensor labels = at::empty({n, cap}, f32.dtype(at::kLong));
    Tensor boxes = at::empty({n, cap, 4}, f32);
    const size_t one_image = cldet_nms_workspace_bytes(1, cap);
    const size_t budget = (size_t)4 << 30;
    const int64_t group = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)(budget / std::max<size_t>(one_image, 1))));
    const size_t nws_bytes = cldet_nms_workspace_bytes((int)group, cap);
    Tensor nws = at::empty({(int64_t)nws_bytes}, u8);
    for (int64_t j0 = 0; j0 < n; j0 += group) {
        const int64_t cnt = std::min<int64_t>(group, n - j0);
        // NMS + gather of the kept candidates in one chain (the resolving block gathers its image)
        check_status(cldet_nms_gather_sorted((const cldet_candidate*)sorted.data_ptr() + j0 * cap, sorted_counts.data_ptr<int32_t>() + j0,
                                             (int)cnt, cap, cap, (float)iou_thresh, (int)nms_mode, vanilla_numel_limit,
                                             keep.data_ptr<int32_t>() + j0 * cap, keep_counts.data_ptr<int32_t>() + j0,
                                             scores.data_ptr<float>() + j0 * cap, labels.data_ptr<int64_t>() + j0 * cap,
                                             boxes.data_ptr<float>() + j0 * cap * 4, nws.data_ptr(), nws_bytes, stream),
                     "cldet_nms_gather_sorted");
    }
    return {scores, labels, boxes, keep_counts, counts};
}

// torchvision.ops.batched_nms / nms drop-in: returns (keep[K] int64 padded, count[1] int32)
std::vector<Tensor> batched_nms(const Tensor& boxes, const Tensor& scores, const c10::optional<Tensor>& idxs, double iou_thresh,
                                int64_t mode, int64_t vanilla_numel_limit) {
    check_f32_cuda(boxes, "boxes");
    check_f32_cuda(scores, "scores");
    const int64_t k = scores.numel();
    TORCH_CHECK(boxes.numel() == 4 * k, "boxes and scores disagree");
    TORCH_CHECK(((uintptr_t)boxes.data_ptr() & 15) == 0, "boxes must be 16-byte aligned");
    c10::cuda::CUDAGuard guard(boxes.device());
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream();
    const auto opts = boxes.options();
    Tensor keep = at::empty({k}, opts.dtype(at::kLong));
    Tensor count = at::zeros({1}, opts.dtype(at::kInt));
    if (k == 0) return {keep, count};
    Tensor ids;
    if (idxs.has_value() && idxs->defined()) {
        ids = idxs->to(at::kLong).contiguous();
        TORCH_CHECK(ids.numel() == k && ids.is_cuda(), "idxs must be a CUDA tensor with one entry per box");
    }
    const size_t ws_bytes = cldet_batched_nms_workspace_bytes(k);
    Tensor ws = at::empty({(int64_t)ws_bytes}, opts.dtype(at::kByte));
    check_status(cldet_batched_nms(boxes.data_ptr<float>(), scores.data_ptr<float>(), ids.defined() ? ids.data_ptr<int64_t>() : nullptr, k,
                                   (float)iou_thresh, ids.defined() ? (int)mode : 1, vanilla_numel_limit, keep.data_ptr<int64_t>(),
                                   count.data_ptr<int32_t>(), ws.data_ptr(), ws_bytes, stream),
                 "cldet_batched_nms");
    return {keep, count};
}

int64_t abi_version() { return cldet_abi_version(); }

}  // namespace

TORCH_LIBRARY(cldet, m) {
    m.def(
        "focal_loss(Tensor cls, Tensor reg, Tensor anchors, Tensor annotations, Tensor hint, float alpha, float gamma, "
        "int incremental, int past_class_num, int ignore_past_class, int new_ignore_past_class, int decrease_positive_by_iou, "
        "int enhance_on_new, float decrease_positive, int image_height, int image_width, bool cls_is_logits, bool want_bg_mask, "
        "bool want_status, int[] peer) -> Tensor[]",
        &focal_loss);
    m.def(
        "detect(Tensor cls, Tensor reg, Tensor anchors, int height, int width, bool is_logits, float score_thresh, "
        "float iou_thresh, int pre_nms_topk, int nms_mode, int vanilla_numel_limit, int capacity) -> Tensor[]",
        &detect);
    m.def("batched_nms(Tensor boxes, Tensor scores, Tensor? idxs, float iou_thresh, int mode, int vanilla_numel_limit) -> Tensor[]",
          &batched_nms);
    m.def("abi_version() -> int", &abi_version);
}
