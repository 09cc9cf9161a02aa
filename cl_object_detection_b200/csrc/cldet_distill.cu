// the head-distillation terms of IL_Loss (retinanet/losses.py:705-737) as two fused passes.
//   prev_fg_mask = sigmoid(prev_cls) > 0.05                      [N,A,P]   (P = number of past classes)
//   reg_mask     = bg_masks & prev_fg_mask.any(2)                [N,A]
//   dist_reg     = SmoothL1Loss()(prev_reg[reg_mask], reg[reg_mask])                 (beta 1, mean over rows*4)
//   dist_cls     = MSELoss()(prev[prev_fg_mask], cur[prev_fg_mask])                  (ignore_GD: rows of reg_mask instead)
//                  on logits (distill_logits) or on probabilities (sigmoid of both)
// Forward: one pass over prev_cls / cls[:, :, :P] / reg / prev_reg -> two sums + two counts (per-block partials reduced in
// fixed order).  Backward: one pass writing dL/dcls [N,A,C] (zeros for the new-class columns) and dL/dreg [N,A,4].
// The reference does this with ~25 eager kernels and four boolean-mask gathers (each a host sync).
#include <math.h>

#include <algorithm>

#include "cldet_common.cuh"

namespace cldet {

__device__ __forceinline__ float sigmoid_exact_d(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

constexpr int kDistillThreads = 256;

struct DistillArgs {
    const float* cls;        // [N,A,C] logits of the current model
    const float* prev_cls;   // [N,A,P] logits of the previous model
    const float* reg;        // [N,A,4]
    const float* prev_reg;   // [N,A,4]
    const uint8_t* bg_mask;  // [N,A] 1 = anchor is not positive (FocalLoss 'bg_masks')
    int64_t rows;            // N*A
    int C, P;
    int use_logits;          // distill_logits
    int ignore_gd;           // ignore_GD
};

// per anchor row: returns reg_mask and accumulates the forward sums
__device__ __forceinline__ bool distill_row_forward(const DistillArgs& a, int64_t row, float& sum_cls, int& cnt_cls,
                                                    float& sum_reg, int& cnt_reg) {
    const float* pc = a.prev_cls + row * a.P;
    const float* cc = a.cls + row * a.C;
    bool any_fg = false;
    float s_fg = 0.0f, s_all = 0.0f;
    int n_fg = 0;
    for (int c = 0; c < a.P; ++c) {
        const float pl = pc[c];
        const float pp = sigmoid_exact_d(pl);
        const bool fg = pp > 0.05f;
        any_fg |= fg;
        const float cur = a.use_logits ? cc[c] : sigmoid_exact_d(cc[c]);
        const float prv = a.use_logits ? pl : pp;
        const float d = prv - cur;
        const float sq = d * d;
        s_all += sq;
        if (fg) {
            s_fg += sq;
            ++n_fg;
        }
    }
    const bool rm = any_fg && a.bg_mask[row] != 0;
    if (a.ignore_gd) {
        if (rm) {
            sum_cls += s_all;
            cnt_cls += a.P;
        }
    } else {
        sum_cls += s_fg;
        cnt_cls += n_fg;
    }
    if (rm) {
        const float4 x = *reinterpret_cast<const float4*>(a.prev_reg + row * 4);
        const float4 y = *reinterpret_cast<const float4*>(a.reg + row * 4);
        const float d[4] = {x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float z = fabsf(d[i]);
            sum_reg += (z < 1.0f) ? 0.5f * z * z : z - 0.5f;
        }
        cnt_reg += 4;
    }
    return rm;
}

__global__ void __launch_bounds__(kDistillThreads)
distill_forward_kernel(const DistillArgs a, double* __restrict__ partial_sums, long long* __restrict__ partial_counts) {
    __shared__ float rs[2][kDistillThreads / 32];
    __shared__ int rc[2][kDistillThreads / 32];
    float sum_cls = 0.f, sum_reg = 0.f;
    int cnt_cls = 0, cnt_reg = 0;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < a.rows; row += (int64_t)gridDim.x * blockDim.x)
        distill_row_forward(a, row, sum_cls, cnt_cls, sum_reg, cnt_reg);
    sum_cls = warp_sum(sum_cls);
    sum_reg = warp_sum(sum_reg);
    cnt_cls = warp_sum_int(cnt_cls);
    cnt_reg = warp_sum_int(cnt_reg);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        rs[0][warp] = sum_cls; rs[1][warp] = sum_reg;
        rc[0][warp] = cnt_cls; rc[1][warp] = cnt_reg;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        double s = 0.0;
        long long c = 0;
        for (int w = 0; w < kDistillThreads / 32; ++w) {
            s += (double)rs[threadIdx.x][w];
            c += rc[threadIdx.x][w];
        }
        partial_sums[2 * blockIdx.x + threadIdx.x] = s;
        partial_counts[2 * blockIdx.x + threadIdx.x] = c;
    }
}

// out[0] = dist_cls_loss, out[1] = dist_reg_loss (0/0 = NaN when nothing is selected, like the reference's mean of an
// empty tensor); counts[0..1] as float for the backward pass
__global__ void distill_finalize_kernel(const double* __restrict__ partial_sums, const long long* __restrict__ partial_counts,
                                        int nblocks, float* __restrict__ out, float* __restrict__ counts) {
    if (threadIdx.x < 2) {
        double s = 0.0;
        long long c = 0;
        for (int b = 0; b < nblocks; ++b) {
            s += partial_sums[2 * b + threadIdx.x];
            c += partial_counts[2 * b + threadIdx.x];
        }
        out[threadIdx.x] = (float)(s / (double)c);
        counts[threadIdx.x] = (float)c;
    }
}

__global__ void __launch_bounds__(kDistillThreads)
distill_backward_kernel(const DistillArgs a, const float* __restrict__ counts, const float* __restrict__ g_cls_loss,
                        const float* __restrict__ g_reg_loss, float* __restrict__ grad_cls, float* __restrict__ grad_reg) {
    // d mean((prev-cur)^2)/dcur = -2 (prev-cur)/K ; d mean(smoothL1(prev-cur))/dcur = -(clip to [-1,1])(prev-cur)/K
    const float k_cls = (g_cls_loss ? *g_cls_loss : 0.0f) * 2.0f / counts[0];
    const float k_reg = (g_reg_loss ? *g_reg_loss : 0.0f) / counts[1];
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < a.rows; row += (int64_t)gridDim.x * blockDim.x) {
        const float* pc = a.prev_cls + row * a.P;
        const float* cc = a.cls + row * a.C;
        float* gc = grad_cls + row * a.C;
        bool any_fg = false;
        for (int c = 0; c < a.P; ++c) any_fg |= sigmoid_exact_d(pc[c]) > 0.05f;
        const bool rm = any_fg && a.bg_mask[row] != 0;
        for (int c = 0; c < a.P; ++c) {
            const float pl = pc[c];
            const float pp = sigmoid_exact_d(pl);
            const bool sel = a.ignore_gd ? rm : (pp > 0.05f);
            float g = 0.0f;
            if (sel) {
                if (a.use_logits) {
                    g = -(pl - cc[c]) * k_cls;
                } else {
                    const float cp = sigmoid_exact_d(cc[c]);
                    g = (-(pp - cp) * k_cls) * (1.0f - cp) * cp;      // through the Sigmoid at losses.py:715
                }
            }
            gc[c] = g;
        }
        for (int c = a.P; c < a.C; ++c) gc[c] = 0.0f;                 // new-class columns are not distilled (losses.py:705)
        float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rm) {
            const float4 x = *reinterpret_cast<const float4*>(a.prev_reg + row * 4);
            const float4 y = *reinterpret_cast<const float4*>(a.reg + row * 4);
            const float d[4] = {x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w};
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = -fminf(fmaxf(d[i], -1.0f), 1.0f) * k_reg;
            gr = make_float4(o[0], o[1], o[2], o[3]);
        }
        *reinterpret_cast<float4*>(grad_reg + row * 4) = gr;
    }
}

// ------------------------------------------------------------------------------------------------
// enhance_error on replay batches (retinanet/losses.py:590-603): over the NEW-class columns (c >= past) of the class
// probabilities, the elements > 0.05 contribute |p| (L1), p^2 (L2) or p^3 (L3); loss = sum / max(count, 1).
// Flat coalesced sweep over [rows, C]; per-block partials, fixed-order fp64 finalize (bit-reproducible).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float enhance_term(float p, int method) { return method == 1 ? fabsf(p) : (method == 2 ? p * p : (p * p) * p); }

__global__ void __launch_bounds__(kDistillThreads)
enhance_error_forward_kernel(const float* __restrict__ cls, int64_t total, int C, int past, int method,
                             double* __restrict__ partial_sums, long long* __restrict__ partial_counts) {
    __shared__ float rs[kDistillThreads / 32];
    __shared__ int rc[kDistillThreads / 32];
    float sum = 0.f;
    int cnt = 0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(e % C);
        if (col < past) continue;
        const float p = cls[e];
        if (p > 0.05f) {
            sum += enhance_term(p, method);
            ++cnt;
        }
    }
    sum = warp_sum(sum);
    cnt = warp_sum_int(cnt);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        rs[warp] = sum;
        rc[warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        long long c = 0;
        for (int w = 0; w < kDistillThreads / 32; ++w) {
            s += (double)rs[w];
            c += rc[w];
        }
        partial_sums[blockIdx.x] = s;
        partial_counts[blockIdx.x] = c;
    }
}

__global__ void enhance_error_finalize_kernel(const double* __restrict__ partial_sums, const long long* __restrict__ partial_counts,
                                              int nblocks, float* __restrict__ out, float* __restrict__ count) {
    if (threadIdx.x == 0) {
        double s = 0.0;
        long long c = 0;
        for (int b = 0; b < nblocks; ++b) {
            s += partial_sums[b];
            c += partial_counts[b];
        }
        const float denom = (float)(c > 1 ? c : 1);                      // max(classification.shape[0], 1)
        *out = (float)s / denom;
        *count = denom;
    }
}

__global__ void __launch_bounds__(kDistillThreads)
enhance_error_backward_kernel(const float* __restrict__ cls, int64_t total, int C, int past, int method,
                              const float* __restrict__ count, const float* __restrict__ g_loss, float* __restrict__ grad) {
    const float k = (g_loss ? *g_loss : 0.0f) / *count;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(e % C);
        float g = 0.0f;
        if (col >= past) {
            const float p = cls[e];
            if (p > 0.05f) g = (method == 1 ? 1.0f : (method == 2 ? 2.0f * p : 3.0f * (p * p))) * k;
        }
        grad[e] = g;
    }
}

// ------------------------------------------------------------------------------------------------
// MAS Output_norm's regression term (IL_method/mas.py:52-55): per image, the mean of |regression| over the rows of its
// positive anchors (0 for an image without positives), summed over images -- without the reference's per-image boolean
// gather + host sync.  One block per image; fixed summation order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDistillThreads)
masked_abs_mean_forward_kernel(const float4* __restrict__ reg, const uint8_t* __restrict__ positive, int64_t A,
                               float* __restrict__ terms, float* __restrict__ counts) {
    __shared__ double rs[kDistillThreads / 32];
    __shared__ int rc[kDistillThreads / 32];
    const int j = blockIdx.x;
    double sum = 0.0;
    int cnt = 0;
    for (int64_t an = threadIdx.x; an < A; an += blockDim.x) {
        if (positive[(int64_t)j * A + an]) {
            const float4 r = reg[(int64_t)j * A + an];
            sum += (double)fabsf(r.x) + (double)fabsf(r.y) + (double)fabsf(r.z) + (double)fabsf(r.w);
            ++cnt;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt = warp_sum_int(cnt);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        rs[warp] = sum;
        rc[warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        int c = 0;
        for (int w = 0; w < kDistillThreads / 32; ++w) {
            s += rs[w];
            c += rc[w];
        }
        terms[j] = c > 0 ? (float)(s / (4.0 * (double)c)) : 0.0f;
        counts[j] = (float)c;
    }
}

__global__ void __launch_bounds__(kDistillThreads)
masked_abs_mean_backward_kernel(const float4* __restrict__ reg, const uint8_t* __restrict__ positive, int64_t A,
                                const float* __restrict__ counts, const float* __restrict__ g_terms, int64_t g_stride,
                                float4* __restrict__ grad) {
    const int j = blockIdx.y;
    const float cnt = counts[j];
    const float k = cnt > 0.0f ? g_terms[(int64_t)j * g_stride] / (4.0f * cnt) : 0.0f;
    for (int64_t an = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; an < A; an += (int64_t)gridDim.x * blockDim.x) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (positive[(int64_t)j * A + an]) {
            const float4 r = reg[(int64_t)j * A + an];
            auto sgn = [](float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); };
            g = make_float4(sgn(r.x) * k, sgn(r.y) * k, sgn(r.z) * k, sgn(r.w) * k);
        }
        grad[(int64_t)j * A + an] = g;
    }
}

static int distill_blocks(int64_t rows) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((rows + kDistillThreads - 1) / kDistillThreads, (int64_t)sm_count() * 8));
}

static int fill_args(DistillArgs& a, const float* d_cls, const float* d_prev_cls, const float* d_reg, const float* d_prev_reg,
                     const uint8_t* d_bg_mask, int num_images, int64_t num_anchors, int num_classes, int past, int use_logits,
                     int ignore_gd) {
    if (!d_cls || !d_prev_cls || !d_reg || !d_prev_reg || !d_bg_mask) return CLDET_ERR_INVALID_ARGUMENT;
    if (num_images <= 0 || num_anchors <= 0 || num_classes <= 0 || past <= 0 || past > num_classes) return CLDET_ERR_INVALID_ARGUMENT;
    if (((uintptr_t)d_reg | (uintptr_t)d_prev_reg) & 15) return CLDET_ERR_INVALID_ARGUMENT;
    a.cls = d_cls; a.prev_cls = d_prev_cls; a.reg = d_reg; a.prev_reg = d_prev_reg; a.bg_mask = d_bg_mask;
    a.rows = (int64_t)num_images * num_anchors; a.C = num_classes; a.P = past; a.use_logits = use_logits ? 1 : 0;
    a.ignore_gd = ignore_gd ? 1 : 0;
    return CLDET_OK;
}

}  // namespace cldet

using namespace cldet;

extern "C" {

size_t cldet_distill_workspace_bytes(int num_images, int64_t num_anchors) {
    if (num_images <= 0 || num_anchors <= 0) return 0;
    const int nb = distill_blocks((int64_t)num_images * num_anchors);
    return (size_t)nb * 2 * (sizeof(double) + sizeof(long long)) + 256;
}

int cldet_distill_forward(const float* d_cls, const float* d_prev_cls, const float* d_reg, const float* d_prev_reg,
                          const uint8_t* d_bg_mask, int num_images, int64_t num_anchors, int num_classes, int past_class_num,
                          int distill_logits, int ignore_gd, float* d_losses, float* d_counts, void* d_workspace,
                          size_t workspace_bytes, void* stream) {
    DistillArgs a;
    int rc = fill_args(a, d_cls, d_prev_cls, d_reg, d_prev_reg, d_bg_mask, num_images, num_anchors, num_classes, past_class_num,
                       distill_logits, ignore_gd);
    if (rc) return rc;
    if (!d_losses || !d_counts || !d_workspace) return CLDET_ERR_INVALID_ARGUMENT;
    if (workspace_bytes < cldet_distill_workspace_bytes(num_images, num_anchors)) return CLDET_ERR_WORKSPACE_TOO_SMALL;
    const int nb = distill_blocks(a.rows);
    double* ps = reinterpret_cast<double*>(d_workspace);
    long long* pc = reinterpret_cast<long long*>(ps + 2 * (size_t)nb);
    cudaStream_t s = (cudaStream_t)stream;
    distill_forward_kernel<<<nb, kDistillThreads, 0, s>>>(a, ps, pc);
    CLDET_LAUNCH_CHECK();
    distill_finalize_kernel<<<1, 32, 0, s>>>(ps, pc, nb, d_losses, d_counts);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

int cldet_distill_backward(const float* d_cls, const float* d_prev_cls, const float* d_reg, const float* d_prev_reg,
                           const uint8_t* d_bg_mask, int num_images, int64_t num_anchors, int num_classes, int past_class_num,
                           int distill_logits, int ignore_gd, const float* d_counts, const float* d_grad_cls_loss,
                           const float* d_grad_reg_loss, float* d_grad_cls, float* d_grad_reg, void* stream) {
    DistillArgs a;
    int rc = fill_args(a, d_cls, d_prev_cls, d_reg, d_prev_reg, d_bg_mask, num_images, num_anchors, num_classes, past_class_num,
                       distill_logits, ignore_gd);
    if (rc) return rc;
    if (!d_counts || !d_grad_cls || !d_grad_reg || ((uintptr_t)d_grad_reg & 15)) return CLDET_ERR_INVALID_ARGUMENT;
    distill_backward_kernel<<<distill_blocks(a.rows), kDistillThreads, 0, (cudaStream_t)stream>>>(a, d_counts, d_grad_cls_loss,
                                                                                                d_grad_reg_loss, d_grad_cls,
                                                                                                d_grad_reg);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

size_t cldet_enhance_error_workspace_bytes(int64_t num_elements) {
    if (num_elements <= 0) return 0;
    return (size_t)distill_blocks(num_elements) * (sizeof(double) + sizeof(long long)) + 256;
}

int cldet_enhance_error_forward(const float* d_cls, int num_images, int64_t num_anchors, int num_classes, int past_class_num,
                                int method, float* d_loss, float* d_count, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!d_cls || !d_loss || !d_count || !d_workspace) return CLDET_ERR_INVALID_ARGUMENT;
    if (num_images <= 0 || num_anchors <= 0 || num_classes <= 0 || past_class_num < 0 || past_class_num > num_classes || method < 1 ||
        method > 3)
        return CLDET_ERR_INVALID_ARGUMENT;
    const int64_t total = (int64_t)num_images * num_anchors * num_classes;
    if (workspace_bytes < cldet_enhance_error_workspace_bytes(total)) return CLDET_ERR_WORKSPACE_TOO_SMALL;
    const int nb = distill_blocks(total);
    double* ps = reinterpret_cast<double*>(d_workspace);
    long long* pc = reinterpret_cast<long long*>(ps + nb);
    cudaStream_t s = (cudaStream_t)stream;
    enhance_error_forward_kernel<<<nb, kDistillThreads, 0, s>>>(d_cls, total, num_classes, past_class_num, method, ps, pc);
    CLDET_LAUNCH_CHECK();
    enhance_error_finalize_kernel<<<1, 32, 0, s>>>(ps, pc, nb, d_loss, d_count);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

int cldet_enhance_error_backward(const float* d_cls, int num_images, int64_t num_anchors, int num_classes, int past_class_num,
                                 int method, const float* d_count, const float* d_grad_loss, float* d_grad_cls, void* stream) {
    if (!d_cls || !d_count || !d_grad_cls) return CLDET_ERR_INVALID_ARGUMENT;
    if (num_images <= 0 || num_anchors <= 0 || num_classes <= 0 || past_class_num < 0 || past_class_num > num_classes || method < 1 ||
        method > 3)
        return CLDET_ERR_INVALID_ARGUMENT;
    const int64_t total = (int64_t)num_images * num_anchors * num_classes;
    enhance_error_backward_kernel<<<distill_blocks(total), kDistillThreads, 0, (cudaStream_t)stream>>>(
        d_cls, total, num_classes, past_class_num, method, d_count, d_grad_loss, d_grad_cls);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

int cldet_masked_abs_mean_forward(const float* d_reg, const uint8_t* d_positive, int num_images, int64_t num_anchors,
                                  float* d_terms, float* d_counts, void* stream) {
    if (!d_reg || !d_positive || !d_terms || !d_counts || num_images <= 0 || num_anchors <= 0 || ((uintptr_t)d_reg & 15))
        return CLDET_ERR_INVALID_ARGUMENT;
    masked_abs_mean_forward_kernel<<<num_images, kDistillThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(d_reg), d_positive, num_anchors, d_terms, d_counts);
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

int cldet_masked_abs_mean_backward(const float* d_reg, const uint8_t* d_positive, int num_images, int64_t num_anchors,
                                   const float* d_counts, const float* d_grad_terms, int64_t grad_stride, float* d_grad_reg,
                                   void* stream) {
    if (!d_reg || !d_positive || !d_counts || !d_grad_terms || !d_grad_reg || num_images <= 0 || num_images > 65535 ||
        num_anchors <= 0 || (((uintptr_t)d_reg | (uintptr_t)d_grad_reg) & 15) || grad_stride < 0)
        return CLDET_ERR_INVALID_ARGUMENT;
    dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((num_anchors + kDistillThreads - 1) / kDistillThreads, 64)),
              (unsigned)num_images);
    masked_abs_mean_backward_kernel<<<grid, kDistillThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(d_reg), d_positive, num_anchors, d_counts, d_grad_terms, grad_stride,
        reinterpret_cast<float4*>(d_grad_reg));
    CLDET_LAUNCH_CHECK();
    return CLDET_OK;
}

}  // extern "C"
