// Shared device/host helpers for libcldet (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cldet.h"

namespace cldet {

// Thread-local record of the last failing CUDA call (no shared mutable state between host threads).
void set_last_cuda_error(cudaError_t e);

#define CLDET_CUDA_TRY(expr)                       \
    do {                                           \
        cudaError_t _e = (expr);                   \
        if (_e != cudaSuccess) {                   \
            ::cldet::set_last_cuda_error(_e);      \
            return CLDET_ERR_CUDA;                 \
        }                                          \
    } while (0)

#define CLDET_LAUNCH_CHECK() CLDET_CUDA_TRY(cudaPeekAtLastError())

constexpr int kNumLevels = 5;
constexpr int kAnchorsPerCell = 9;

__host__ __device__ __forceinline__ uint32_t meta_pack(uint32_t state, uint32_t label, uint32_t raw_row) {
    return (state & 3u) | ((label & 0x1fffu) << 3) | (raw_row << 16);
}
__host__ __device__ __forceinline__ uint32_t meta_state(uint32_t m) { return m & 3u; }
__host__ __device__ __forceinline__ uint32_t meta_label(uint32_t m) { return (m >> 3) & 0x1fffu; }
__host__ __device__ __forceinline__ uint32_t meta_row(uint32_t m) { return m >> 16; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Streaming 128-bit accesses: data touched exactly once, keep it out of L1.
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// GT-centric assignment for the standard anchor grid of an (height, width) image (cldet_assign.cu): for every valid GT row,
// visit only the anchors that can reach IoU >= 0.4 with it and atomicMax the key (exact IoU bits << 32 | ~row) into best[N,A]
// (zero on entry);
// counts anchors crossing 0.5 into npos_acc and valid rows into nvalid.  Returns CLDET_ERR_UNSUPPORTED when A does not match
// the grid of (height, width).
int launch_gt_scatter(int height, int width, const float* d_anchors, int64_t num_anchors, const float* d_annotations,
                      int num_images, int gt_rows, unsigned long long* d_best, uint32_t* d_touched, int32_t* d_npos_acc,
                      int32_t* d_nvalid, cudaStream_t s);

// ---- programmatic dependent launch (PDL) ----
// A kernel launched with launch_pdl() may be scheduled while its predecessor on the stream is still draining (the
// predecessor's blocks have all started and called pdl_launch_dependents(), or exited); it must call pdl_wait() -- in EVERY
// block, before the first access to memory the predecessor touches -- which returns once the predecessor has completed and
// its writes are visible.  Both instructions are no-ops in a kernel that was launched normally.  The chains here are bound by
// the latency of short dependent launches, which this hides.  CLDET_NO_PDL=1 launches everything normally (A/B only).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- bulk asynchronous copies (the TMA unit's 1-D mode: cp.async.bulk, SASS UBLKCP) + mbarrier completion ----
// A contiguous, 16-byte aligned run of bytes moves between global and shared memory without passing through registers: one
// thread issues it, the copy engine lands the bytes and signals an mbarrier (loads) / a bulk group (stores).  Memory-level
// parallelism then no longer depends on how many loads the warps hold in registers while they compute.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// makes mbarrier initialisations (and earlier generic-proxy writes to shared memory) visible to the async proxy
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst_gmem, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest `N` committed store groups have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

inline int sm_count() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

}  // namespace cldet
