// Shared device/host helpers for the vmasr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/vmasr_b200.h"

namespace vmasr {

// ---- error plumbing (api.cu) -----------------------------------------------------------------------
int fail(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
        target = dev;
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != target) cudaSetDevice(prev);
    }
    int target = -1;
};

int sm_count(int device);

// One-time-per-DEVICE flag for function attributes (cudaFuncSetAttribute is per device context: a process that drives
// several GPUs must set the dynamic shared-memory limit on each of them).
struct PerDeviceOnce {
    bool done[64] = {};
    bool &operator()() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
        return done[dev];
    }
};

// ---- tuning knobs --------------------------------------------------------------------------------------------------
// The product library reads no environment variables.  A build with -DVMASR_TUNING (make TUNING=1) turns the knobs of
// DESIGN.md 5.1 back on for measurement sessions.
inline const char *tuning_env(const char *name) {
#ifdef VMASR_TUNING
    return getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}

// ---- per-CTA phase timeline (VMASR_TUNING builds only) ------------------------------------------------------------------
// tools/timeline.py hands the library a device buffer of 16 timestamps per CTA (vmasr_debug_timeline); the multi-chunk scan
// kernels then record %globaltimer at their phase boundaries.  Compiles to nothing in the product build.
#ifdef VMASR_TUNING
unsigned long long *debug_timeline();
#define VMASR_TL(args, slot)                                                                    \
    do {                                                                                        \
        if ((args).timeline) {                                                                  \
            unsigned long long t_;                                                              \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                              \
            (args).timeline[(size_t)blockIdx.x * 16 + (slot)] = t_;                             \
        }                                                                                       \
    } while (0)
#else
#define VMASR_TL(args, slot) do { } while (0)
#endif

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// The scan kernels of a model run back to back on one stream.  Launched with the programmatic-stream-serialization
// attribute, the CTAs of kernel N + 1 are scheduled while the last wave of kernel N drains; they park on
// `griddepcontrol.wait` (first statement that touches global memory) until kernel N has completed and flushed.
// What overlaps is the launch latency and the CTA start-up.  Measured on the bench step (34 launches in one CUDA graph): nothing
// while every kernel ran one-tile CTAs (+0.7 % / -0.4 %, rounds 1 and 2); 0.6 % (same box, four alternating runs) to 1 % now
// that the backward's persistent CTAs have a longer start-up (barriers, producer warp, first copies).  ON in the product
// library; measurement builds turn it off with VMASR_PDL=0.  Every kernel launched this way executes griddepcontrol.wait
// before its first global access, so it is correct behind any predecessor, PDL-aware or not.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
int launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, const char *what, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return check_cuda(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...), what);
}
// device side: dependents may be scheduled from now on / wait for the grid this one depends on
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- dtype helpers ---------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- ITEMS consecutive elements per thread, 128-bit (or 64-bit) vector access when aligned ----------
// `row` points at element 0 of the (b, d) row, `pos` is the first position of this thread (a multiple of
// ITEMS), `len` the row length.  Out-of-range positions read as `fill`.
template <typename T, int ITEMS, bool VEC>
__device__ __forceinline__ void load_items(const T *__restrict__ row, int pos, int len, float (&v)[ITEMS], float fill) {
    if (VEC && pos + ITEMS <= len) {
        constexpr int BYTES = ITEMS * sizeof(T);
        if constexpr (BYTES == 32) {
            const float4 *p = reinterpret_cast<const float4 *>(row + pos);
            float4 a = __ldg(p), b = __ldg(p + 1);
            alignas(16) T tmp[ITEMS];
            *reinterpret_cast<float4 *>(tmp) = a;
            *reinterpret_cast<float4 *>(reinterpret_cast<char *>(tmp) + 16) = b;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) v[i] = to_f32<T>(tmp[i]);
        } else if constexpr (BYTES == 16) {
            float4 a = __ldg(reinterpret_cast<const float4 *>(row + pos));
            alignas(16) T tmp[ITEMS];
            *reinterpret_cast<float4 *>(tmp) = a;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) v[i] = to_f32<T>(tmp[i]);
        } else {
            static_assert(BYTES == 8, "unsupported vector width");
            float2 a = __ldg(reinterpret_cast<const float2 *>(row + pos));
            alignas(16) T tmp[ITEMS];
            *reinterpret_cast<float2 *>(tmp) = a;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) v[i] = to_f32<T>(tmp[i]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) v[i] = (pos + i < len) ? to_f32<T>(row[pos + i]) : fill;
    }
}

template <typename T, int ITEMS, bool VEC>
__device__ __forceinline__ void store_items(T *__restrict__ row, int pos, int len, const float (&v)[ITEMS]) {
    if (VEC && pos + ITEMS <= len) {
        constexpr int BYTES = ITEMS * sizeof(T);
        alignas(16) T tmp[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) tmp[i] = from_f32<T>(v[i]);
        if constexpr (BYTES == 32) {
            float4 *p = reinterpret_cast<float4 *>(row + pos);
            p[0] = *reinterpret_cast<float4 *>(tmp);
            p[1] = *reinterpret_cast<float4 *>(reinterpret_cast<char *>(tmp) + 16);
        } else if constexpr (BYTES == 16) {
            *reinterpret_cast<float4 *>(row + pos) = *reinterpret_cast<float4 *>(tmp);
        } else {
            *reinterpret_cast<float2 *>(row + pos) = *reinterpret_cast<float2 *>(tmp);
        }
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            if (pos + i < len) row[pos + i] = from_f32<T>(v[i]);
    }
}

// ---- fast, accuracy-checked transcendental pieces ---------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// softplus(x) with the reference's threshold (identity above 20, fwd_kernel.cuh:117) and its derivative
// sigmoid(x) (1 above the threshold, bwd_kernel.cuh:234-238).  MUFU based; relative error < 1e-6 over the
// whole range: for tiny e = exp(x) the log is replaced by its alternating series (log(1+e) would lose
// relative accuracy once 1+e rounds).
template <bool WITH_SIG>
__device__ __forceinline__ float softplus_sig(float x, float &sig) {
    const float e = ex2_approx(x * 1.4426950408889634f);
    const float one_pe = 1.0f + e;
    float sp = (e < 0.03125f) ? e * (1.0f - e * (0.5f - e * (0.33333334f - 0.25f * e)))
                              : lg2_approx(one_pe) * 0.6931471805599453f;
    if (WITH_SIG) sig = __fdividef(e, one_pe);
    if (x > 20.0f) {
        sp = x;
        if (WITH_SIG) sig = 1.0f;
    }
    return sp;
}

// ---- release / acquire accessors for the chunk-carry exchange --------------------------------------
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float2 ld_relaxed_f2(const float2 *p) {
    float2 v;
    asm volatile("ld.relaxed.gpu.global.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_f2(float2 *p, float2 v) {
    asm volatile("st.relaxed.gpu.global.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

}  // namespace vmasr
