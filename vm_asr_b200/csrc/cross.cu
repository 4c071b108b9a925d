// 4-direction cross scan / cross merge for sm_100a.
// Replaces CrossScan / CrossMerge (model/vmamba.py:27-73) and triton_cross_scan / triton_cross_merge
// (model/csm_triton.py:7-154).  One read of a 64x64 tile of the (H, W) map feeds all four directions:
//   k=0  l = h*W + w          k=1  l = w*H + h          k=2, k=3: the same two walked backwards (L-1-l)
// The row-major pair is moved with 128-bit accesses along w, the column-major pair with 128-bit accesses
// along h after a shared-memory transpose, so every global access of every direction is coalesced (the
// Triton kernel's transposed stores are stride-H).  Pure data movement (scan) / three additions in the
// reference's association (merge): bit-exact with the PyTorch versions.
#include "common.cuh"

namespace vmasr {

constexpr int kTile = 64;
constexpr int kPad = 65;  // shared-memory row stride (floats)

__device__ __forceinline__ float4 rev4(float4 v) { return make_float4(v.w, v.z, v.y, v.x); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}

// ---- fp32, H % 4 == 0 and W % 4 == 0, 16-byte aligned bases ---------------------------------------
__global__ void __launch_bounds__(256) cross_scan_vec4(const float *__restrict__ x, float *__restrict__ xs, int C, int H, int W) {
    __shared__ float s[kTile * kPad];
    const int bc = blockIdx.z;
    const int b = bc / C, c = bc - b * C;
    const long long L = (long long)H * W;
    const int h0 = blockIdx.y * kTile, w0 = blockIdx.x * kTile;
    const float *src = x + (long long)bc * L;
    float *d0 = xs + (((long long)b * 4 + 0) * C + c) * L;
    float *d1 = xs + (((long long)b * 4 + 1) * C + c) * L;
    float *d2 = xs + (((long long)b * 4 + 2) * C + c) * L;
    float *d3 = xs + (((long long)b * 4 + 3) * C + c) * L;
    const int q = threadIdx.x & 15, r0 = threadIdx.x >> 4;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int hl = p * 16 + r0, wl = q * 4;
        const int h = h0 + hl, w = w0 + wl;
        if (h < H && w < W) {
            const long long l = (long long)h * W + w;
            const float4 v = __ldg(reinterpret_cast<const float4 *>(src + l));
            *reinterpret_cast<float4 *>(d0 + l) = v;
            *reinterpret_cast<float4 *>(d2 + (L - 4 - l)) = rev4(v);
            float *row = s + hl * kPad + wl;
            row[0] = v.x; row[1] = v.y; row[2] = v.z; row[3] = v.w;
        }
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int wl = p * 16 + r0, hl = q * 4;
        const int h = h0 + hl, w = w0 + wl;
        if (h < H && w < W) {
            const float4 t = make_float4(s[hl * kPad + wl], s[(hl + 1) * kPad + wl], s[(hl + 2) * kPad + wl], s[(hl + 3) * kPad + wl]);
            const long long l = (long long)w * H + h;
            *reinterpret_cast<float4 *>(d1 + l) = t;
            *reinterpret_cast<float4 *>(d3 + (L - 4 - l)) = rev4(t);
        }
    }
}

__global__ void __launch_bounds__(256) cross_merge_vec4(const float *__restrict__ ys, float *__restrict__ y, int C, int H, int W) {
    __shared__ float s[kTile * kPad];
    const int bc = blockIdx.z;
    const int b = bc / C, c = bc - b * C;
    const long long L = (long long)H * W;
    const int h0 = blockIdx.y * kTile, w0 = blockIdx.x * kTile;
    const float *s0 = ys + (((long long)b * 4 + 0) * C + c) * L;
    const float *s1 = ys + (((long long)b * 4 + 1) * C + c) * L;
    const float *s2 = ys + (((long long)b * 4 + 2) * C + c) * L;
    const float *s3 = ys + (((long long)b * 4 + 3) * C + c) * L;
    float *dst = y + (long long)bc * L;
    const int q = threadIdx.x & 15, r0 = threadIdx.x >> 4;
    // column-major pair first: (ys1 + flip ys3), transposed into shared memory
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int wl = p * 16 + r0, hl = q * 4;
        const int h = h0 + hl, w = w0 + wl;
        if (h < H && w < W) {
            const long long l = (long long)w * H + h;
            const float4 a = __ldg(reinterpret_cast<const float4 *>(s1 + l));
            const float4 bb = rev4(__ldg(reinterpret_cast<const float4 *>(s3 + (L - 4 - l))));
            const float4 t = add4(a, bb);
            s[hl * kPad + wl] = t.x; s[(hl + 1) * kPad + wl] = t.y; s[(hl + 2) * kPad + wl] = t.z; s[(hl + 3) * kPad + wl] = t.w;
        }
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int hl = p * 16 + r0, wl = q * 4;
        const int h = h0 + hl, w = w0 + wl;
        if (h < H && w < W) {
            const long long l = (long long)h * W + w;
            const float4 a = __ldg(reinterpret_cast<const float4 *>(s0 + l));
            const float4 bb = rev4(__ldg(reinterpret_cast<const float4 *>(s2 + (L - 4 - l))));
            const float4 rowp = add4(a, bb);
            const float *row = s + hl * kPad + wl;
            const float4 colp = make_float4(row[0], row[1], row[2], row[3]);
            *reinterpret_cast<float4 *>(dst + l) = add4(rowp, colp);
        }
    }
}

// ---- any shape, any of the three dtypes: 32x32 tiles, one element per access ------------------------
template <typename T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
// fp16/bf16 + fp16/bf16 is exact in fp32, so one rounding of the fp32 sum equals the native half add
template <> __device__ __forceinline__ __half add_rn<__half>(__half a, __half b) { return __float2half_rn(__half2float(a) + __half2float(b)); }
template <> __device__ __forceinline__ __nv_bfloat16 add_rn<__nv_bfloat16>(__nv_bfloat16 a, __nv_bfloat16 b) {
    return __float2bfloat16_rn(__bfloat162float(a) + __bfloat162float(b));
}

template <typename T>
__global__ void __launch_bounds__(256) cross_scan_any(const T *__restrict__ x, T *__restrict__ xs, int C, int H, int W) {
    __shared__ T s[32][33];
    const int bc = blockIdx.z;
    const int b = bc / C, c = bc - b * C;
    const long long L = (long long)H * W;
    const int h0 = blockIdx.y * 32, w0 = blockIdx.x * 32;
    const T *src = x + (long long)bc * L;
    T *d0 = xs + (((long long)b * 4 + 0) * C + c) * L;
    T *d1 = xs + (((long long)b * 4 + 1) * C + c) * L;
    T *d2 = xs + (((long long)b * 4 + 2) * C + c) * L;
    T *d3 = xs + (((long long)b * 4 + 3) * C + c) * L;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int h = h0 + r, w = w0 + tx;
        if (h < H && w < W) {
            const long long l = (long long)h * W + w;
            const T v = src[l];
            d0[l] = v;
            d2[L - 1 - l] = v;
            s[r][tx] = v;
        }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int w = w0 + r, h = h0 + tx;
        if (h < H && w < W) {
            const long long l = (long long)w * H + h;
            const T v = s[tx][r];
            d1[l] = v;
            d3[L - 1 - l] = v;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) cross_merge_any(const T *__restrict__ ys, T *__restrict__ y, int C, int H, int W) {
    __shared__ T s[32][33];
    const int bc = blockIdx.z;
    const int b = bc / C, c = bc - b * C;
    const long long L = (long long)H * W;
    const int h0 = blockIdx.y * 32, w0 = blockIdx.x * 32;
    const T *s0 = ys + (((long long)b * 4 + 0) * C + c) * L;
    const T *s1 = ys + (((long long)b * 4 + 1) * C + c) * L;
    const T *s2 = ys + (((long long)b * 4 + 2) * C + c) * L;
    const T *s3 = ys + (((long long)b * 4 + 3) * C + c) * L;
    T *dst = y + (long long)bc * L;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int w = w0 + r, h = h0 + tx;
        if (h < H && w < W) {
            const long long l = (long long)w * H + h;
            s[tx][r] = add_rn<T>(s1[l], s3[L - 1 - l]);
        }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int h = h0 + r, w = w0 + tx;
        if (h < H && w < W) {
            const long long l = (long long)h * W + w;
            dst[l] = add_rn<T>(add_rn<T>(s0[l], s2[L - 1 - l]), s[r][tx]);
        }
    }
}

static int check_shape(const void *a, const void *b, int B, int C, int H, int W, int dtype, const char *who) {
    if (!a || !b) return fail("%s: null tensor", who);
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail("%s: sizes must be positive (B %d C %d H %d W %d)", who, B, C, H, W);
    if (dtype != VMASR_F32 && dtype != VMASR_F16 && dtype != VMASR_BF16) return fail("%s: unsupported dtype %d", who, dtype);
    if ((long long)B * C > 65535) return fail("%s: B*C = %lld exceeds the grid limit 65535", who, (long long)B * C);
    return 0;
}

template <bool MERGE>
static int run_cross(const void *in, void *out, int B, int C, int H, int W, int dtype, int device, void *stream_) {
    const char *who = MERGE ? "cross_merge" : "cross_scan";
    if (int rc = check_shape(in, out, B, C, H, W, dtype, who)) return rc;
    DeviceGuard guard(device);
    if (!guard.ok) return fail("%s: cannot select CUDA device %d", who, device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const bool vec = dtype == VMASR_F32 && (H % 4 == 0) && (W % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    if (vec) {
        dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, B * C);
        if (MERGE) cross_merge_vec4<<<grid, 256, 0, stream>>>(static_cast<const float *>(in), static_cast<float *>(out), C, H, W);
        else cross_scan_vec4<<<grid, 256, 0, stream>>>(static_cast<const float *>(in), static_cast<float *>(out), C, H, W);
    } else {
        dim3 grid((W + 31) / 32, (H + 31) / 32, B * C);
        switch (dtype) {
            case VMASR_F32:
                if (MERGE) cross_merge_any<float><<<grid, 256, 0, stream>>>(static_cast<const float *>(in), static_cast<float *>(out), C, H, W);
                else cross_scan_any<float><<<grid, 256, 0, stream>>>(static_cast<const float *>(in), static_cast<float *>(out), C, H, W);
                break;
            case VMASR_F16:
                if (MERGE) cross_merge_any<__half><<<grid, 256, 0, stream>>>(static_cast<const __half *>(in), static_cast<__half *>(out), C, H, W);
                else cross_scan_any<__half><<<grid, 256, 0, stream>>>(static_cast<const __half *>(in), static_cast<__half *>(out), C, H, W);
                break;
            default:
                if (MERGE) cross_merge_any<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(in), static_cast<__nv_bfloat16 *>(out), C, H, W);
                else cross_scan_any<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(in), static_cast<__nv_bfloat16 *>(out), C, H, W);
        }
    }
    return check_cuda(cudaGetLastError(), who);
}

}  // namespace vmasr

extern "C" int vmasr_cross_scan(const void *x, void *xs, int B, int C, int H, int W, int dtype, int device, void *stream) {
    return vmasr::run_cross<false>(x, xs, B, C, H, W, dtype, device, stream);
}
extern "C" int vmasr_cross_merge(const void *ys, void *y, int B, int C, int H, int W, int dtype, int device, void *stream) {
    return vmasr::run_cross<true>(ys, y, B, C, H, W, dtype, device, stream);
}
