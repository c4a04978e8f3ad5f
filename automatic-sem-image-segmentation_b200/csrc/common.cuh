// Shared helpers for the sm_100a kernels of libsemb200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/semb200.h"

namespace semb {

typedef __nv_bfloat16 bf16;

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  check_launch(const char* what);            // cudaGetLastError -> SEMB_ECUDA, counts the launch
extern long long g_launch_count;

#define SEMB_REQUIRE(cond, code, ...)            \
    do {                                         \
        if (!(cond)) {                           \
            semb::set_error(__VA_ARGS__);        \
            return (code);                       \
        }                                        \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch ------------------------------------------------------------
// A training step is ~400 dependent launches of 10-80 us kernels: without overlap every kernel boundary costs the launch
// latency plus the next kernel's prologue (barrier init, TMEM allocation, shared-memory zeroing, tensor-map prefetch).
// Kernels launched through launch_pdl() may start while their predecessor in the stream drains; they call pdl_wait()
// before their first access to global memory (it returns once the predecessor grid has completed and flushed), and
// pdl_trigger() at their top so that their own successor can be scheduled early.  SEMB_NO_PDL=1 launches them serially.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline long long cdivl(long long a, long long b) { return (a + b - 1) / b; }

// ---- 8-channel vector access (the unit every tensor is padded to) ----------------------------
// load8 / store8 move 8 consecutive channels between memory (T = float | bf16) and 8 fp32 registers.
template <typename T> struct Vec8;

template <> struct Vec8<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        const float4 a = *reinterpret_cast<const float4*>(p);
        const float4 b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};

template <> struct Vec8<bf16> {
    static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) {
        const uint4 raw = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

// 4-channel variants used by the SIMT conv tiles
template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        const float4 a = *reinterpret_cast<const float4*>(p);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec4<bf16> {
    static __device__ __forceinline__ void load(const bf16* p, float (&v)[4]) {
        const uint2 raw = *reinterpret_cast<const uint2*>(p);
        v[0] = __uint_as_float(raw.x << 16);
        v[1] = __uint_as_float(raw.x & 0xffff0000u);
        v[2] = __uint_as_float(raw.y << 16);
        v[3] = __uint_as_float(raw.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void store(bf16* p, const float (&v)[4]) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]);
        const __nv_bfloat162 h1 = __floats2bfloat162_rn(v[2], v[3]);
        uint2 raw;
        raw.x = *reinterpret_cast<const uint32_t*>(&h0);
        raw.y = *reinterpret_cast<const uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(p) = raw;
    }
};

__device__ __forceinline__ float act_fwd(float u, int act) {
    switch (act) {
        case SEMB_ACT_RELU: return fmaxf(u, 0.f);
        case SEMB_ACT_LEAKY: return u > 0.f ? u : 0.2f * u;
        case SEMB_ACT_SIGMOID: return 1.f / (1.f + expf(-u));
        case SEMB_ACT_TANH: return tanhf(u);
        default: return u;
    }
}
// derivative expressed through the OUTPUT y = act(u)
__device__ __forceinline__ float act_bwd_from_y(float y, int act) {
    switch (act) {
        case SEMB_ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case SEMB_ACT_LEAKY: return y > 0.f ? 1.f : 0.2f;
        case SEMB_ACT_SIGMOID: return y * (1.f - y);
        case SEMB_ACT_TANH: return 1.f - y * y;
        default: return 1.f;
    }
}

__device__ __forceinline__ int reflect_index(int i, int n) {
    // np.pad(mode="reflect"): edge not repeated.  Valid for -n < i < 2n-1.
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

static inline bool view_ok(const semb_tensor* t) {
    return t && t->ptr && t->C > 0 && (t->C % 8) == 0 && (t->pitch % 8) == 0 && (t->coff % 8) == 0 &&
           t->coff + t->C <= t->pitch && (reinterpret_cast<uintptr_t>(t->ptr) % 16) == 0;
}

}  // namespace semb
